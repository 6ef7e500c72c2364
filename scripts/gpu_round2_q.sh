#!/bin/bash
# Round 2, call Q: K0 fast build with two positions per thread
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider > $D/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $D/pytest_gpu.log
tail -5 $D/pytest_gpu.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $D/bench_collab.json 2> $D/bench_collab.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_collab.json").read().strip().splitlines()[-1])
h=d["hot_path_fwd"]; r=d.get("e2e_resident_dataset") or {}
print("collab ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "fwd us", round(h["us"],1), "k0+k0b", round(h["graph_build_us"],1), "k0b", d["roofline"].get("k0b_us"), "resident", r.get("device_step_us"), "e2e", round(d["e2e"]["value"]))
PY
