#!/bin/bash
# Round 2, last call: the whole GPU suite, smoke, the headline bench line and the reference arm on the final code
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider > $D/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $D/pytest_gpu.log; tail -4 $D/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1; echo "smoke exit $?" >> $D/smoke.log; tail -2 $D/smoke.log
timeout 900 python bench.py --steps 40 --warmup 5 > $D/bench_collab.json 2> $D/bench_collab.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_collab.json").read().strip().splitlines()[-1])
r=d.get("e2e_resident_dataset") or {}; h=d["hot_path_fwd"]
print("collab ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "KS us", round(h["us"],1), "frac", round(d["roofline"]["frac"],3), "k0+k0b", round(h["graph_build_us"],1), "launches", d["gpu_launches_per_step"])
print("resident dev us", r.get("device_step_us"), "device_value", r.get("device_value"), "e2e resident", r.get("value"), "driver epoch", r.get("driver_epoch_value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
print(open("gpurun_out/bench_reference.json").read().strip()[:300])
PY
