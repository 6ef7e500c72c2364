#!/bin/bash
# Round 2, call M: KSB conv5 variant with the d(h1) slab staged in shared memory
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider > $D/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $D/pytest_gpu.log
tail -5 $D/pytest_gpu.log
timeout 600 python scripts/ks_vs_largest.py 0 492 > $D/ks_vs_largest2.log 2>&1; tail -2 $D/ks_vs_largest2.log
timeout 300 python scripts/trace_stack_bwd.py collab conv5 > $D/trace_bwd5.log 2>&1; grep "^#" gpurun_out/trace_bwd_collab_conv5.txt
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $D/bench_collab.json 2> $D/bench_collab.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_collab.json").read().strip().splitlines()[-1])
h=d["hot_path_fwd"]; r=d.get("e2e_resident_dataset") or {}
print("collab ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "fwd us", round(h["us"],1), "resident", r.get("device_step_us"), r.get("value"), h.get("conv5_fused_variant"))
PY
