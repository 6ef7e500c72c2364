#!/bin/bash
# Round 2, call J: cluster-pair split in KSB + mandatory split of graphs beyond one CTA (KS and KSB)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline.py -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider -k "split or pair" > $D/pytest_split.log 2>&1
echo "pytest split exit $?" >> $D/pytest_split.log
tail -25 $D/pytest_split.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider > $D/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $D/pytest_gpu.log
tail -8 $D/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1
echo "smoke exit $?" >> $D/smoke.log; tail -2 $D/smoke.log
timeout 900 python scripts/ks_vs_largest.py > $D/ks_vs_largest.log 2>&1; tail -9 $D/ks_vs_largest.log
timeout 300 python scripts/trace_stack_bwd.py collab > $D/trace_bwd.log 2>&1; tail -6 gpurun_out/trace_bwd_collab.txt
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $D/bench_collab.json 2> $D/bench_collab.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_collab.json").read().strip().splitlines()[-1])
h=d["hot_path_fwd"]; print("collab ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "fwd us", round(h["us"],1), h.get("conv5_fused_variant"))
PY
