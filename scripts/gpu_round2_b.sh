#!/bin/bash
# Round 2, call B: the reworked KS (cluster-pair split, in-team zero fill) -- tests, split sweep, trace, bench.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > $D/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $D/pytest_gpu.log
timeout 300 python scripts/sweep_ks_split.py collab proteins mutag > $D/sweep_ks_split.log 2>&1
timeout 300 python scripts/trace_stack_fwd.py collab > $D/trace_fwd.log 2>&1
timeout 300 python scripts/time_hot_path.py collab > $D/time_collab.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 3 > $D/bench_collab.json 2> $D/bench_collab.err
tail -12 $D/pytest_gpu.log; cat $D/sweep_ks_split.log; tail -8 $D/time_collab.log
grep "^#" $D/trace_collab.txt | head -12
cut -c1-600 $D/bench_collab.json; tail -3 $D/bench_collab.err
