#!/bin/bash
# Round 2, call H: tests + smoke on the latest kernels, cost of the batch's largest graph (KS / KSB /
# step against its size), the KSB per-graph timeline, initcheck with plain-store padding.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider > $D/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $D/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1
echo "smoke exit $?" >> $D/smoke.log
timeout 900 python scripts/ks_vs_largest.py > $D/ks_vs_largest.log 2>&1
timeout 300 python scripts/trace_stack_bwd.py collab > $D/trace_bwd.log 2>&1
DGCNN_KS_PLAIN_ZERO=1 timeout 900 compute-sanitizer --tool initcheck python scripts/sanitize_small.py > $D/sanitize_initcheck_plain_zero.log 2>&1
echo "initcheck exit $?" >> $D/sanitize_initcheck_plain_zero.log
tail -4 $D/pytest_gpu.log; tail -2 $D/smoke.log; cat $D/ks_vs_largest.log | tail -12
grep -E "SUMMARY|exit" $D/sanitize_initcheck_plain_zero.log | tail -3
tail -5 gpurun_out/trace_bwd_collab.txt
