#!/bin/bash
# Round 2, call D: DSMEM bulk exchange + TMA zero fill + fused conv5 head -- check, tests, timings.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o /tmp/dsmem_check scripts/dsmem_bulk_check.cu > $D/dsmem_check.log 2>&1
timeout 60 /tmp/dsmem_check >> $D/dsmem_check.log 2>&1
echo "dsmem check exit $?" >> $D/dsmem_check.log
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -q --maxfail=10 --tb=short -p no:cacheprovider -x > $D/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $D/pytest_gpu.log
timeout 300 python scripts/sweep_ks_split.py collab proteins > $D/sweep_ks_split.log 2>&1
timeout 300 python scripts/ks_cold_vs_ring.py 6 30 > $D/ks_cold_vs_ring.log 2>&1
timeout 300 python scripts/trace_stack_fwd.py collab > $D/trace_fwd.log 2>&1
cat $D/dsmem_check.log; tail -15 $D/pytest_gpu.log; grep -v Warn $D/sweep_ks_split.log; grep -v Warn $D/ks_cold_vs_ring.log | tail -8
grep "^#" $D/trace_collab.txt | head -4; grep -v "^#" $D/trace_collab.txt | head -6
