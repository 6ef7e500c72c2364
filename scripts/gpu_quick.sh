#!/bin/bash
# quick GPU round: parity tests, hot-path kernel times, forward phase trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/time_hot_path.py collab 20 > gpurun_out/time_collab.log 2>&1
timeout 300 python scripts/trace_stack_fwd.py collab > gpurun_out/trace.log 2>&1
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/time_collab.log
