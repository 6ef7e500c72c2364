#!/bin/bash
# forward-kernel iteration: parity tests (-x), kernel times, phase trace, one ncu --set full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/time_hot_path.py collab 20 > gpurun_out/time_collab.log 2>&1
timeout 300 python scripts/trace_stack_fwd.py collab > gpurun_out/trace.log 2>&1
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-stack_fwd_mma} -s 1 -c 1 \
    -f -o gpurun_out/prof_ks python scripts/profile_hot_path.py collab 3 fwd > gpurun_out/prof_ks.log 2>&1
fi
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/time_collab.log
