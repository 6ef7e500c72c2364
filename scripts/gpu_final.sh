#!/bin/bash
# End-of-round check: the full GPU suite, the default bench line, one ncu capture of n1_gather.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 240 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -2 gpurun_out/bench.err
timeout 120 ncu --set full --clock-control none --import-source on -k regex:n1_gather -s 2 -c 1 -o gpurun_out/prof_n1_gather -f python scripts/profile_collate.py > gpurun_out/prof_n1.log 2>&1
echo "ncu exit $?"
