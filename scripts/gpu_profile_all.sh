#!/bin/bash
# ncu evidence for profiles/: --set full captures of the forward and backward stack kernels,
# the K0 launch list, and the launch list of one eager training step of bench.py
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_fwd_mma -s 1 -c 2 \
    -f -o gpurun_out/prof_stack_fwd_mma python scripts/profile_hot_path.py collab 3 fwd > gpurun_out/prof_fwd.log 2>&1
echo "fwd exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_bwd_mma -s 1 -c 2 \
    -f -o gpurun_out/prof_stack_bwd_mma python scripts/profile_hot_path.py collab 3 bwd > gpurun_out/prof_bwd.log 2>&1
echo "bwd exit $?"
timeout 600 ncu --set full --clock-control none -k regex:k0_fast_build -s 1 -c 1 \
    -f -o gpurun_out/prof_k0_fast_build python scripts/profile_k0.py collab > gpurun_out/prof_k0.log 2>&1
echo "k0 exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/k0_launches.csv python scripts/profile_k0.py collab > gpurun_out/k0.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph \
    > gpurun_out/bench_under_ncu.log 2>&1
echo "list exit $?"
