#!/bin/bash
# ncu captures: launch list of a bench step and one --set full capture of a kernel.
# usage: KERNEL=stack_fwd bash scripts/gpu_profile.sh
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
K=${KERNEL:-stack_fwd}
MODE=${MODE:-fwd}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 2 \
    -f -o gpurun_out/prof_$K python scripts/profile_hot_path.py collab 3 $MODE > gpurun_out/prof_$K.log 2>&1
echo "ncu full exit $?" >> gpurun_out/prof_$K.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph \
    > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list exit $?" >> gpurun_out/bench_under_ncu.log
tail -3 gpurun_out/prof_$K.log
