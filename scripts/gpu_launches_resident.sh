#!/bin/bash
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/resident_launches.csv python scripts/profile_resident_step.py > gpurun_out/profile_resident_step.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/profile_resident_step.log
