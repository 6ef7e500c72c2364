#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_resident.py -q > gpurun_out/resident_tests.log 2>&1
echo "resident tests exit $?"
tail -25 gpurun_out/resident_tests.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/collate_launches.csv python scripts/profile_collate.py > gpurun_out/profile_collate.log 2>&1
echo "ncu exit $?"
grep -E "n1_|k0b" gpurun_out/collate_launches.csv | awk -F'","' '{print $5, $NF}' | tail -20
