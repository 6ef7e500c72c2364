#!/bin/bash
# Round 2, call K: warm-cache launch list of the resident training step (ncu --cache-control none),
# KS after the plan's loads were hoisted
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $D/resident_launches_warm.csv \
    python scripts/profile_resident_step.py > $D/profile_resident_step_warm.log 2>&1
python scripts/summarize_ncu.py launches $D/resident_launches_warm.csv $D/resident_launches_warm.md "resident step, warm caches (ncu --cache-control none)"
head -40 $D/resident_launches_warm.md
timeout 600 python scripts/ks_vs_largest.py 0 492 > $D/ks_vs_largest2.log 2>&1; tail -3 $D/ks_vs_largest2.log
timeout 600 python -m pytest tests/test_gpu_headline.py -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider -k "split or pair" 2>&1 | tail -3
