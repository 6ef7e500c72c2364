#!/bin/bash
# Round 2, call C: where does the KS time go -- flush vs ring timing, ncu --set full of KS (cache-control
# all = cold, and none = warm instructions), warm per-graph trace.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 300 python scripts/ks_cold_vs_ring.py 6 30 > $D/ks_cold_vs_ring.log 2>&1
TRACE_WARM=1 timeout 300 python scripts/trace_stack_fwd.py collab > $D/trace_fwd_warm.log 2>&1
cp $D/trace_collab.txt $D/trace_collab_warm.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_fwd -s 1 -c 2 \
    -f -o $D/prof_ks_r2 python scripts/profile_hot_path.py collab 3 fwd > $D/prof_ks_r2.log 2>&1
echo "ncu full exit $?" >> $D/prof_ks_r2.log
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:stack_fwd -s 1 -c 2 \
    -f -o $D/prof_ks_r2_warm python scripts/profile_hot_path.py collab 3 fwd > $D/prof_ks_r2_warm.log 2>&1
echo "ncu full (warm) exit $?" >> $D/prof_ks_r2_warm.log
cat $D/ks_cold_vs_ring.log; grep "^#" $D/trace_collab_warm.txt | head -12; tail -2 $D/prof_ks_r2.log $D/prof_ks_r2_warm.log
