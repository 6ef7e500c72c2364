#!/bin/bash
# Round 2, call L: per-graph timelines of KSB, plain and conv5 variant
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 300 python scripts/trace_stack_bwd.py collab > $D/trace_bwd.log 2>&1; grep "^#" gpurun_out/trace_bwd_collab.txt
timeout 300 python scripts/trace_stack_bwd.py collab conv5 > $D/trace_bwd5.log 2>&1; grep "^#" gpurun_out/trace_bwd_collab_conv5.txt; tail -3 $D/trace_bwd5.log
