#!/bin/bash
# Round 2, call E: N2 backward (KSB fed with d h1) parity.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline.py -m gpu -q --maxfail=10 --tb=short -p no:cacheprovider -k "conv5 or trainer or poisoned" > $D/pytest_n2.log 2>&1
echo "pytest exit $?" >> $D/pytest_n2.log
tail -40 $D/pytest_n2.log
