#!/bin/bash
# tuning run: CTA size of the fused kernels (rebuilds the library on the GPU box per variant)
mkdir -p gpurun_out
for t in ${SWEEP:-512 640 768}; do
  DGCNN_NVCC_EXTRA="-DDGCNN_FWD_THREADS=${FWD:-640} -DDGCNN_BWD_THREADS=$t" python -c "from dgcnn_b200 import _lib; _lib.build_library(force=True)" > /dev/null 2>&1
  echo "== FWD_THREADS=${FWD:-640} BWD_THREADS=$t" >> gpurun_out/sweep.log
  timeout 300 python scripts/time_hot_path.py collab 20 2>&1 | grep "KS hot\|KSB" >> gpurun_out/sweep.log
done
cat gpurun_out/sweep.log
