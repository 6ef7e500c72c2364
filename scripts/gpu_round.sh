#!/bin/bash
# One gpurun call: parity tests, smoke, bench, ncu launch list (+ optional sanitizer).
# Everything is wrapped in `timeout`; logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
if [ "${SANITIZE:-0}" = "1" ]; then
  for tool in memcheck racecheck synccheck; do
    timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=dgcnn python scripts/sanitize_small.py \
        > gpurun_out/sanitize_$tool.log 2>&1
    echo "$tool exit $?" >> gpurun_out/sanitize_$tool.log
  done
fi
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph \
    > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/bench_under_ncu.log
fi
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for f in gpurun_out/sanitize_*.log; do [ -f "$f" ] && tail -4 "$f"; done
