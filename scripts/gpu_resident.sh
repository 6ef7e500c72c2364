#!/bin/bash
# Round-1 N1/N4 check on the GPU box: resident-data-set tests, smoke, one bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_resident.py -x -q > gpurun_out/resident_tests.log 2>&1
echo "resident tests exit $?"
tail -15 gpurun_out/resident_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"
tail -5 gpurun_out/smoke.log
timeout 200 python bench.py --steps 20 --warmup 3 --cpu-seconds 3 > gpurun_out/bench_resident.json 2> gpurun_out/bench_resident.err
echo "bench exit $?"
tail -3 gpurun_out/bench_resident.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_resident.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "resident", d.get("e2e_resident_dataset"))
except Exception as e:
    print("no bench line", e)
PY
