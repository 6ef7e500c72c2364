#!/bin/bash
# Round 2, call S: lean host side of driver.train_epoch (one H2D of the epoch's ids, vectorised plans)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_headline.py -m gpu -q --maxfail=10 --tb=short -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $D/bench_collab.json 2> $D/bench_collab.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_collab.json").read().strip().splitlines()[-1])
r=d.get("e2e_resident_dataset") or {}
print("collab ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "resident dev us", r.get("device_step_us"), "device_value", r.get("device_value"), "e2e resident", r.get("value"), "driver epoch", r.get("driver_epoch_value"))
PY
tail -3 $D/bench_collab.err | cut -c1-200
