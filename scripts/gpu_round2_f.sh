#!/bin/bash
# Round 2, call F: everything after the N2 wiring -- all GPU tests, smoke, bench with / without the fusion.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider > $D/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $D/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1
echo "smoke exit $?" >> $D/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 3 > $D/bench_collab.json 2> $D/bench_collab.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fuse-conv5 > $D/bench_collab_nofuse.json 2> $D/bench_collab_nofuse.err
timeout 300 python scripts/time_hot_path.py collab > $D/time_collab.log 2>&1
tail -25 $D/pytest_gpu.log; tail -3 $D/smoke.log
for f in bench_collab bench_collab_nofuse; do python - <<PY
import json
d=json.loads(open("$D/$f.json").read().strip().splitlines()[-1])
r=d.get("e2e_resident_dataset") or {}
print("$f", "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches_per_step"],
      "resident step us", r.get("device_step_us"), "resident e2e", r.get("value"), "driver epoch", r.get("driver_epoch_value"), "roofline us", d["roofline"]["launch_us"], d["roofline"]["frac"])
PY
done
tail -3 $D/bench_collab.err; tail -6 $D/time_collab.log
