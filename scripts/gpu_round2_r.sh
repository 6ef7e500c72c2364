#!/bin/bash
# Round 2, call R: the driver's own N-GPU bench command on the final code (default flags)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 3 > $D/bench_driver_n$N.json 2> $D/bench_driver_n$N.err
echo "exit $?"
python - <<PY
import json
d=json.loads(open("$D/bench_driver_n$N.json").read().strip().splitlines()[-1])
print("N=$N ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "first", round(d["first_timed_step_ms"],4), "steady", round(d["ms_per_step_after_first"],4), "e2e", round(d["e2e"]["value"]), "params_equal", d["params_equal_across_ranks"], d["comm_status_per_rank"], d["clocks"], d["gpu_launches_per_step"])
PY
tail -3 $D/bench_driver_n$N.err | cut -c1-200
