#!/bin/bash
# N-GPU diagnostic: where does the step time go at N ranks?  Independent batches against the same
# batches on every rank, each with the exchange trace (per-rank compute / wait) and per-rank clocks.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nproc > $D/host_n$N.txt; nvidia-smi topo -m >> $D/host_n$N.txt 2>&1
for sh in ${SHARDS:-independent same}; do
  timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-resident \
      --shards $sh --trace-exchange $D/exchange_trace_collab_${sh}_n$N.json > $D/bench_collab_${sh}_n$N.json 2> $D/bench_collab_${sh}_n$N.err
  echo "bench $sh n$N exit $?" >> $D/bench_collab_${sh}_n$N.err
done
python - <<PY
import json
import os
for sh in os.environ.get("SHARDS", "independent same").split():
    try:
        d=json.loads(open("$D/bench_collab_%s_n$N.json"%sh).read().strip().splitlines()[-1])
        print(sh, "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "first", d.get("first_timed_step_ms"), "steady", d.get("ms_per_step_after_first"))
        for r in d["per_rank"]: print("   ", r)
        t=json.load(open("$D/exchange_trace_collab_%s_n$N.json"%sh))
        print("   compute/rank", t.get("mean_compute_us_per_rank"), "max", t.get("mean_of_max_compute_us"), "mean", t.get("mean_of_mean_compute_us"))
        print("   wait/rank", t["mean_wait_us_per_rank"], "push", t["mean_push_us"], "sum+adam", t["mean_sum_adam_us"])
    except Exception as e:
        print(sh, "failed", e)
PY
head -3 $D/host_n$N.txt
