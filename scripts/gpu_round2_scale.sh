#!/bin/bash
# Multi-GPU call (gpurun --gpus N): correctness checks, then the scaling bench lines.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
N=${N:-2}
W=${WORKLOADS:-collab}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 scripts/check_p2p_allreduce.py > $D/check_p2p_n$N.log 2>&1
echo "check_p2p exit $?" >> $D/check_p2p_n$N.log
timeout 300 $TR --master-port 29513 scripts/check_dp_resident.py > $D/check_dp_resident_n$N.log 2>&1
echo "check_dp_resident exit $?" >> $D/check_dp_resident_n$N.log
for w in $W; do
  timeout 900 $TR --master-port 29514 bench.py --gpus $N --workload $w --steps 20 --warmup 3 --no-cpu-baseline \
      --trace-exchange $D/exchange_trace_${w}_n$N.json > $D/bench_${w}_n$N.json 2> $D/bench_${w}_n$N.err
  echo "bench $w n$N exit $?" >> $D/bench_${w}_n$N.err
done
if [ "${INDEP:-1}" = "1" ]; then
  timeout 900 $TR --master-port 29515 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-resident \
      --shards independent --trace-exchange $D/exchange_trace_collab_indep_n$N.json > $D/bench_collab_indep_n$N.json 2> $D/bench_collab_indep_n$N.err
fi
grep -v Warn $D/check_p2p_n$N.log | tail -3; grep -v Warn $D/check_dp_resident_n$N.log | tail -5
for f in $D/bench_*_n$N.json; do python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    r=d.get("e2e_resident_dataset") or {}
    print("$f", "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "resident dev us", r.get("device_step_us"), "resident e2e", r.get("value"), "params_equal", d.get("params_equal_across_ranks"), d.get("comm_status_per_rank"), d.get("shards"))
except Exception as e:
    print("$f", "failed", e)
PY
done
for f in $D/exchange_trace_*_n$N.json; do python - <<PY
import json
d=json.load(open("$f")); print("$f", "compute/rank", d.get("mean_compute_us_per_rank"), "max", d.get("mean_of_max_compute_us"), "mean", d.get("mean_of_mean_compute_us"), "wait/rank", d["mean_wait_us_per_rank"], "push", d["mean_push_us"], "sum+adam", d["mean_sum_adam_us"], "skew", d["mean_enter_skew_us"])
PY
done
tail -3 $D/bench_collab_n$N.err
