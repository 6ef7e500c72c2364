#!/bin/bash
# Round 2, call G: per-layer path rework (vectorised K1, project-first) -- tests, dd / powerlaw bench lines + launch lists.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider > $D/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $D/pytest_gpu.log
for w in dd powerlaw; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 3 > $D/bench_$w.json 2> $D/bench_$w.err
  echo "bench $w exit $?" >> $D/bench_$w.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/launches_$w.csv \
      python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-graph > $D/bench_${w}_ncu.log 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $D/bench_collab.json 2> $D/bench_collab.err
tail -15 $D/pytest_gpu.log
python - <<'PY'
import json
for w in ("collab","dd","powerlaw"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{w}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(w, "no bench line", e); continue
    h=d["hot_path_fwd"]; r=d["roofline"]
    print(w, "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "fwd us", round(h["us"],1), "frac", round(h["frac_of_peak"],3), "per-layer us", round(h["per_layer_kernel"]["launch_us"],1), round(h["per_layer_kernel"]["frac_of_peak"],3), "k0", round(h["graph_build_us"],1))
PY
