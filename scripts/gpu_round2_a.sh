#!/bin/bash
# Round 2, call A: all GPU tests, smoke plain + under ncu (the driver's command), bench lines of every
# BASELINE workload as the code stands, host-oracle diagnosis.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > $D/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $D/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1
echo "smoke exit $?" >> $D/smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $D/smoke_ncu.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke_ncu.log 2>&1
echo "smoke under ncu exit $?" >> $D/smoke_ncu.log
timeout 120 python scripts/diag_oracle_cpu.py 300 > $D/diag_oracle_cpu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 10 --csv --log-file $D/diag_oracle_ncu.csv \
    python scripts/diag_oracle_cpu.py 300 > $D/diag_oracle_cpu_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 3 > $D/bench_collab.json 2> $D/bench_collab.err
for w in dd powerlaw proteins mutag; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 3 > $D/bench_$w.json 2> $D/bench_$w.err
  echo "bench $w exit $?" >> $D/bench_$w.err
done
tail -15 $D/pytest_gpu.log; tail -3 $D/smoke.log; tail -3 $D/smoke_ncu.log
cat $D/diag_oracle_cpu.log; grep -v "^==" $D/diag_oracle_cpu_ncu.log
for w in collab dd powerlaw proteins mutag; do echo "== $w"; cut -c1-1500 $D/bench_$w.json; tail -2 $D/bench_$w.err; done
