#!/bin/bash
# smoke()'s per-layer path under ncu / sanitizers, bit-for-bit against a plain run.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
timeout 300 python scripts/diag_smoke.py plain > $D/diag_plain.log 2>&1
timeout 300 python scripts/diag_smoke.py poison poison > $D/diag_poison.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum -c 1000 --csv --log-file $D/diag_ncu_default.csv \
    python scripts/diag_smoke.py ncu_default > $D/diag_ncu_default.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $D/diag_ncu_noclock.csv \
    python scripts/diag_smoke.py ncu_noclock > $D/diag_ncu_noclock.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1000 --csv \
    --log-file $D/diag_ncu_nocache.csv python scripts/diag_smoke.py ncu_nocache > $D/diag_ncu_nocache.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $D/smoke_ncu.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke_ncu.log 2>&1
echo "smoke under ncu exit $?" >> $D/smoke_ncu.log
for tool in initcheck racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-name kns=dgcnn python scripts/diag_smoke.py san_$tool \
      > $D/diag_san_$tool.log 2>&1
  echo "$tool exit $?" >> $D/diag_san_$tool.log
done
for f in plain poison ncu_default ncu_noclock ncu_nocache; do echo "== $f"; grep -v Warning $D/diag_$f.log | tail -14; done
tail -4 $D/smoke_ncu.log
for tool in initcheck racecheck synccheck memcheck; do echo "== $tool"; tail -6 $D/diag_san_$tool.log; done
