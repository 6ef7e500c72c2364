#!/bin/bash
# gpurun_out/ (scratch) -> profiles/ (tracked): round-2 summaries.  Run here, after the GPU calls.
set -u
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles
S="python scripts/summarize_ncu.py"
[ -f $G/prof_ks_r2.ncu-rep ] && $S kernel $G/prof_ks_r2.ncu-rep $P/r02_stack_fwd_mma.md "KS stack_fwd_mma_kernel, round 2 (cluster pairs, DSMEM bulk exchange, TMA zero fill), writes pooled: COLLAB-synth bs512"
[ -f $G/prof_ks_conv5_r2.ncu-rep ] && $S kernel $G/prof_ks_conv5_r2.ncu-rep $P/r02_stack_fwd_conv5.md "KS with conv5 + ReLU + max-pool fused (dgcnn_stack_fwd_conv5), inside the resident training step"
[ -f $G/prof_ksb_conv5_r2.ncu-rep ] && $S kernel $G/prof_ksb_conv5_r2.ncu-rep $P/r02_stack_bwd_conv5.md "KSB fed with d(h1) (dgcnn_stack_bwd_conv5), inside the resident training step"
[ -f $G/prof_k1_staged_r2.ncu-rep ] && $S kernel $G/prof_k1_staged_r2.ncu-rep $P/r02_k1_staged.md "K1 gc_aggregate_staged, power-law bs256 (1000-node graphs), one 32->32 layer"
[ -f $G/resident_launches.csv ] && $S launches $G/resident_launches.csv $P/r02_launches_resident_step.md "Training steps fed from the resident data set (dgcnn_train_step_resident), COLLAB-synth bs512, round 2"
for w in dd powerlaw; do
  [ -f $G/launches_fwd_$w.csv ] && $S launches $G/launches_fwd_$w.csv $P/r02_launches_fwd_$w.md "Forward hot path (K0 + per-layer kernels + K2) on $w-synth at BASELINE size, round 2"
done
{
  echo "# Source-line hot spots, round 2 (ncu --set full --import-source on, COLLAB-synth bs512)"
  echo
  for k in ks_r2 ks_conv5_r2 ksb_conv5_r2 k1_staged_r2; do
    [ -f $G/prof_$k.ncu-rep ] && { echo "## $k"; echo; python scripts/ncu_hotspots.py $G/prof_$k.ncu-rep 14; echo; }
  done
} > $P/r02_hotspots.md
for w in collab dd powerlaw proteins mutag; do
  [ -s $G/bench_$w.json ] && tail -1 $G/bench_$w.json > $P/r02_bench_$w.json
done
[ -s $G/bench_reference.json ] && tail -1 $G/bench_reference.json > $P/r02_bench_reference.json
for n in 2 4 8; do
  for w in collab powerlaw collab_indep collab_balanced collab_independent collab_same; do
    [ -s $G/bench_${w}_n$n.json ] && tail -1 $G/bench_${w}_n$n.json > $P/r02_bench_${w}_n$n.json
    [ -s $G/exchange_trace_${w}_n$n.json ] && python - <<PY
import json
d=json.load(open("$G/exchange_trace_${w}_n$n.json"))
d["rows"]=d["rows"][:8]
json.dump(d, open("$P/r02_exchange_trace_${w}_n$n.json","w"), indent=1)
PY
  done
  for c in check_p2p check_dp_resident; do
    [ -s $G/${c}_n$n.log ] && grep -v "CUDAEvent\|Warning\|OMP_NUM\|\*\*\*\*" $G/${c}_n$n.log | tail -4 > $P/r02_${c}_n$n.log
  done
done
[ -s $G/ks_vs_largest.txt ] && cp $G/ks_vs_largest.txt $P/r02_ks_vs_largest.txt
[ -s $G/resident_launches_warm.md ] && cp $G/resident_launches_warm.md $P/r02_launches_resident_step_warm.md
[ -s $G/sanitize_initcheck_plain_zero.log ] && grep -E "SUMMARY|exit|sanitize workload|round-2 paths|Error" $G/sanitize_initcheck_plain_zero.log | head -20 > $P/r02_sanitize_initcheck_plain_zero.log
for t in memcheck racecheck synccheck initcheck; do
  [ -s $G/sanitize_$t.log ] && grep -E "SUMMARY|exit|sanitize workload|round-2 paths|Error|Hazard" $G/sanitize_$t.log | head -20 > $P/r02_sanitize_$t.log
done
[ -s $G/smoke_ncu.log ] && { tail -2 $G/smoke_ncu.log; grep -c dgcnn $G/smoke_ncu.csv; } > $P/r02_smoke_under_ncu_driver_command.log
ls $P | grep r02
