#!/bin/bash
# Round 2, call O: PDL with the implicit trigger only (no early launch of the dependents), A/B
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
D=gpurun_out
for pdl in 1 0 1 0; do
  DGCNN_PDL=$pdl timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $D/bench_collab_pdl$pdl.json 2> $D/bench_collab_pdl$pdl.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_collab_pdl$pdl.json").read().strip().splitlines()[-1])
h=d["hot_path_fwd"]; r=d.get("e2e_resident_dataset") or {}
print("PDL=$pdl collab ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "fwd us", round(h["us"],1), "k0", round(h["graph_build_us"],1), "resident", r.get("device_step_us"), r.get("value"), "e2e", round(d["e2e"]["value"]))
PY
done
