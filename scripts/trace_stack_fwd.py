"""Per-graph phase timeline of the fused forward kernel (debug hook dgcnn_stack_fwd_set_trace).
    python scripts/trace_stack_fwd.py [workload]   ->  gpurun_out/trace_<workload>.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import dgcnn_b200 as dg
from dgcnn_b200 import _lib
from dgcnn_b200.synth import CONFIGS, make_batch

name = sys.argv[1] if len(sys.argv) > 1 else "collab"
dev = torch.device("cuda:0")
cfg = CONFIGS[name]
hb = make_batch(name)
data = hb.to(dev)
data.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
torch.manual_seed(324)
model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).eval()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load_library()
b = hb.num_graphs
trace = torch.zeros(b, 16, dtype=torch.int64, device=dev)
with torch.no_grad():
    g0 = model.build_graph(data)
    for _ in range(3):
        model.hot_path(data.x, g0)
    if os.environ.get("TRACE_WARM", "0") != "1":
        flush.zero_()                                  # cold L2 (data AND the kernel's instructions)
    lib.dgcnn_stack_fwd_set_trace(trace.data_ptr())
    model.hot_path(data.x, g0)
    torch.cuda.synchronize()
    lib.dgcnn_stack_fwd_set_trace(None)
t = trace.cpu().numpy()
meta = t[:, 15]
smid = meta >> 32
nthr = meta & 0xfff
n = (meta & 0xffffffff) >> 12
t0 = t[:, 0].min()
names = ["start", "phase0", "xs", "L1", "L2", "L3", "L4", "sort", "gather"]
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/trace_{name}.txt", "w") as f:
    f.write("# cycles (clock64 of the SM; SMs are not synchronised, 'start' is relative to the earliest)\n")
    f.write("graph n threads smid start " + " ".join("d_" + x for x in names[1:]) + " total\n")
    order = np.argsort(-n)
    for gi in order:
        row = t[gi, :9]
        d = np.diff(row)
        f.write(f"{gi} {n[gi]} {nthr[gi]} {smid[gi]} {row[0]-t0} " + " ".join(str(int(x)) for x in d)
                + f" {int(row[8]-row[0])}\n")
    ns0 = t[:, 11].min()
    f.write(f"# globaltimer (ns): first CTA start 0, last CTA start {t[:, 11].max() - ns0}, "
            f"first graph start {t[:, 9].min() - ns0}, last graph end {t[:, 10].max() - ns0}\n")
    late = np.argsort(-t[:, 10])[:8]
    for gi in late:
        f.write(f"#   late finisher: graph {gi} n {n[gi]} threads {nthr[gi]} sm {smid[gi]} "
                f"start {t[gi, 9] - ns0} end {t[gi, 10] - ns0} ns\n")
    big = int(np.argmax(n))
    f.write(f"# CTA prologue (cycles, mean | largest graph): entry->plan+weights "
            f"{np.mean(t[:, 13] - t[:, 12]):.0f} | {t[big, 13] - t[big, 12]}, ->padding zeroed "
            f"{np.mean(t[:, 14] - t[:, 13]):.0f} | {t[big, 14] - t[big, 13]}, ->graph start "
            f"{np.mean(t[:, 0] - t[:, 14]):.0f} | {t[big, 0] - t[big, 14]}\n")
    tot = t[:, 8] - t[:, 0]
    f.write(f"# sum over graphs of total cycles: {tot.sum()}  mean {tot.mean():.0f}  max {tot.max()}\n")
    for s in sorted(set(smid.tolist())):
        sel = smid == s
        f.write(f"# sm {s}: graphs {sel.sum()} span {t[sel, 8].max() - t[sel, 0].min()} busy {tot[sel].sum()}\n")
print(open(f"gpurun_out/trace_{name}.txt").read()[:6000])
