"""Per-graph phase timeline of the fused backward kernel (debug hook dgcnn_stack_bwd_set_trace).
    python scripts/trace_stack_bwd.py [workload]   ->  gpurun_out/trace_bwd_<workload>.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import dgcnn_b200 as dg
from dgcnn_b200 import _lib, ops
from dgcnn_b200.synth import CONFIGS, make_batch

name = sys.argv[1] if len(sys.argv) > 1 else "collab"
conv5 = len(sys.argv) > 2 and sys.argv[2] == "conv5"      # the SURVEY 8f N2 variant (fed with d(h1))
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
g0 = model.build_graph(data)
with torch.no_grad():
    pooled, xcat, perm = model.hot_path(data.x, g0)
    dp = torch.randn_like(pooled)
    weights = [c.lin.weight for c in (model.conv1, model.conv2, model.conv3, model.conv4)]
    if conv5:
        biases = [c.bias for c in (model.conv1, model.conv2, model.conv3, model.conv4)]
        h1, arg, xcat, perm, _ = ops.stack_fwd_conv5(data.x, g0, weights, biases, model.conv5.weight,
                                                     model.conv5.bias, cfg.k, 0)
        dh1 = torch.randn_like(h1)
        run = lambda: ops.stack_bwd_conv5(dh1, arg, perm, xcat, data.x, g0, weights, model.conv5.weight, cfg.k, 0)
    else:
        run = lambda: ops.stack_bwd(dp, perm, xcat, data.x, g0, weights, cfg.k, 0)
    for _ in range(3):
        run()
    flush.zero_()
    lib.dgcnn_stack_bwd_set_trace(trace.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.dgcnn_stack_bwd_set_trace(None)
t = trace.cpu().numpy()
meta = t[:, 15]
smid = meta >> 32
nthr = meta & 0xfff
n = (meta & 0xffffffff) >> 12
names = ["phase0", "L4", "A3", "B3", "C3", "A2", "B2", "C2", "A1", "B1", "C1", "out"]
os.makedirs("gpurun_out", exist_ok=True)
tag = name + ("_conv5" if conv5 else "")
with open(f"gpurun_out/trace_bwd_{tag}.txt", "w") as f:
    f.write("graph n threads smid " + " ".join("d_" + x for x in names) + " total\n")
    for gi in np.argsort(-n):
        row = t[gi, :13]
        d = np.diff(row)
        f.write(f"{gi} {n[gi]} {nthr[gi]} {smid[gi]} " + " ".join(str(int(x)) for x in d)
                + f" {int(row[12]-row[0])}\n")
    for thr in sorted(set(nthr.tolist())):
        sel = nthr == thr
        d = np.diff(t[sel, :13], axis=1).mean(0).astype(int)
        f.write(f"# threads {thr}: graphs {sel.sum()} n mean {n[sel].mean():.0f} phases {d.tolist()} "
                f"total {int((t[sel, 12] - t[sel, 0]).mean())}\n")
print(open(f"gpurun_out/trace_bwd_{tag}.txt").read()[:1500])
