"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every C-ABI
entry point on tiny inputs, both hot-path implementations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dgcnn_b200 as dg
from dgcnn_b200.synth import CONFIGS, make_batch

dev = torch.device("cuda:0")
for name, count in (("mutag", 12), ("proteins", 6), ("collab", 6)):
    cfg = CONFIGS[name]
    hb = make_batch(name, num_graphs=count)
    data = hb.to(dev)
    data.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
    torch.manual_seed(0)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).eval()
    for fused in (True, False):
        dg.set_fused(fused)
        out = model(data)
        torch.nn.functional.nll_loss(out, data.y).backward()
        torch.cuda.synchronize()
        print(name, "fused" if fused else "per-layer", float(out.sum()))
print("sanitize workload done", dg.ops.LAUNCHES)
