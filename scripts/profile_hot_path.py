"""Short workload for ncu: graph build + hot-path forward (+ backward) on one COLLAB-synth
batch, a few repetitions, L2 flushed between them.  No timing is reported from here."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dgcnn_b200 as dg
from dgcnn_b200.synth import CONFIGS, make_batch

name = sys.argv[1] if len(sys.argv) > 1 else "collab"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
backward = len(sys.argv) > 3 and sys.argv[3] == "bwd"
dev = torch.device("cuda:0")
cfg = CONFIGS[name]
hb = make_batch(name)
data = hb.to(dev)
data.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
torch.manual_seed(324)
model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(reps):
    flush.zero_()
    g = model.build_graph(data)
    if backward:
        pooled, _, _ = model.hot_path(data.x, g)
        pooled.sum().backward()
    else:
        with torch.no_grad():
            model.hot_path(data.x, g)
torch.cuda.synchronize()
print("done", dg.ops.LAUNCHES)
