"""ncu workload: K0 + K0b (graph build) on one COLLAB-synth batch, L2 flushed between repetitions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dgcnn_b200 as dg
from dgcnn_b200.synth import CONFIGS, make_batch

name = sys.argv[1] if len(sys.argv) > 1 else "collab"
dev = torch.device("cuda:0")
cfg = CONFIGS[name]
hb = make_batch(name)
data = hb.to(dev)
data.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    g = model.build_graph(data)
torch.cuda.synchronize()
print("done")
