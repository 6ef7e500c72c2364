"""Launch list of the resident-data-set gather (N1) on one COLLAB-synth batch: run under
   ncu --metrics gpu__time_duration.sum --clock-control none  (scripts/gpu_resident.sh)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgcnn_b200 as dg
from dgcnn_b200.synth import CONFIGS, make_graphs

cfg = CONFIGS["collab"]
dev = torch.device("cuda:0")
graphs = make_graphs(cfg, 1024, seed=324)
ds = dg.DeviceDataset(graphs, dev, num_classes=cfg.num_classes)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rng = np.random.RandomState(0)
for rep in range(3):
    ids = rng.permutation(1024)[:512]
    ids_dev = ds.ids_to_device(ids)
    flush.zero_()
    torch.cuda.synchronize()
    rb = ds.batch(ids, ids_dev, bitmaps=(rep == 2))
    torch.cuda.synchronize()
print("ok", rb.num_nodes)
