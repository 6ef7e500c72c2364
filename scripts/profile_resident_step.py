"""Three training steps fed from a resident data set (FusedTrainer.step_resident) on
COLLAB-synth bs512, L2 flushed in between: run under
   ncu --metrics gpu__time_duration.sum --clock-control none   for the launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgcnn_b200 as dg
from dgcnn_b200.synth import CONFIGS, make_graphs

cfg = CONFIGS["collab"]
dev = torch.device("cuda:0")
graphs = make_graphs(cfg, 1024, seed=324)
ds = dg.DeviceDataset(graphs, dev, num_classes=cfg.num_classes)
torch.manual_seed(324)
model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).train()
trainer = dg.FusedTrainer(model)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rng = np.random.RandomState(0)
for rep in range(4):
    ids = rng.permutation(1024)[:512]
    ids_dev = ds.ids_to_device(ids)
    flush.zero_()
    torch.cuda.synchronize()
    stats = trainer.step_resident(ds, ids, ids_dev)
    torch.cuda.synchronize()
print("ok", float(stats[0]) / 512)
