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
# N1: resident data set -- dgcnn_dataset_prepare, n1_gather (fused tables), n1_plan + gather
# (> 1024 graphs per batch), a symmetric and a generic (multigraph) data set, one resident step
import numpy as np
from dgcnn_b200.synth import make_graphs
dg.set_fused(True)
cfg = CONFIGS["mutag"]
graphs = make_graphs(cfg, 24, seed=1)
ds = dg.DeviceDataset(graphs, dev, num_classes=cfg.num_classes)
rb = ds.batch(np.array([3, 3, 0, 23, 7]))
big = ds.batch(np.random.RandomState(0).randint(0, 24, size=1100), bitmaps=False)
rng = np.random.RandomState(2)
multi = [{"x": rng.standard_normal((n, 3)).astype(np.float32),
          "edge_index": np.stack([rng.randint(0, max(n, 1), 3 * n), rng.randint(0, max(n, 1), 3 * n)]).astype(np.int64),
          "y": 0} for n in (5, 0, 7, 1, 9)]
gds = dg.DeviceDataset(multi, dev, num_classes=2)
gb = gds.batch(np.array([4, 1, 0, 2, 2]))
model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).train()
trainer = dg.FusedTrainer(model)
stats = trainer.step_resident(ds, np.arange(12))
torch.cuda.synchronize()
print("resident", rb.num_nodes, big.num_nodes, gb.num_nodes, float(stats[0]))
# round 2: the per-layer path with the batch's graph offsets (gc_aggregate_staged needs >= 96 graphs;
# one graph above its 1728-row cap goes to the row-parallel kernel), K2's radix select (a graph with
# more than 4k nodes), the conv5 fusion on and off, the cluster-split KS (graphs above 112 nodes when
# CTAs are spare: the collab batches above)
from dgcnn_b200.synth import collate
cfgp = CONFIGS["proteins"]
many = make_graphs(cfgp, 100, seed=3, tie_free=True)
rngb = np.random.RandomState(5)
for nbig in (1800, 1400):
    pr = np.unique(np.sort(rngb.randint(0, nbig, (3 * nbig, 2)), 1), axis=0)
    pr = pr[pr[:, 0] != pr[:, 1]]
    src = np.concatenate([pr[:, 0], pr[:, 1]]); dst = np.concatenate([pr[:, 1], pr[:, 0]])
    order = np.argsort(src * nbig + dst, kind="stable")
    many.append({"x": rngb.standard_normal((nbig, cfgp.num_features)).astype(np.float32),
                 "edge_index": np.stack([src[order], dst[order]]), "y": 0})
hb = collate(many)
data = hb.to(dev)
torch.manual_seed(1)
model = dg.Model(cfgp.num_features, cfgp.num_classes, 30).to(dev).eval()
out = model(data)
torch.nn.functional.nll_loss(out, data.y).backward()
for fuse in (False, True):
    dg.ops.set_fuse_conv5(fuse)
    hb2 = make_batch("collab", num_graphs=6)
    d2 = hb2.to(dev)
    m2 = dg.Model(1, 3, 130).to(dev).train()
    tr = dg.FusedTrainer(m2)
    tr.step(d2)
    out2 = m2(d2)
    torch.nn.functional.nll_loss(out2, d2.y).backward()
torch.cuda.synchronize()
print("round-2 paths", float(out.sum()), float(out2.sum()))
print("sanitize workload done", dg.ops.LAUNCHES)
