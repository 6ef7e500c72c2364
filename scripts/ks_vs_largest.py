"""How much of KS / KSB / the training step is the largest graph of the batch?
The bench batch with its first graph replaced by a COLLAB-like graph of m nodes:
    python scripts/ks_vs_largest.py [m ...]   ->  gpurun_out/ks_vs_largest.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import dgcnn_b200 as dg
from bench import timed
from dgcnn_b200 import ops
from dgcnn_b200.synth import CONFIGS, collate, make_graphs, _gnm_pairs, _symmetrise_sorted

dev = torch.device("cuda:0")
cfg = CONFIGS["collab"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sizes = [int(a) for a in sys.argv[1:]] or [0, 256, 330, 400, 432, 448, 480, 492]
base = make_graphs(cfg, cfg.batch_size, seed=324)
lines = []
for m in sizes:
    graphs = list(base)
    if m:
        rng = np.random.RandomState(m)
        ei = _symmetrise_sorted(_gnm_pairs(rng, m, 33 * m), m)
        deg = np.bincount(ei[1], minlength=m).astype(np.float32)
        graphs[0] = {"x": (deg / deg.max())[:, None].astype(np.float32), "edge_index": ei, "y": 0}
    hb = collate(graphs)
    data = hb.to(dev)
    data.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
    torch.manual_seed(324)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).train()
    convs = (model.conv1, model.conv2, model.conv3, model.conv4)
    with torch.enable_grad():
        g_t = model.build_graph(data)
    with torch.no_grad():
        ws = [c.lin.weight for c in convs]
        bs = [c.bias for c in convs]
        t_ks = timed(lambda: model.hot_path(data.x, g_t), flush)
        pooled, xcat, perm = model.hot_path(data.x, g_t)
        dp = torch.randn_like(pooled)
        t_ksb = timed(lambda: ops.stack_bwd(dp, perm, xcat, data.x, g_t, ws, cfg.k, 0), flush)
        t_ks5 = t_ksb5 = float("nan")
        if ops.conv5_fusable(cfg.num_features, data.max_nodes):
            w5, b5 = model.conv5.weight, model.conv5.bias
            t_ks5 = timed(lambda: ops.stack_fwd_conv5(data.x, g_t, ws, bs, w5, b5, cfg.k, 0), flush)
            h1, arg, xcat5, perm5, _ = ops.stack_fwd_conv5(data.x, g_t, ws, bs, w5, b5, cfg.k, 0)
            dh1 = torch.randn_like(h1)
            t_ksb5 = timed(lambda: ops.stack_bwd_conv5(dh1, arg, perm5, xcat5, data.x, g_t, ws, w5, cfg.k, 0), flush)
    trainer = dg.FusedTrainer(model, lr=1e-3)
    before = ops.launches_total()
    trainer.step(data)
    torch.cuda.synchronize()
    launches = ops.launches_total() - before
    t_step = timed(lambda: trainer.step(data), flush)
    lines.append(f"largest {data.max_nodes:4d} nodes {hb.num_nodes} edges {hb.num_edges}: KS {t_ks*1e6:6.1f} KSB {t_ksb*1e6:6.1f} "
                 f"KS-conv5 {t_ks5*1e6:6.1f} KSB-conv5 {t_ksb5*1e6:6.1f} step {t_step*1e6:6.1f} us ({launches} launches)")
    print(lines[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/ks_vs_largest.txt", "w").write("\n".join(lines) + "\n")
