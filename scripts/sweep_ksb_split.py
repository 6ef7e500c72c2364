"""KSB (conv5 variant) and KS-conv5 time against the cluster-pair split threshold
(dgcnn_stack_fwd_configure sets it for both kernels):  python scripts/sweep_ksb_split.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dgcnn_b200 as dg
from bench import timed
from dgcnn_b200 import _lib, ops
from dgcnn_b200.synth import CONFIGS, make_batch

dev = torch.device("cuda:0")
lib = _lib.load_library()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cfg = CONFIGS["collab"]
for seed in (324, 326):                                  # largest graph 303 / 410 nodes
    hb = make_batch("collab", seed=seed)
    data = hb.to(dev)
    data.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
    torch.manual_seed(324)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).train()
    convs = (model.conv1, model.conv2, model.conv3, model.conv4)
    ws, bs = [c.lin.weight for c in convs], [c.bias for c in convs]
    w5, b5 = model.conv5.weight, model.conv5.bias
    with torch.enable_grad():
        g_t = model.build_graph(data)
    for pairs, pct in ((0, 80), (1, 10000), (1, 160), (1, 120), (1, 80), (1, 60), (1, 40), (1, 25)):
        lib.dgcnn_stack_fwd_configure(pairs, pct)
        with torch.no_grad():
            t_ks = timed(lambda: ops.stack_fwd_conv5(data.x, g_t, ws, bs, w5, b5, cfg.k, 0), flush, reps=12)
            h1, arg, xcat, perm, _ = ops.stack_fwd_conv5(data.x, g_t, ws, bs, w5, b5, cfg.k, 0)
            dh1 = torch.randn_like(h1)
            t_ksb = timed(lambda: ops.stack_bwd_conv5(dh1, arg, perm, xcat, data.x, g_t, ws, w5, cfg.k, 0), flush, reps=12)
        print(f"largest {data.max_nodes} pairs {pairs} split_pct {pct:5d}: KS-conv5 {t_ks*1e6:6.1f} us  KSB-conv5 {t_ksb*1e6:6.1f} us", flush=True)
lib.dgcnn_stack_fwd_configure(-1, 80)
