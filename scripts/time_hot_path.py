"""Device time of the hot-path kernels on one synthetic batch (CUDA-graph replay, L2 flushed
before every replay, CUDA events): K0 graph build, fused forward, fused backward.
    python scripts/time_hot_path.py [workload] [reps]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dgcnn_b200 as dg
from dgcnn_b200 import ops
from dgcnn_b200.synth import CONFIGS, make_batch

name = sys.argv[1] if len(sys.argv) > 1 else "collab"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
cfg = CONFIGS[name]
hb = make_batch(name)
data = hb.to(dev)
data.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
torch.manual_seed(324)
model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).eval()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True, external=True)     # event nodes inside the graph
    b = torch.cuda.Event(enable_timing=True, external=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        a.record()
        fn()
        b.record()
    ts = []
    for _ in range(reps):
        flush.zero_()
        g.replay()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.mean(ts), min(ts)


with torch.enable_grad():                       # training-mode graph: also A_hat^T for backward
    g0 = model.build_graph(data)
    t_k0 = timed(lambda: model.build_graph(data))
    t_k0a = timed(lambda: ops.build_graph(data.edge_index, data.batch, data.x.size(0), hb.num_graphs,
                                          transpose=True, max_nodes=0))
    t_k0n = timed(lambda: ops.build_graph(data.edge_index, data.batch, data.x.size(0), hb.num_graphs,
                                          transpose=False, max_nodes=data.max_nodes))
with torch.no_grad():
    print("workload", name, "N", hb.num_nodes, "E", hb.num_edges, "max_nodes", data.max_nodes)
    print("K0 build_graph  us mean/min: %.1f %.1f" % t_k0)
    print("   K0 only (no bitmaps)     : %.1f %.1f" % t_k0a)
    print("   K0 + K0b, no transpose   : %.1f %.1f" % t_k0n)
    print("KS hot_path fwd us mean/min: %.1f %.1f" % timed(lambda: model.hot_path(data.x, g0)))
    pooled, xcat, perm = model.hot_path(data.x, g0)
    dp = torch.randn_like(pooled)
    weights = [c.lin.weight for c in (model.conv1, model.conv2, model.conv3, model.conv4)]
    if ops.stack_bwd_supported(cfg.num_features, data.max_nodes) and g0.rowptr_t is not None:
        print("KSB stack_bwd   us mean/min: %.1f %.1f" % timed(
            lambda: ops.stack_bwd(dp, perm, xcat, data.x, g0, weights, cfg.k, 0)))
