"""KS forward time against the cluster-pair split setting (dgcnn_stack_fwd_configure):
    python scripts/sweep_ks_split.py [workload ...]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dgcnn_b200 as dg
from dgcnn_b200 import _lib
from dgcnn_b200.synth import CONFIGS, make_batch

dev = torch.device("cuda:0")
lib = _lib.load_library()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True, external=True)
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


for name in (sys.argv[1:] or ["collab", "proteins", "mutag"]):
    cfg = CONFIGS[name]
    hb = make_batch(name)
    data = hb.to(dev)
    torch.manual_seed(324)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).eval()
    with torch.no_grad():
        g0 = model.build_graph(data)
        sizes = sorted((hb.ptr[1:] - hb.ptr[:-1]).tolist(), reverse=True)[:8]
        print(f"{name}: graphs {hb.num_graphs} N {hb.num_nodes} E {hb.num_edges} largest {sizes}")
        for pairs, pct in ((0, 80), (1, 10000), (1, 120), (1, 100), (1, 80), (1, 60), (1, 40)):
            lib.dgcnn_stack_fwd_configure(pairs, pct)
            mean, best = timed(lambda: model.hot_path(data.x, g0))
            print(f"  pairs {pairs} split_pct {pct:5d}: KS {mean:6.1f} us (min {best:6.1f})")
    lib.dgcnn_stack_fwd_configure(-1, 80)
