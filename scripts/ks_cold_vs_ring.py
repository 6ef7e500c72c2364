"""KS forward (stack_fwd_mma_kernel) device time under the two legal ways of keeping the inputs
out of L2 (B200_PROFILING.md "timing hygiene"): (a) a 256 MiB write before every launch -- which
also evicts the kernel's own 100 KB of instructions -- and (b) a ring of distinct batches whose
combined footprint exceeds the 126 MB L2, no flush (data cold, instructions warm).  (c) same
batch every time, no flush: everything warm, for reference only.
    python scripts/ks_cold_vs_ring.py [ring=6] [reps=30]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dgcnn_b200 as dg
from dgcnn_b200.synth import CONFIGS, make_batch

ring = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0")
cfg = CONFIGS["collab"]
torch.manual_seed(324)
model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).eval()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
slots = []
with torch.no_grad():
    for i in range(ring):
        data = make_batch("collab", seed=324 + i).to(dev)
        g0 = model.build_graph(data)
        for _ in range(2):
            model.hot_path(data.x, g0)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True, external=True)
        b = torch.cuda.Event(enable_timing=True, external=True)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model.hot_path(data.x, g0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            a.record()
            out = model.hot_path(data.x, g0)
            b.record()
        n, e = data.x.size(0), data.edge_index.size(1)
        foot = 4 * (n + 1) + 4 * e // 17 + 4 * n + 4 * n + 4 * n * 100 + 4 * 512 * cfg.k * 98   # fragmaps ~ col/17
        slots.append((g, a, b, out, foot, data, g0))   # data / g0 stay alive: the captured graph reads them
print(f"ring {ring} batches, KS footprint per batch ~{slots[0][4] / 1e6:.0f} MB (outputs + inputs), "
      f"ring total ~{sum(s[4] for s in slots) / 1e6:.0f} MB vs 126 MB L2")


def run(mode):
    ts = []
    for it in range(reps + 2 * ring):
        g, a, b = slots[0 if mode in ("flush", "warm") else it % ring][:3]
        if mode == "flush":
            flush.zero_()
        g.replay()
        torch.cuda.synchronize()
        if it >= 2 * ring:
            ts.append(a.elapsed_time(b) * 1e3)
    return statistics.mean(ts), min(ts), statistics.median(ts)


for mode in ("flush", "ring", "warm", "ring", "flush"):
    m, lo, med = run(mode)
    print(f"{mode:6s}: KS {m:6.2f} us mean, {med:6.2f} median, {lo:6.2f} min")
