"""Which CPU op of the fp32 oracle is host-dependent?  (profiles/r02_smoke_under_ncu.md: on the
GPU box's host the float32 oracle sometimes lands 5e-5 from the float64 oracle in ONE 56-row
chunk, while the GPU result and the float64 oracle are bit-stable.)  Repeats each torch CPU op
of oracle.gcn_conv in float32 and counts the runs that leave the float64 result by > 1e-6."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dgcnn_b200.synth import make_batch

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
b = make_batch("mutag", seed=324, num_graphs=16, tie_free=True)
torch.manual_seed(324)
x, n = b.x, b.x.size(0)
w = (torch.rand(32, 8) - 0.5)
src, dst = b.edge_index
wts = torch.rand(src.numel())
print("threads", torch.get_num_threads(), "mkldnn", torch.backends.mkldnn.is_available(),
      "fp32 matmul precision", torch.get_float32_matmul_precision())
ops = {
    "matmul x@W.T": (lambda: x @ w.t(), lambda: x.double() @ w.double().t()),
    "linear": (lambda: torch.nn.functional.linear(x, w), lambda: torch.nn.functional.linear(x.double(), w.double())),
    "index_select*w": (lambda: (x @ w.t()).index_select(0, src) * wts[:, None],
                       lambda: (x.double() @ w.double().t()).index_select(0, src) * wts.double()[:, None]),
    "scatter_add": (lambda: torch.zeros(n, 8).scatter_add_(0, dst[:, None].expand(-1, 8), x.index_select(0, src)),
                    lambda: torch.zeros(n, 8, dtype=torch.float64).scatter_add_(0, dst[:, None].expand(-1, 8),
                                                                              x.double().index_select(0, src))),
    "tanh": (lambda: torch.tanh(x), lambda: torch.tanh(x.double())),
    "pow -0.5": (lambda: (wts + 1).pow(-0.5), lambda: (wts.double() + 1).pow(-0.5)),
}
for name, (f32, f64) in ops.items():
    want = f64()
    bad, worst, rows = 0, 0.0, None
    for _ in range(reps):
        got = f32().double()
        err = (got - want).abs()
        m = float(err.max())
        if m > 1e-6:
            bad += 1
            if m > worst:
                worst = m
                r = torch.nonzero(err.reshape(err.size(0), -1).max(1).values > 1e-6).flatten()
                rows = (int(r.min()), int(r.max()), int(r.numel()))
    print(f"{name:16s} runs off by > 1e-6: {bad}/{reps}  worst {worst:.3e}  rows (first, last, count) {rows}")
