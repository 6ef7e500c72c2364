"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference itself cannot be imported
here (torch_geometric is absent), so these vectors pin the ORACLE, in float32 and
float64, on small seeded inputs; tests/test_oracle.py additionally checks them
against closed-form dense evaluations that do not share code with the oracle.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import dgcnn_oracle as orc            # noqa: E402
from dgcnn_b200.synth import make_batch           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hand_fixture():
    """graph 0: path 0-1-2.  graph 1: triangle 3-4-5 with a duplicated edge 3->4, a
    self loop on 5 and an isolated node 6.  graph 2: empty.  graph 3: single node 7."""
    und = [(0, 1), (1, 2), (3, 4), (4, 5), (3, 5)]
    src = [a for a, b in und] + [b for a, b in und] + [3, 5]
    dst = [b for a, b in und] + [a for a, b in und] + [4, 5]
    edge_index = np.array([src, dst], dtype=np.int64)
    batch = np.array([0, 0, 0, 1, 1, 1, 1, 3], dtype=np.int64)
    x = (np.arange(8 * 3, dtype=np.float32).reshape(8, 3) % 5 - 2.0) / 4.0
    return x, edge_index, batch, 4


def stack_case(name, x, edge_index, batch, num_graphs, k, seed, norm):
    g = torch.Generator().manual_seed(seed)
    f = x.shape[1]
    dims = [(f, 32), (32, 32), (32, 32), (32, 1)]
    ws = [(torch.rand(co, ci, generator=g) * 2 - 1) * (6.0 / (ci + co)) ** 0.5 for ci, co in dims]
    bs = [(torch.rand(co, generator=g) * 2 - 1) * 0.1 for _, co in dims]
    xt, ei, bt = torch.from_numpy(x), torch.from_numpy(edge_index), torch.from_numpy(batch)
    out = {"x": x, "edge_index": edge_index, "batch": batch,
           "num_graphs": np.int64(num_graphs), "k": np.int64(k), "norm": np.int64(norm)}
    for i, (w, b) in enumerate(zip(ws, bs)):
        out[f"w{i+1}"], out[f"b{i+1}"] = w.numpy(), b.numpy()
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        xs = xt.detach().clone().to(dt).requires_grad_(True)
        wsd = [w.detach().clone().to(dt).requires_grad_(True) for w in ws]
        bsd = [b.detach().clone().to(dt).requires_grad_(True) for b in bs]
        xcat = orc.graph_conv_stack(xs, ei, wsd, bsd, norm)
        pooled, perm = orc.sort_aggregation(xcat, bt, k, num_graphs, return_perm=True)
        # a fixed, non-trivial cotangent so that backward is pinned too
        gg = torch.Generator().manual_seed(seed + 1)
        cot = torch.randn(pooled.shape, generator=gg, dtype=torch.float64).to(dt)
        (pooled * cot).sum().backward()
        out[f"xcat_{tag}"] = xcat.detach().numpy()
        out[f"pooled_{tag}"] = pooled.detach().numpy()
        out[f"perm_{tag}"] = perm.numpy()
        out[f"dx_{tag}"] = xs.grad.numpy()
        for i in range(4):
            out[f"dw{i+1}_{tag}"] = wsd[i].grad.numpy()
            out[f"db{i+1}_{tag}"] = bsd[i].grad.numpy()
        out["cotangent"] = cot.to(torch.float32).numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k_: v.shape for k_, v in out.items() if hasattr(v, "shape") and v.ndim})


def sortpool_cases():
    """keys with ties, signed zeros, all-equal, n<k, n=k, n>k, an empty graph."""
    d, k = 5, 4
    sizes = [2, 4, 7, 0, 6, 1, 9]
    keys = [
        [0.5, 0.5],                                  # n<k, tie
        [0.1, -0.3, 0.7, 0.2],                       # n=k
        [0.0, -0.0, 0.0, 1.0, -1.0, -0.0, 0.0],      # n>k, signed-zero ties
        [],                                          # empty graph
        [0.25] * 6,                                  # all equal
        [-2.0],                                      # single node
        [3.0, 1.0, 3.0, 2.0, 3.0, 1.0, 2.0, 3.0, 0.0],
    ]
    n = sum(sizes)
    rng = np.random.RandomState(7)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[:, -1] = np.concatenate([np.array(kk, dtype=np.float32) for kk in keys])
    batch = np.repeat(np.arange(len(sizes)), sizes).astype(np.int64)
    out, perm = orc.sort_aggregation(torch.from_numpy(x), torch.from_numpy(batch), k, len(sizes),
                                     return_perm=True)
    np.savez_compressed(os.path.join(HERE, "sortpool_cases.npz"), x=x, batch=batch,
                        num_graphs=np.int64(len(sizes)), k=np.int64(k), out=out.numpy(),
                        perm=perm.numpy())
    print("sortpool_cases", out.shape, perm.tolist())


def collate_case():
    """A data set of six hand-sized graphs (path, triangle with a duplicated edge and a self
    loop, an EMPTY graph, a single node, a directed 3-cycle, a star) and a batch of ids with a
    repeat: the batch a resident data set must gather (tests/test_gpu_resident.py), as the
    oracle's Batch.from_data_list + CSR restatement build it."""
    def und(pairs):
        return np.array([[a for a, b in pairs] + [b for a, b in pairs],
                         [b for a, b in pairs] + [a for a, b in pairs]], dtype=np.int64)
    eis = [und([(0, 1), (1, 2)]),
           np.concatenate([und([(0, 1), (1, 2), (0, 2)]), np.array([[0, 2], [1, 2]])], axis=1),
           np.zeros((2, 0), np.int64),
           np.zeros((2, 0), np.int64),
           np.array([[0, 1, 2], [1, 2, 0]], dtype=np.int64),
           und([(0, 1), (0, 2), (0, 3), (0, 4)])]
    sizes = [3, 4, 0, 1, 3, 5]
    rng = np.random.RandomState(3)
    xs = [rng.standard_normal((n, 2)).astype(np.float32) for n in sizes]
    ys = [0, 1, 1, 0, 1, 0]
    ids = np.array([5, 1, 2, 4, 1, 3, 0], dtype=np.int64)
    x, ei, batch, ptr, y = orc.from_data_list([(xs[i], eis[i], ys[i]) for i in ids])
    rowptr, col, rowptr_t, col_t, dis = orc.batch_csr(ei, int(ptr[-1]))
    n_b = [sizes[i] for i in ids]
    gorder = sorted(range(len(ids)), key=lambda q: (-n_b[q], q))
    out = {"ids": ids, "sizes": np.array(sizes), "ys": np.array(ys), "x": x.numpy(),
           "edge_index": ei.numpy(), "batch": batch.numpy(), "ptr": ptr.numpy(), "y": y.numpy(),
           "rowptr": rowptr.numpy(), "col": col.numpy(), "rowptr_t": rowptr_t.numpy(), "col_t": col_t.numpy(),
           "dis": dis.numpy(), "gorder": np.array(gorder)}
    for i in range(len(sizes)):
        out[f"g{i}_x"], out[f"g{i}_edge_index"] = xs[i], eis[i]
    np.savez_compressed(os.path.join(HERE, "collate_case.npz"), **out)
    print("collate_case", {k_: v.shape for k_, v in out.items() if not k_.startswith("g")})


if __name__ == "__main__":
    collate_case()
    x, ei, b, nb = hand_fixture()
    stack_case("hand_sym", x, ei, b, nb, k=3, seed=11, norm=orc.NORM_SYM)
    stack_case("hand_rw", x, ei, b, nb, k=3, seed=11, norm=orc.NORM_RW)
    mb = make_batch("mutag", seed=324, num_graphs=6, tie_free=True)
    stack_case("mutag6_sym", mb.x.numpy(), mb.edge_index.numpy(), mb.batch.numpy(), 6, k=30,
               seed=324, norm=orc.NORM_SYM)
    pb = make_batch("proteins", seed=5, num_graphs=5, tie_free=True)
    stack_case("proteins5_sym", pb.x.numpy(), pb.edge_index.numpy(), pb.batch.numpy(), 5, k=20,
               seed=5, norm=orc.NORM_SYM)
    sortpool_cases()
