"""Pins the CPU oracle (oracle/dgcnn_oracle.py).

The reference ships no tests or golden vectors and its arithmetic lives in PyG,
which is absent here ("parity unpinned", see the oracle's header).  What CAN be
pinned: the README.md:96-104 parameter counts, closed-form dense evaluations of
the published GCN formula in float64 that share no code with the oracle, an
independent pure-Python SortPooling, and the committed golden vectors.
"""
import os

import numpy as np
import pytest
import torch

from oracle import dgcnn_oracle as orc
from dgcnn_b200.synth import make_batch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# README.md:62-104 (dataset, feature dim incl. degree column, classes, parameters)
README_PARAMS = [("MUTAG", 8, 2, 52035), ("PTC", 19, 2, 52387), ("NCI1", 38, 2, 52995),
                 ("PROTEINS", 5, 2, 51939), ("DD", 90, 2, 54659), ("COLLAB", 1, 3, 51940),
                 ("IMDB-B", 1, 2, 51811), ("IMDB-M", 1, 3, 51940)]


@pytest.mark.parametrize("name,f,c,count", README_PARAMS)
def test_parameter_counts_match_readme(name, f, c, count):
    m = orc.OracleModel(f, c)
    assert sum(p.numel() for p in m.parameters()) == count
    keys = set(m.state_dict().keys())
    assert {"conv1.lin.weight", "conv1.bias", "conv4.lin.weight", "conv5.weight",
            "classifier_1.weight", "classifier_2.bias"} <= keys
    assert m.conv1.lin.weight.shape == (32, f) and m.classifier_1.in_features == 352


@pytest.mark.parametrize("k,expected", [(30, 352), (60, 832), (291, 4512), (130, 1952), (512, 8064)])
def test_classifier_width(k, expected):
    assert orc.classifier_in_features(k) == expected


def dense_gcn(x, edge_index, w, b, norm):
    """Independent float64 restatement: build A (multi-edges counted, loops dropped),
    add I, normalise, matmul.  out = N(A+I) x W^T + b."""
    n = x.shape[0]
    a = np.zeros((n, n))
    for s, d in edge_index.T:
        if s != d:
            a[d, s] += 1.0                       # row = target, column = source
    a += np.eye(n)
    deg = a.sum(1)
    if norm == orc.NORM_SYM:
        ah = a / np.sqrt(deg)[:, None] / np.sqrt(deg)[None, :]
    else:
        ah = a / deg[:, None]
    return ah @ (x.astype(np.float64) @ w.astype(np.float64).T) + b.astype(np.float64)


@pytest.mark.parametrize("norm", [orc.NORM_SYM, orc.NORM_RW])
def test_gcn_conv_matches_dense_formula(norm):
    rng = np.random.RandomState(0)
    n, f, c = 23, 7, 5
    ei = rng.randint(0, n, size=(2, 90))         # loops and duplicates included on purpose
    x, w, b = rng.randn(n, f), rng.randn(c, f), rng.randn(c)
    got = orc.gcn_conv(torch.from_numpy(x), torch.from_numpy(ei), torch.from_numpy(w),
                       torch.from_numpy(b), norm).numpy()
    np.testing.assert_allclose(got, dense_gcn(x, ei, w, b, norm), rtol=0, atol=1e-12)


def test_gcn_norm_hand_values():
    """P3 path 0-1-2: deg+1 = (2,3,2); weight(0->1) = 1/sqrt(2*3); loop weights 1/deg."""
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    ei2, w = orc.gcn_norm(ei, 3, torch.float64)
    assert ei2.tolist() == [[0, 1, 1, 2, 0, 1, 2], [1, 0, 2, 1, 0, 1, 2]]
    s6 = 1 / np.sqrt(6)
    np.testing.assert_allclose(w.numpy(), [s6, s6, s6, s6, 0.5, 1 / 3, 0.5], atol=1e-15)


def test_remove_self_loops_keeps_order():
    ei = torch.tensor([[0, 1, 1, 2, 2], [0, 2, 1, 1, 0]])
    out, none = orc.remove_self_loops(ei)
    assert none is None and out.tolist() == [[1, 2, 2], [2, 1, 0]]


def python_sort_pool(x, batch, k, num_graphs):
    """Independent SortPooling: Python's stable sort per graph on (-key)."""
    d = x.shape[1]
    out = np.zeros((num_graphs, k, d), dtype=x.dtype)
    perm = -np.ones((num_graphs, k), dtype=np.int64)
    for g in range(num_graphs):
        nodes = [i for i in range(len(batch)) if batch[i] == g]
        nodes.sort(key=lambda i: -float(x[i, -1]) + 0.0)     # stable; -0.0 + 0.0 == 0.0
        for r, i in enumerate(nodes[:k]):
            out[g, r], perm[g, r] = x[i], i
    return out.reshape(num_graphs, k * d), perm


def test_sort_aggregation_fixture_against_python_sort():
    z = np.load(os.path.join(GOLDEN, "sortpool_cases.npz"))
    x, batch, k, b = z["x"], z["batch"], int(z["k"]), int(z["num_graphs"])
    out, perm = orc.sort_aggregation(torch.from_numpy(x), torch.from_numpy(batch), k, b,
                                     return_perm=True)
    eo, ep = python_sort_pool(x, batch, k, b)
    np.testing.assert_array_equal(perm.numpy(), ep)
    np.testing.assert_array_equal(out.numpy(), eo)
    np.testing.assert_array_equal(z["perm"], ep)
    np.testing.assert_array_equal(z["out"], eo)


@pytest.mark.parametrize("seed", range(5))
def test_sort_aggregation_random_against_python_sort(seed):
    rng = np.random.RandomState(seed)
    sizes = rng.randint(0, 40, size=9)
    sizes[-1] = max(sizes[-1], 1)        # PyG infers B from batch.max(): last graph non-empty
    n = int(sizes.sum())
    x = rng.randn(n, 6).astype(np.float32)
    x[:, -1] = np.round(x[:, -1], 1)     # plenty of exact ties
    batch = np.repeat(np.arange(9), sizes)
    for k in (1, 7, 50):
        out, perm = orc.sort_aggregation(torch.from_numpy(x), torch.from_numpy(batch), k, 9,
                                         return_perm=True)
        eo, ep = python_sort_pool(x, batch, k, 9)
        np.testing.assert_array_equal(perm.numpy(), ep)
        np.testing.assert_array_equal(out.numpy(), eo)


def test_sort_aggregation_nan_first():
    x = torch.tensor([[0.0, 1.0], [1.0, float("nan")], [2.0, 3.0]])
    _, perm = orc.sort_aggregation(x, torch.zeros(3, dtype=torch.long), 3, 1, return_perm=True)
    assert perm.tolist() == [[1, 2, 0]]


def test_pool_gradient_reaches_exactly_min_n_k_rows():
    b = make_batch("mutag", seed=1, num_graphs=5, tie_free=True)
    x = torch.randn(b.num_nodes, 97, requires_grad=True)
    k = 15
    orc.sort_aggregation(x, b.batch, k, 5).sum().backward()
    sizes = (b.ptr[1:] - b.ptr[:-1]).clamp(max=k)
    rows_with_grad = (x.grad.abs().sum(1) > 0).long()
    per_graph = torch.zeros(5, dtype=torch.long).scatter_add_(0, b.batch, rows_with_grad)
    assert per_graph.tolist() == sizes.tolist()


def test_indegree_feature():
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    out = orc.indegree_feature(ei, 3, torch.ones(3, 2))
    np.testing.assert_allclose(out[:, -1].numpy(), [0.5, 1.0, 0.5])
    assert out.shape == (3, 3)


@pytest.mark.parametrize("name", ["hand_sym", "hand_rw", "mutag6_sym", "proteins5_sym"])
def test_golden_vectors(name):
    """The committed vectors are what the oracle produces today, and their float32
    and float64 halves agree to the 1e-5 bar with >10x headroom; x_cat also matches
    the independent dense formula."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    norm, k, b = int(z["norm"]), int(z["k"]), int(z["num_graphs"])
    ws = [torch.from_numpy(z[f"w{i}"]) for i in range(1, 5)]
    bs = [torch.from_numpy(z[f"b{i}"]) for i in range(1, 5)]
    x, ei, bt = (torch.from_numpy(z[n]) for n in ("x", "edge_index", "batch"))
    xcat = orc.graph_conv_stack(x, ei, ws, bs, norm)
    pooled, perm = orc.sort_aggregation(xcat, bt, k, b, return_perm=True)
    np.testing.assert_allclose(xcat.numpy(), z["xcat_f32"], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(perm.numpy(), z["perm_f32"])
    np.testing.assert_allclose(pooled.numpy(), z["pooled_f32"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(z["xcat_f32"], z["xcat_f64"], rtol=0, atol=1e-6)
    for key in ["dx"] + [f"dw{i}" for i in range(1, 5)] + [f"db{i}" for i in range(1, 5)]:
        scale = max(1.0, float(np.abs(z[key + "_f64"]).max()))
        np.testing.assert_allclose(z[key + "_f32"], z[key + "_f64"], rtol=0, atol=2e-5 * scale)
    # independent dense chain in float64
    h = z["x"].astype(np.float64)
    cols = []
    for i in range(4):
        h = np.tanh(dense_gcn(h, z["edge_index"], z[f"w{i+1}"], z[f"b{i+1}"], norm))
        cols.append(h)
    np.testing.assert_allclose(np.concatenate(cols, 1), z["xcat_f64"], rtol=0, atol=1e-12)


def test_oracle_model_forward_backward_runs():
    b = make_batch("mutag", seed=3, num_graphs=8)
    torch.manual_seed(324)
    m = orc.OracleModel(8, 2, k=30).eval()
    out = m(b)
    assert out.shape == (8, 2)
    np.testing.assert_allclose(out.exp().sum(1).detach().numpy(), 1.0, atol=1e-5)
    torch.nn.functional.nll_loss(out, b.y).backward()
    assert all(p.grad is not None for p in m.parameters())
