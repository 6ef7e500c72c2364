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


# ----------------------------------------------------------------------------------------
# property tests: hypothesis draws small block-diagonal multigraphs
# ----------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st


@st.composite
def graph_lists(draw, max_graphs=5, max_nodes=9, features=3):
    """A list of small directed multigraphs (loops and duplicates allowed, empty graphs too),
    tie-free float64 features."""
    seed = draw(st.integers(0, 2 ** 31 - 1))
    sizes = draw(st.lists(st.integers(0, max_nodes), min_size=1, max_size=max_graphs))
    if sum(sizes) == 0:
        sizes[0] = 1
    rng = np.random.RandomState(seed)
    graphs = []
    for n in sizes:
        m = int(rng.randint(0, 3 * n + 1)) if n else 0
        ei = np.stack([rng.randint(0, max(n, 1), m), rng.randint(0, max(n, 1), m)]).astype(np.int64)
        graphs.append((rng.standard_normal((n, features)), ei, int(rng.randint(0, 2))))
    return graphs


def oracle_weights(features, seed=0):
    g = torch.Generator().manual_seed(seed)
    dims = [(features, 32), (32, 32), (32, 32), (32, 1)]
    ws = [(torch.rand(co, ci, generator=g, dtype=torch.float64) * 2 - 1) * (6.0 / (ci + co)) ** 0.5 for ci, co in dims]
    bs = [(torch.rand(co, generator=g, dtype=torch.float64) * 2 - 1) * 0.1 for _, co in dims]
    return ws, bs


@settings(max_examples=25, deadline=None)
@given(graph_lists(), st.sampled_from([orc.NORM_SYM, orc.NORM_RW]))
def test_property_stack_matches_dense_matmul_in_float64(graphs, norm):
    """4 x tanh(N(A+I) X W^T + b) on the batched multigraph == the dense closed form."""
    x, ei, batch, ptr, y = orc.from_data_list(graphs)
    ws, bs = oracle_weights(x.size(1))
    got = orc.graph_conv_stack(x, ei, ws, bs, norm).numpy()
    h, outs = x.numpy(), []
    for w, b in zip(ws, bs):
        h = np.tanh(dense_gcn(h, ei.numpy(), w.numpy(), b.numpy(), norm))
        outs.append(h)
    np.testing.assert_allclose(got, np.concatenate(outs, 1), rtol=0, atol=1e-12)


@settings(max_examples=25, deadline=None)
@given(graph_lists(), st.integers(1, 12))
def test_property_batch_of_one_equals_batched(graphs, k):
    """Graphs never interact (model.py:30-35): the batched hot path == every graph on its own."""
    x, ei, batch, ptr, y = orc.from_data_list(graphs)
    ws, bs = oracle_weights(x.size(1))
    xcat = orc.graph_conv_stack(x, ei, ws, bs)
    pooled, perm = orc.sort_aggregation(xcat, batch, k, len(graphs), return_perm=True)
    for g, (gx, gei, _) in enumerate(graphs):
        lo, hi = int(ptr[g]), int(ptr[g + 1])
        if hi == lo:
            assert pooled[g].abs().sum() == 0 and (perm[g] == -1).all()      # empty graph: all padding
            continue
        own = orc.graph_conv_stack(torch.as_tensor(gx), torch.as_tensor(gei), ws, bs)
        np.testing.assert_allclose(xcat[lo:hi].numpy(), own.numpy(), rtol=0, atol=1e-13)
        own_pool, own_perm = orc.sort_aggregation(own, torch.zeros(hi - lo, dtype=torch.long), k, 1, return_perm=True)
        np.testing.assert_allclose(pooled[g].numpy(), own_pool[0].numpy(), rtol=0, atol=1e-13)
        # ranks are only defined up to rounding when two keys agree to ~1e-13 (the two runs sum in
        # different orders); the pooled rows were compared above
        ks = torch.sort(own[:, -1], descending=True).values
        if hi - lo < 2 or float((ks[:-1] - ks[1:]).min()) > 1e-9:
            assert torch.equal(torch.where(perm[g] >= 0, perm[g] - lo, perm[g]), own_perm[0])


@settings(max_examples=25, deadline=None)
@given(graph_lists(), st.integers(1, 12), st.integers(0, 2 ** 31 - 1))
def test_property_relabelling_nodes_keeps_the_pooled_output(graphs, k, seed):
    """Permutation equivariance: relabel the nodes of every graph => same pooled rows (keys are
    tie-free with probability 1, so the tie rule does not enter), x_cat rows permuted."""
    rng = np.random.RandomState(seed)
    relabelled, perms = [], []
    for gx, gei, gy in graphs:
        p = rng.permutation(gx.shape[0])                     # new id of old node i is p[i]
        inv = np.argsort(p)
        relabelled.append((gx[inv], p[gei] if gei.size else gei, gy))
        perms.append(p)
    ws, bs = oracle_weights(graphs[0][0].shape[1])
    x, ei, batch, ptr, _ = orc.from_data_list(graphs)
    x2, ei2, batch2, _, _ = orc.from_data_list(relabelled)
    xcat, xcat2 = orc.graph_conv_stack(x, ei, ws, bs), orc.graph_conv_stack(x2, ei2, ws, bs)
    for g, p in enumerate(perms):
        lo = int(ptr[g])
        np.testing.assert_allclose(xcat2[lo + p].numpy(), xcat[lo:lo + len(p)].numpy(), rtol=0, atol=1e-13)
    keys = xcat[:, -1].numpy()
    for g in range(len(graphs)):                             # the property needs well-separated keys
        seg = np.sort(keys[int(ptr[g]):int(ptr[g + 1])])
        if seg.size > 1 and np.min(np.diff(seg)) < 1e-9:
            return
    pooled = orc.sort_aggregation(xcat, batch, k, len(graphs))
    pooled2 = orc.sort_aggregation(xcat2, batch2, k, len(graphs))
    np.testing.assert_allclose(pooled2.numpy(), pooled.numpy(), rtol=0, atol=1e-12)


@settings(max_examples=40, deadline=None)
@given(graph_lists(max_graphs=6, max_nodes=12))
def test_property_batch_csr_is_the_edge_multiset(graphs):
    """from_data_list + batch_csr: both CSRs hold exactly the non-loop edges (multi-edges kept),
    rows ascending, and dis = (1 + in-degree)^-1/2."""
    x, ei, batch, ptr, _ = orc.from_data_list(graphs)
    n = int(ptr[-1])
    rowptr, col, rowptr_t, col_t, dis = orc.batch_csr(ei, n)
    want = sorted((int(s), int(d)) for s, d in ei.t().tolist() if s != d)
    by_target = sorted((int(col[e]), i) for i in range(n) for e in range(int(rowptr[i]), int(rowptr[i + 1])))
    by_source = sorted((i, int(col_t[e])) for i in range(n) for e in range(int(rowptr_t[i]), int(rowptr_t[i + 1])))
    assert by_target == want and by_source == want
    for i in range(n):
        row = col[int(rowptr[i]):int(rowptr[i + 1])].tolist()
        assert row == sorted(row)
        assert all(int(batch[j]) == int(batch[i]) for j in row)       # edges never leave their graph
    indeg = np.bincount([d for _, d in want], minlength=n)
    np.testing.assert_allclose(dis.numpy(), (1.0 + indeg) ** -0.5, rtol=1e-6)
