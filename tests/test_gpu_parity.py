"""GPU parity tests: every CUDA entry point of include/dgcnn_b200.h, called through
the ctypes binding (dgcnn_b200.ops), against the CPU oracle on the same seeded
inputs, the committed golden vectors, and size-independent properties.

Tolerances (north_star): fp32 features within 1e-5 absolute of the oracle (outputs
are tanh-bounded); permutation indices and CSR arrays bit-exact.
"""
import os
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import dgcnn_b200 as dg
from dgcnn_b200 import ops
from dgcnn_b200.synth import CONFIGS, collate, make_batch, make_graphs
from oracle import dgcnn_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ATOL = 1e-5
DEV = "cuda:0"
# the dense tail is stock torch: keep cuDNN/cuBLAS in true fp32 so that the end-to-end
# comparison against the float64 oracle measures OUR kernels, not TF32 rounding
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


# ------------------------------------------------------------------ helpers
def ref_csr(edge_index: np.ndarray, n: int):
    src, dst = edge_index
    keep = src != dst
    src, dst = src[keep], dst[keep]
    o = np.lexsort((src, dst))
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=n))]).astype(np.int32)
    ot = np.lexsort((dst, src))
    rowptr_t = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=n))]).astype(np.int32)
    indeg = np.bincount(dst, minlength=n)
    dis = (1.0 / np.sqrt((indeg + 1).astype(np.float32))).astype(np.float32)
    return rowptr, src[o].astype(np.int32), rowptr_t, dst[ot].astype(np.int32), dis


def random_multigraph(rng, sizes, avg_deg=4.0, loops=True):
    """Block-diagonal directed multigraph: duplicates, self loops, isolated nodes,
    edges in random order."""
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    srcs, dsts = [], []
    for g, n in enumerate(sizes):
        if n == 0:
            continue
        m = int(avg_deg * n)
        s = rng.randint(0, n, size=m) + ptr[g]
        d = rng.randint(0, n, size=m) + ptr[g]
        if not loops:
            ok = s != d
            s, d = s[ok], d[ok]
        srcs.append(s)
        dsts.append(d)
    src = np.concatenate(srcs) if srcs else np.zeros(0, np.int64)
    dst = np.concatenate(dsts) if dsts else np.zeros(0, np.int64)
    p = rng.permutation(src.size)
    ei = np.stack([src[p], dst[p]]).astype(np.int64)
    batch = np.repeat(np.arange(len(sizes)), sizes).astype(np.int64)
    return ei, batch, int(ptr[-1])


def gpu_graph(ei, batch, n, b, **kw):
    return ops.build_graph(torch.from_numpy(ei).to(DEV), torch.from_numpy(batch).to(DEV), n, b, **kw)


def assert_perm_matches(perm_gpu, perm_ref, keys64, gptr, k, tol=1e-5):
    """perm must equal the oracle's wherever the oracle's neighbouring sorted keys are
    further apart than the fp32 tolerance (near-ties may legitimately swap)."""
    perm_gpu, perm_ref = np.asarray(perm_gpu), np.asarray(perm_ref)
    assert ((perm_gpu < 0) == (perm_ref < 0)).all()
    bad = 0
    for g in range(perm_ref.shape[0]):
        n = gptr[g + 1] - gptr[g]
        if n == 0:
            continue
        ks = np.sort(keys64[gptr[g]:gptr[g + 1]])[::-1]
        for r in range(min(n, k)):
            if perm_gpu[g, r] == perm_ref[g, r]:
                continue
            near = (r > 0 and abs(ks[r] - ks[r - 1]) < tol) or (r + 1 < n and abs(ks[r] - ks[r + 1]) < tol)
            assert near, f"graph {g} rank {r}: {perm_gpu[g, r]} vs {perm_ref[g, r]} without a near-tie"
            bad += 1
    return bad


# ------------------------------------------------------------------ K0
@pytest.mark.parametrize("sizes", [[5, 0, 7, 1, 0], [2048], [2047, 1], [4096, 100], [1], [0, 0, 3],
                                   list(range(0, 60)), [30000, 20000, 15000]])
def test_build_graph_matches_reference_csr(sizes):
    rng = np.random.RandomState(len(sizes) + sum(sizes))
    ei, batch, n = random_multigraph(rng, sizes, avg_deg=5.0)
    b = len(sizes)
    g = gpu_graph(ei, batch, n, b)
    g.check()
    rowptr, col, rowptr_t, col_t, dis = ref_csr(ei, n)
    e = int(rowptr[-1])
    np.testing.assert_array_equal(g.rowptr.cpu().numpy(), rowptr)
    np.testing.assert_array_equal(g.col.cpu().numpy()[:e], col)
    np.testing.assert_array_equal(g.rowptr_t.cpu().numpy(), rowptr_t)
    np.testing.assert_array_equal(g.col_t.cpu().numpy()[:e], col_t)
    np.testing.assert_allclose(g.dis.cpu().numpy(), dis, rtol=2e-7, atol=0)
    np.testing.assert_array_equal(g.gptr.cpu().numpy(), np.concatenate([[0], np.cumsum(sizes)]))
    # gorder: graph ids by descending size, ties by id
    want = sorted(range(b), key=lambda i: (-sizes[i], i))
    assert g.gorder.cpu().tolist() == want


def test_build_graph_fast_path_equals_generic_path():
    """A sorted symmetric edge list takes the streaming fast path; shuffling the same
    edges forces the generic count/scan/fill/sort path.  Both must give the same CSR."""
    for name in ("proteins", "collab"):
        b = make_batch(name, num_graphs=40)
        ei = b.edge_index
        perm = torch.randperm(ei.size(1), generator=torch.Generator().manual_seed(0))
        g1 = ops.build_graph(ei.to(DEV), b.batch.to(DEV), b.num_nodes, b.num_graphs)
        g2 = ops.build_graph(ei[:, perm].contiguous().to(DEV), b.batch.to(DEV), b.num_nodes, b.num_graphs)
        e = ei.size(1)
        for a, c in ((g1.rowptr, g2.rowptr), (g1.col[:e], g2.col[:e]), (g1.rowptr_t, g2.rowptr_t),
                     (g1.col_t[:e], g2.col_t[:e]), (g1.dis, g2.dis), (g1.gptr, g2.gptr)):
            assert torch.equal(a, c)
        # one missing reverse edge (asymmetric) must fall back and still be right
        ei3 = ei[:, 1:].contiguous()
        g3 = ops.build_graph(ei3.to(DEV), b.batch.to(DEV), b.num_nodes, b.num_graphs)
        rowptr, col, rowptr_t, col_t, dis = ref_csr(ei3.numpy(), b.num_nodes)
        np.testing.assert_array_equal(g3.rowptr.cpu().numpy(), rowptr)
        np.testing.assert_array_equal(g3.col.cpu().numpy()[:rowptr[-1]], col)
        np.testing.assert_array_equal(g3.rowptr_t.cpu().numpy(), rowptr_t)
        np.testing.assert_array_equal(g3.col_t.cpu().numpy()[:rowptr[-1]], col_t)
        np.testing.assert_allclose(g3.dis.cpu().numpy(), dis, rtol=2e-7)


def test_build_graph_int32_indices_equal_int64():
    """dgcnn_build_graph_i32 (compact host batches) gives the same graph as the int64 entry
    point, on the fast path (sorted symmetric batch) and on the generic one (shuffled edges)."""
    b = make_batch("collab", num_graphs=24)
    mx = int((b.ptr[1:] - b.ptr[:-1]).max())
    perm = torch.randperm(b.edge_index.size(1), generator=torch.Generator().manual_seed(1))
    for ei in (b.edge_index, b.edge_index[:, perm].contiguous()):
        g64 = ops.build_graph(ei.to(DEV), b.batch.to(DEV), b.num_nodes, b.num_graphs, max_nodes=mx)
        g32 = ops.build_graph(ei.to(torch.int32).to(DEV), b.batch.to(torch.int32).to(DEV), b.num_nodes,
                              b.num_graphs, max_nodes=mx)
        assert int(g64.status.item()) == int(g32.status.item())
        e = ei.size(1)
        for name in ("rowptr", "rowptr_t", "dis", "gptr", "gorder", "bitmap", "bmoff", "gflags", "fgoff",
                     "gdesc"):
            assert torch.equal(getattr(g64, name), getattr(g32, name)), name
        used = int(g64.fgoff[-1])                   # the buffer is sized by an upper bound
        assert torch.equal(g64.fragmap[:used], g32.fragmap[:used])
        assert torch.equal(g64.col[:e], g32.col[:e]) and torch.equal(g64.col_t[:e], g32.col_t[:e])
    # and the whole model runs on a compact batch
    cfg = CONFIGS["collab"]
    torch.manual_seed(0)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).eval()
    d64, d32 = b.to(DEV), b.compact().to(DEV)
    d64.max_nodes = d32.max_nodes = mx
    with torch.no_grad():
        assert torch.equal(model(d64), model(d32))


def test_fingerprint_symmetry_check_rejects_asymmetric_sorted_lists(monkeypatch):
    """K0's fast path proves symmetry with multiset fingerprints.  Sorted, duplicate-free lists
    that are NOT symmetric -- including ones that keep every in/out degree intact -- must take
    the generic path and give the reference CSR; the exact binary-search check
    (DGCNN_EXACT_SYMMETRY=1) must agree on both verdicts."""
    b = make_batch("proteins", num_graphs=30)
    ei = b.edge_index.numpy()
    n = b.num_nodes
    # a directed 3-cycle a->b->c->a inside graph 0 replaces nothing: add it and keep the order
    rng = np.random.RandomState(3)
    cases = {"symmetric": ei}
    extra = np.array([[0, 1, 2], [1, 2, 0]])            # nodes 0,1,2 belong to graph 0
    have = set(map(tuple, ei.T.tolist()))
    add = np.array([e for e in extra.T.tolist() if tuple(e) not in have] or [[0, 3]]).T
    asym = np.concatenate([ei, add], 1)
    asym = asym[:, np.lexsort((asym[1], asym[0]))]
    cases["one-way edges"] = asym
    # reverse the direction of one existing edge pair: (u,v),(v,u) -> keep only (u,v) twice is a
    # duplicate, so instead drop (v,u) and add (v,w) for another neighbour w of u: degrees change
    # little, symmetry breaks
    drop = int(rng.randint(ei.shape[1]))
    cases["missing reverse edge"] = np.delete(ei, drop, axis=1)
    for exact in (False, True):
        monkeypatch.setattr(ops, "EXACT_SYMMETRY_CHECK", exact)
        for name, e in cases.items():
            e = np.ascontiguousarray(e)
            g = gpu_graph(e, b.batch.numpy(), n, b.num_graphs)
            generic = bool(int(g.status.item()) & ops.GRAPH_GENERIC)
            assert generic == (name != "symmetric"), (name, exact)
            rowptr, col, rowptr_t, col_t, dis = ref_csr(e, n)
            m = int(rowptr[-1])
            np.testing.assert_array_equal(g.rowptr.cpu().numpy(), rowptr)
            np.testing.assert_array_equal(g.col.cpu().numpy()[:m], col)
            np.testing.assert_array_equal(g.rowptr_t.cpu().numpy(), rowptr_t)
            np.testing.assert_array_equal(g.col_t.cpu().numpy()[:m], col_t)
            np.testing.assert_allclose(g.dis.cpu().numpy(), dis, rtol=2e-7)


def test_bitmaps_match_the_adjacency():
    """K0b: bit (r,c) of graph g  <=>  edge c->r or r == c; duplicates flagged per graph;
    the transposed bitmap is only built when the input is not a sorted symmetric list."""
    rng = np.random.RandomState(5)
    sizes = [5, 0, 33, 64, 1, 100, 17]
    ei, batch, n = random_multigraph(rng, sizes, avg_deg=3.0)
    b = len(sizes)
    mx = max(sizes)
    g = gpu_graph(ei, batch, n, b, max_nodes=mx)
    assert int(g.status.item()) & ops.GRAPH_GENERIC
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    for bitmap, bmoff, gflags, rows, cols in ((g.bitmap, g.bmoff, g.gflags, ei[1], ei[0]),
                                              (g.bitmap_t, g.bmoff_t, g.gflags_t, ei[0], ei[1])):
        bm, off, fl = bitmap.cpu().numpy().view(np.uint32), bmoff.cpu().numpy(), gflags.cpu().numpy()
        for gi, ng in enumerate(sizes):
            npad = (ng + 15) // 16 * 16
            wpr = (npad + 31) // 32
            assert off[gi + 1] - off[gi] == npad * wpr
            if ng == 0:
                continue
            dense = np.zeros((npad, wpr * 32), dtype=bool)
            sel = (batch[rows] == gi) & (rows != cols)
            dense[rows[sel] - ptr[gi], cols[sel] - ptr[gi]] = True
            dense[np.arange(ng), np.arange(ng)] = True
            words = bm[off[gi]:off[gi + 1]].reshape(npad, wpr)
            got = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(npad, -1)
            np.testing.assert_array_equal(got, dense)
            pairs = np.stack([rows[sel], cols[sel]], 1)
            has_dup = len(np.unique(pairs, axis=0)) != len(pairs)
            assert bool(fl[gi] & 1) == has_dup
            if bitmap is g.bitmap:
                # fragment-major copy (tensor-core kernels): word (mt, grp, lane) carries the
                # lane's m16k16 A-fragment bits of blocks kt = 4 grp + q
                fm = g.fragmap.cpu().numpy().view(np.uint32)
                fo = g.fgoff.cpu().numpy()
                tiles = npad // 16
                groups = (tiles + 3) // 4
                assert fo[gi + 1] - fo[gi] == tiles * groups * 32
                want = np.zeros((tiles, groups, 32), dtype=np.uint32)
                full = np.zeros((npad, groups * 64), dtype=bool)
                full[:, :dense.shape[1]] = dense[:, :groups * 64]
                for lane in range(32):
                    gq, t = lane >> 2, lane & 3
                    for q in range(4):
                        for i in range(4):
                            r = np.arange(tiles) * 16 + gq + 8 * (i & 1)
                            for grp in range(groups):
                                c = grp * 64 + q * 16 + 2 * t + 8 * (i >> 1)
                                m = 4 * q + i
                                want[:, grp, lane] |= (full[r, c].astype(np.uint32) << m) | \
                                                      (full[r, c + 1].astype(np.uint32) << (16 + m))
                np.testing.assert_array_equal(fm[fo[gi]:fo[gi + 1]].reshape(tiles, groups, 32), want)
    # sorted symmetric input: proven symmetric on the device, transposed bitmap left empty
    bt = make_batch("proteins", num_graphs=20)
    g2 = ops.build_graph(bt.edge_index.to(DEV), bt.batch.to(DEV), bt.num_nodes, 20,
                         max_nodes=int((bt.ptr[1:] - bt.ptr[:-1]).max()))
    assert not int(g2.status.item()) & ops.GRAPH_GENERIC
    assert int(g2.bitmap_t.abs().sum()) == 0 and int(g2.bitmap.abs().sum()) > 0


def test_build_graph_heavy_rows_and_no_edges():
    # star: one row of degree 5000 (rank sort beyond one warp chunk), plus an edgeless graph
    n = 5001
    src = np.arange(1, n, dtype=np.int64)
    ei = np.stack([np.concatenate([src, np.zeros(n - 1, np.int64)]),
                   np.concatenate([np.zeros(n - 1, np.int64), src])])
    ei = ei[:, np.random.RandomState(0).permutation(ei.shape[1])]
    batch = np.zeros(n, np.int64)
    g = gpu_graph(ei, batch, n, 1)
    rowptr, col, rowptr_t, col_t, dis = ref_csr(ei, n)
    np.testing.assert_array_equal(g.rowptr.cpu().numpy(), rowptr)
    np.testing.assert_array_equal(g.col.cpu().numpy()[:rowptr[-1]], col)
    np.testing.assert_array_equal(g.col_t.cpu().numpy()[:rowptr[-1]], col_t)
    g2 = gpu_graph(np.zeros((2, 0), np.int64), np.zeros(4, np.int64), 4, 1, transpose=False)
    assert g2.rowptr.cpu().tolist() == [0] * 5 and g2.dis.cpu().tolist() == [1.0] * 4
    assert g2.rowptr_t is None


def test_build_graph_flags_bad_input():
    ei = torch.tensor([[0, 9], [1, 0]], device=DEV)
    g = ops.build_graph(ei, torch.tensor([0, 0, 0], device=DEV), 3, 1)
    with pytest.raises(ValueError, match="edge_index"):
        g.check()
    g = ops.build_graph(torch.tensor([[0], [1]], device=DEV), torch.tensor([1, 0, 0], device=DEV), 3, 2)
    with pytest.raises(ValueError, match="batch"):
        g.check()


def test_synthetic_batches_build():
    for name in ("mutag", "proteins", "dd", "collab"):
        b = make_batch(name)
        g = ops.build_graph(b.edge_index.to(DEV), b.batch.to(DEV), b.num_nodes, b.num_graphs)
        g.check()
        rowptr, col, rowptr_t, col_t, dis = ref_csr(b.edge_index.numpy(), b.num_nodes)
        np.testing.assert_array_equal(g.rowptr.cpu().numpy(), rowptr)
        np.testing.assert_array_equal(g.col.cpu().numpy()[:rowptr[-1]], col)
        # symmetric inputs: the transpose equals the matrix
        np.testing.assert_array_equal(g.rowptr_t.cpu().numpy(), rowptr)
        np.testing.assert_array_equal(g.col_t.cpu().numpy()[:rowptr[-1]], col)
        np.testing.assert_array_equal(g.gptr.cpu().numpy(), b.ptr.numpy().astype(np.int32))


# ------------------------------------------------------------------ K1 / K3
DIMS = [(1, 32), (2, 5), (5, 32), (8, 32), (3, 7), (19, 32), (32, 32), (32, 1), (38, 32), (90, 32),
        (64, 64), (128, 128), (100, 3), (33, 40)]


def conv_case(seed, cin, cout, sizes=(17, 1, 0, 40, 9), avg_deg=3.0):
    rng = np.random.RandomState(seed)
    ei, batch, n = random_multigraph(rng, list(sizes), avg_deg)
    x = rng.randn(n, cin).astype(np.float32)
    w = (rng.randn(cout, cin) / np.sqrt(cin)).astype(np.float32)
    b = (rng.randn(cout) * 0.1).astype(np.float32)
    return ei, batch, n, x, w, b


@pytest.mark.parametrize("cin,cout", DIMS)
@pytest.mark.parametrize("norm", [0, 1])
@pytest.mark.parametrize("act", [0, 1])
def test_graph_conv_forward(cin, cout, norm, act):
    ei, batch, n, x, w, b = conv_case(cin * 131 + cout, cin, cout)
    g = gpu_graph(ei, batch, n, 5)
    # x and out live inside wider buffers (column slices), like the [N,97] concat buffer
    xbuf = torch.zeros(n, cin + 3, device=DEV)
    xbuf[:, 2:2 + cin] = torch.from_numpy(x).to(DEV)
    obuf = torch.full((n, cout + 5), 7.0, device=DEV)
    ops.graph_conv_fwd(xbuf[:, 2:2 + cin], g.rowptr, g.col, g.dis, torch.from_numpy(w).to(DEV),
                       torch.from_numpy(b).to(DEV), norm, act, obuf[:, 1:1 + cout])
    ref = orc.gcn_conv(torch.from_numpy(x), torch.from_numpy(ei), torch.from_numpy(w),
                       torch.from_numpy(b), norm)
    ref64 = orc.gcn_conv(torch.from_numpy(x).double(), torch.from_numpy(ei),
                         torch.from_numpy(w).double(), torch.from_numpy(b).double(), norm)
    if act:
        ref, ref64 = torch.tanh(ref), torch.tanh(ref64)
    got = obuf[:, 1:1 + cout].cpu()
    scale = max(1.0, float(ref64.abs().max()))
    err64 = (got.double() - ref64).abs().max().item()
    err32 = (got - ref).abs().max().item()
    orc32 = (ref.double() - ref64).abs().max().item()
    assert err64 <= ATOL * scale, f"vs fp64 oracle {err64:.3e}; vs fp32 oracle {err32:.3e}; fp32 oracle vs fp64 {orc32:.3e}"
    assert err32 <= ATOL * scale + orc32, f"vs fp32 oracle {err32:.3e} (fp32 oracle itself is {orc32:.3e} from fp64)"
    # nothing outside the slice was touched
    assert (obuf[:, 0] == 7.0).all() and (obuf[:, 1 + cout:] == 7.0).all()


def test_graph_conv_forward_without_bias_and_via_torch_ops():
    ei, batch, n, x, w, b = conv_case(5, 32, 32)
    g = gpu_graph(ei, batch, n, 5)
    out = torch.empty(n, 32, device=DEV)
    torch.ops.dgcnn_b200.graph_conv_fwd(torch.from_numpy(x).to(DEV), g.rowptr, g.col, g.dis,
                                        torch.from_numpy(w).to(DEV), None, 0, 0, out)
    ref = orc.gcn_conv(torch.from_numpy(x).double(), torch.from_numpy(ei), torch.from_numpy(w).double(), None)
    assert (out.cpu().double() - ref).abs().max().item() <= ATOL


@pytest.mark.parametrize("cin,cout", DIMS)
@pytest.mark.parametrize("norm", [0, 1])
def test_graph_conv_backward(cin, cout, norm):
    ei, batch, n, x, w, b = conv_case(cin * 17 + cout + 1, cin, cout)
    g = gpu_graph(ei, batch, n, 5)
    rng = np.random.RandomState(1)
    dy = rng.randn(n, cout).astype(np.float32)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    bt = torch.from_numpy(b).double().requires_grad_(True)
    y64 = torch.tanh(orc.gcn_conv(xt, torch.from_numpy(ei), wt, bt, norm))
    y64.backward(torch.from_numpy(dy).double())

    xd, wd = torch.from_numpy(x).to(DEV), torch.from_numpy(w).to(DEV)
    yd = torch.empty(n, cout, device=DEV)
    ops.graph_conv_fwd(xd, g.rowptr, g.col, g.dis, wd, torch.from_numpy(b).to(DEV), norm, 1, yd)
    dx0 = torch.from_numpy(rng.randn(n, cin).astype(np.float32)).to(DEV)
    dx = dx0.clone()
    dw, db = ops.graph_conv_bwd(torch.from_numpy(dy).to(DEV), yd, xd, g.rowptr_t, g.col_t, g.dis, wd,
                                norm, 1, dx, True)
    for got, ref in ((dx - dx0, xt.grad), (dw, wt.grad), (db, bt.grad)):
        scale = max(1.0, float(ref.abs().max()))
        assert (got.cpu().double() - ref).abs().max().item() <= 2e-5 * scale
    # overwrite mode, no db, no dx
    dx2 = torch.full((n, cin), 3.0, device=DEV)
    dw2, none = ops.graph_conv_bwd(torch.from_numpy(dy).to(DEV), yd, xd, g.rowptr_t, g.col_t, g.dis, wd,
                                   norm, 1, dx2, False, need_db=False)
    assert none is None
    assert (dx2.cpu().double() - xt.grad).abs().max().item() <= 2e-5 * max(1.0, float(xt.grad.abs().max()))
    dw3, _ = ops.graph_conv_bwd(torch.from_numpy(dy).to(DEV), yd, xd, g.rowptr_t, g.col_t, g.dis, wd,
                                norm, 1, None, False)
    assert (dw3 - dw2).abs().max().item() <= 1e-4 * max(1.0, float(dw2.abs().max()))


def test_gcnconv_module_autograd_matches_oracle():
    ei, batch, n, x, w, b = conv_case(77, 19, 32, sizes=(25, 30))
    torch.manual_seed(0)
    conv = dg.GCNConv(19, 32).to(DEV)
    assert set(conv.state_dict()) == {"lin.weight", "bias"}
    with torch.no_grad():
        conv.bias.uniform_(-0.1, 0.1)
    xd = torch.from_numpy(x).to(DEV).requires_grad_(True)
    out = conv(xd, torch.from_numpy(ei).to(DEV))
    out.square().sum().backward()
    xr = torch.from_numpy(x).double().requires_grad_(True)           # float64 oracle: host-independent
    wr = conv.lin.weight.detach().cpu().double().requires_grad_(True)
    br = conv.bias.detach().cpu().double().requires_grad_(True)
    ref = orc.gcn_conv(xr, torch.from_numpy(ei), wr, br)
    ref.square().sum().backward()
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= ATOL * max(1, ref.abs().max().item())
    for got, want in ((xd.grad, xr.grad), (conv.lin.weight.grad, wr.grad), (conv.bias.grad, br.grad)):
        assert (got.cpu().double() - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())


# ------------------------------------------------------------------ K2 / K4
def test_sort_pool_golden_cases_bit_exact():
    z = np.load(os.path.join(GOLDEN, "sortpool_cases.npz"))
    x = torch.from_numpy(z["x"]).to(DEV)
    gptr = ops.graph_ptr(torch.from_numpy(z["batch"]).to(DEV), int(z["num_graphs"]))
    out, perm = ops.sort_pool_fwd(x, gptr, int(z["k"]))
    np.testing.assert_array_equal(perm.cpu().numpy(), z["perm"])
    np.testing.assert_array_equal(out.cpu().numpy(), z["out"])


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("k", [1, 7, 30, 200])
@pytest.mark.parametrize("hint", [0, 16, 5000])
def test_sort_pool_random_bit_exact(seed, k, hint):
    """ragged sizes with empty graphs and many exact ties; hint=16 forces graphs larger
    than the shared buffer through the global-memory sort path."""
    rng = np.random.RandomState(seed)
    sizes = list(rng.randint(0, 90, size=12)) + [1, 0, 300, 1025]
    n = int(np.sum(sizes))
    d = 97
    xbuf = rng.randn(n, d + 3).astype(np.float32)
    xbuf[:, d - 1] = np.round(xbuf[:, d - 1], 1)
    xbuf[::17, d - 1] = -0.0
    batch = np.repeat(np.arange(len(sizes)), sizes).astype(np.int64)
    xg = torch.from_numpy(xbuf).to(DEV)[:, :d]
    gptr = ops.graph_ptr(torch.from_numpy(batch).to(DEV), len(sizes))
    out, perm = ops.sort_pool_fwd(xg, gptr, k, hint)
    ro, rp = orc.sort_aggregation(torch.from_numpy(xbuf[:, :d].copy()), torch.from_numpy(batch), k,
                                  len(sizes), return_perm=True)
    np.testing.assert_array_equal(perm.cpu().numpy(), rp.numpy())
    np.testing.assert_array_equal(out.cpu().numpy(), ro.numpy())


def test_sort_pool_giant_graph_and_nan():
    rng = np.random.RandomState(3)
    sizes = [20000, 5, 5748]
    n, d, k = sum(sizes), 4, 291
    x = rng.randn(n, d).astype(np.float32)
    x[123, -1] = np.inf
    x[20003, -1] = 1e30      # (not -inf: PyG's fill = x.min()-1 would collide and zero the entry)
    batch = np.repeat(np.arange(3), sizes).astype(np.int64)
    gptr = ops.graph_ptr(torch.from_numpy(batch).to(DEV), 3)
    for hint in (0, 20000, 5748):
        out, perm = ops.sort_pool_fwd(torch.from_numpy(x).to(DEV), gptr, k, hint)
        ro, rp = orc.sort_aggregation(torch.from_numpy(x), torch.from_numpy(batch), k, 3, return_perm=True)
        np.testing.assert_array_equal(perm.cpu().numpy(), rp.numpy())
        assert perm[0, 0].item() == 123 and perm[1, 0].item() == 20003
        np.testing.assert_array_equal(out.cpu().numpy(), ro.numpy())
    # NaN keys sort first.  Single graph only: in PyG a NaN anywhere turns the pad value
    # x.min()-1 into NaN, so padded batches are garbage in the reference itself.
    x1 = rng.randn(3000, d).astype(np.float32)
    x1[[7, 2048], -1] = np.nan
    b1 = np.zeros(3000, np.int64)
    out, perm = ops.sort_pool_fwd(torch.from_numpy(x1).to(DEV), ops.graph_ptr(torch.from_numpy(b1).to(DEV), 1), 10)
    ro, rp = orc.sort_aggregation(torch.from_numpy(x1), torch.from_numpy(b1), 10, 1, return_perm=True)
    np.testing.assert_array_equal(perm.cpu().numpy(), rp.numpy())
    assert perm[0, :2].tolist() == [7, 2048]
    np.testing.assert_array_equal(out.cpu().numpy(), ro.numpy())


def test_sort_pool_backward_and_module():
    rng = np.random.RandomState(9)
    sizes = [3, 0, 50, 12]
    n, d, k = sum(sizes), 97, 10
    x = rng.randn(n, d).astype(np.float32)
    batch = np.repeat(np.arange(4), sizes).astype(np.int64)
    pool = dg.SortAggregation(k)
    assert dg.SortPool is dg.SortAggregation and pool.k == k
    xd = torch.from_numpy(x).to(DEV).requires_grad_(True)
    out = pool(xd, torch.from_numpy(batch).to(DEV))          # B inferred from index.max(), like PyG
    cot = torch.from_numpy(rng.randn(4, k * d).astype(np.float32))
    (out * cot.to(DEV)).sum().backward()
    xr = torch.from_numpy(x).requires_grad_(True)
    (orc.sort_aggregation(xr, torch.from_numpy(batch), k) * cot).sum().backward()
    np.testing.assert_array_equal(xd.grad.cpu().numpy(), xr.grad.numpy())
    # strided destination (ld > d) through the raw op
    perm = pool(xd, torch.from_numpy(batch).to(DEV), return_perm=True)[1]
    buf = torch.full((n, d + 4), 5.0, device=DEV)
    ops.sort_pool_bwd(cot.to(DEV), perm, n, out=buf[:, :d])
    np.testing.assert_array_equal(buf[:, :d].cpu().numpy(), xr.grad.numpy())
    assert (buf[:, d:] == 5.0).all()


# ------------------------------------------------------------------ stack + model
@pytest.fixture(params=["fused-mma", "fused-fma", "per-layer"])
def fused(request):
    """All CUDA implementations of the hot path: the one-launch fused stack kernel in its
    tensor-core (graph_stack_mma.cu) and FMA-gather (graph_stack.cu) variants, and the
    per-layer kernels (graph_conv.cu + sort_pool.cu)."""
    dg.set_fused(request.param != "per-layer")
    old = ops.STACK_VARIANT
    ops.STACK_VARIANT = ops.STACK_FMA if request.param == "fused-fma" else ops.STACK_MMA
    yield request.param
    ops.STACK_VARIANT = old
    dg.set_fused(True)


def max_graph(batch_np, b):
    return int(np.bincount(batch_np, minlength=max(b, 1)).max()) if len(batch_np) else 0


def run_stack(z, norm, k, b):
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    ws = [dev(z[f"w{i}"]).requires_grad_(True) for i in range(1, 5)]
    bs = [dev(z[f"b{i}"]).requires_grad_(True) for i in range(1, 5)]
    x = dev(z["x"]).requires_grad_(True)
    g = ops.build_graph(dev(z["edge_index"]), dev(z["batch"]), x.size(0), b,
                        max_nodes=max_graph(z["batch"], b))
    before = ops.LAUNCHES["stack_fwd"]
    pooled, xcat, perm = dg.graph_conv_stack(x, g, ws, bs, k, norm)
    assert (ops.LAUNCHES["stack_fwd"] - before == 1) == dg.fused_enabled()
    g.check()
    return x, ws, bs, g, pooled, xcat, perm


@pytest.mark.parametrize("name", ["hand_sym", "hand_rw", "mutag6_sym", "proteins5_sym"])
def test_stack_against_golden(name, fused):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    norm, k, b = int(z["norm"]), int(z["k"]), int(z["num_graphs"])
    x, ws, bs, g, pooled, xcat, perm = run_stack(z, norm, k, b)
    assert np.abs(xcat.detach().cpu().numpy() - z["xcat_f32"]).max() <= ATOL
    assert np.abs(xcat.detach().cpu().numpy() - z["xcat_f64"]).max() <= ATOL
    gptr = g.gptr.cpu().numpy()
    swaps = assert_perm_matches(perm.cpu().numpy(), z["perm_f64"], z["xcat_f64"][:, -1], gptr, k)
    if swaps == 0:
        assert np.abs(pooled.detach().cpu().numpy() - z["pooled_f64"]).max() <= ATOL
        (pooled * torch.from_numpy(z["cotangent"]).to(DEV)).sum().backward()
        pairs = [(x.grad, z["dx_f64"])]
        pairs += [(ws[i].grad, z[f"dw{i+1}_f64"]) for i in range(4)]
        pairs += [(bs[i].grad, z[f"db{i+1}_f64"]) for i in range(4)]
        for got, want in pairs:
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(got.cpu().numpy() - want).max() <= 3e-5 * scale


def stack_inputs(rng, sizes, f, avg_deg, loops, seed_w=0):
    ei, batch, n = random_multigraph(rng, list(sizes), avg_deg, loops)
    z = {"x": rng.randn(n, f).astype(np.float32), "edge_index": ei, "batch": batch}
    dims = [(f, 32), (32, 32), (32, 32), (32, 1)]
    for i, (ci, co) in enumerate(dims, 1):
        z[f"w{i}"] = (rng.randn(co, ci) * (1.5 / np.sqrt(ci))).astype(np.float32)
        z[f"b{i}"] = (rng.randn(co) * 0.1).astype(np.float32)
    return z, n


def dedup_symmetric(ei):
    """simple undirected graph from a random multigraph (what TU data looks like)."""
    s, d = ei
    keep = s != d
    pairs = np.unique(np.stack([np.concatenate([s[keep], d[keep]]),
                                np.concatenate([d[keep], s[keep]])], 1), axis=0)
    return np.ascontiguousarray(pairs.T)


STACK_CASES = [
    # name, sizes, F, avg_deg, simple graph?, k, norm
    ("tiny-mixed", [1, 0, 2, 33, 5, 64, 65, 0, 31, 32], 1, 3.0, True, 7, 0),
    ("dense-complement", [40, 70, 90, 33, 128], 1, 60.0, True, 30, 0),
    ("dense-F5", [40, 70, 90, 33], 5, 50.0, True, 60, 0),
    ("rank-vs-bitonic", [255, 256, 257, 300], 8, 6.0, True, 130, 0),
    ("project-first-F19", [20, 45, 100, 7], 19, 4.0, True, 30, 0),
    ("project-first-F90-dense", [60, 80, 35], 90, 40.0, True, 30, 1),
    ("multigraph-csr-path", [30, 50, 70, 3], 6, 8.0, False, 20, 0),
    ("multigraph-F38-rw", [30, 90], 38, 30.0, False, 20, 1),
    ("rw-norm", [50, 60, 10], 3, 5.0, True, 10, 1),
    ("large-n-480", [480, 100], 4, 20.0, True, 130, 0),
    # per-CTA plan of the tensor-core kernels (graph_mma.cuh plan_pass): more graphs than one
    # pass holds (8 teams x 148 CTAs), a giant that gets an SM of its own, and graphs so large
    # that shared memory, not the team count, ends a pass
    ("plan-many-small", [1 + (i * 7) % 40 for i in range(1500)], 3, 3.0, True, 10, 0),
    ("plan-giant-alone", [480] + [12 + (i % 20) for i in range(300)], 2, 5.0, True, 30, 0),
    ("plan-smem-limited", [280] * 200 + [300] * 100, 1, 4.0, True, 130, 0),
]


@pytest.mark.parametrize("variant", ["mma", "fma"])
@pytest.mark.parametrize("case", STACK_CASES, ids=[c[0] for c in STACK_CASES])
def test_fused_stack_forward_matches_oracle_and_per_layer(case, variant):
    name, sizes, f, avg_deg, simple, k, norm = case
    rng = np.random.RandomState(zlib.crc32(name.encode()) % (2 ** 31))
    z, n = stack_inputs(rng, sizes, f, avg_deg, loops=True)
    if simple:
        z["edge_index"] = dedup_symmetric(z["edge_index"])
    b = len(sizes)
    dg.set_fused(False)
    try:
        _, _, _, _, pooled_l, xcat_l, perm_l = run_stack(z, norm, k, b)
    finally:
        dg.set_fused(True)
    old = ops.STACK_VARIANT
    ops.STACK_VARIANT = ops.STACK_FMA if variant == "fma" else ops.STACK_MMA
    try:
        x, ws, bs, g, pooled, xcat, perm = run_stack(z, norm, k, b)
    finally:
        ops.STACK_VARIANT = old
    ref = orc.graph_conv_stack(torch.from_numpy(z["x"]).double(), torch.from_numpy(z["edge_index"]),
                               [torch.from_numpy(z[f"w{i}"]).double() for i in range(1, 5)],
                               [torch.from_numpy(z[f"b{i}"]).double() for i in range(1, 5)], norm)
    err = (xcat.detach().cpu().double() - ref).abs().max().item()
    err_l = (xcat_l.detach().cpu().double() - ref).abs().max().item()
    assert err <= ATOL, f"fused x_cat err {err:.3e} (per-layer {err_l:.3e})"
    assert err_l <= ATOL
    # SortPool inside the fused kernel is bit-exact w.r.t. the oracle run on ITS OWN x_cat
    ro, rp = orc.sort_aggregation(xcat.detach().cpu(), torch.from_numpy(z["batch"]), k, b, return_perm=True)
    np.testing.assert_array_equal(perm.cpu().numpy(), rp.numpy())
    np.testing.assert_array_equal(pooled.detach().cpu().numpy(), ro.numpy())


@pytest.mark.parametrize("variant", ["mma", "fma"])
@pytest.mark.parametrize("case", STACK_CASES, ids=[c[0] for c in STACK_CASES])
def test_fused_stack_backward_matches_oracle_and_per_layer(case, variant):
    """KSB (two launches) against float64 autograd of the oracle and against K3/K4.  The
    oracle continues from OUR permutation (validated bit-exact above) so that near-tied
    keys cannot turn a legitimate rank swap into a gradient mismatch."""
    name, sizes, f, avg_deg, simple, k, norm = case
    rng = np.random.RandomState(zlib.crc32(name.encode()) % (2 ** 31))
    z, n = stack_inputs(rng, sizes, f, avg_deg, loops=True)
    if simple:
        z["edge_index"] = dedup_symmetric(z["edge_index"])
    b = len(sizes)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    g = ops.build_graph(dev(z["edge_index"]), dev(z["batch"]), n, b, max_nodes=max_graph(z["batch"], b))
    cot = torch.from_numpy(rng.randn(b, k * 97).astype(np.float32))

    def run(fused_flag):
        dg.set_fused(fused_flag)
        old_variant = ops.STACK_VARIANT
        ops.STACK_VARIANT = ops.STACK_FMA if variant == "fma" else ops.STACK_MMA
        try:
            ws = [dev(z[f"w{i}"]).requires_grad_(True) for i in range(1, 5)]
            bs = [dev(z[f"b{i}"]).requires_grad_(True) for i in range(1, 5)]
            before = ops.LAUNCHES["stack_bwd"]
            pooled, xcat, perm = dg.graph_conv_stack(dev(z["x"]), g, ws, bs, k, norm)
            (pooled * cot.to(DEV)).sum().backward()
            expect_fused = fused_flag and ops.stack_bwd_supported(f, g.max_nodes)
            assert (ops.LAUNCHES["stack_bwd"] - before == 2) == expect_fused
            return [w.grad.cpu() for w in ws], [b_.grad.cpu() for b_ in bs], perm.cpu().long()
        finally:
            ops.STACK_VARIANT = old_variant
            dg.set_fused(True)

    def oracle_grads(perm):
        wr = [torch.from_numpy(z[f"w{i}"]).double().requires_grad_(True) for i in range(1, 5)]
        br = [torch.from_numpy(z[f"b{i}"]).double().requires_grad_(True) for i in range(1, 5)]
        rx = orc.graph_conv_stack(torch.from_numpy(z["x"]).double(), torch.from_numpy(z["edge_index"]),
                                  wr, br, norm)
        rpool = torch.where((perm >= 0).unsqueeze(-1), rx[perm.clamp(min=0)], rx.new_zeros(()))
        (rpool.reshape(b, k * 97) * cot.double()).sum().backward()
        return [w.grad for w in wr] + [b_.grad for b_ in br]

    # the two CUDA paths round x_4 differently, so exactly tied keys may rank differently:
    # each path is compared with the oracle continued from ITS OWN permutation
    for fused_flag in (True, False):
        dws, dbs, perm = run(fused_flag)
        for got, want in zip(dws + dbs, oracle_grads(perm)):
            scale = max(1.0, float(want.abs().max()))
            err = (got.double() - want).abs().max().item()
            assert err <= 3e-5 * scale, f"fused={fused_flag}: {err:.3e} scale {scale:.2e}"


def test_fused_backward_is_deterministic():
    """Same inputs, same cotangent -> bit-identical parameter gradients (fixed graph->CTA
    assignment, ordered reduction, no float atomics).  The cotangent is fixed rather than
    taken through the dense tail because cuDNN's wgrad is not bit-reproducible."""
    cfg = CONFIGS["collab"]
    batch = make_batch("collab", num_graphs=200)
    data = batch.to(DEV)
    data.max_nodes = int((batch.ptr[1:] - batch.ptr[:-1]).max())
    torch.manual_seed(3)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).eval()
    cot = torch.randn(200, cfg.k * 97, device=DEV)
    grads = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        before = ops.LAUNCHES["stack_bwd"]
        pooled, _, _ = model.hot_path(data.x, model.build_graph(data))
        (pooled * cot).sum().backward()
        assert ops.LAUNCHES["stack_bwd"] - before == 2
        grads.append([p.grad.clone() for p in (model.conv1.lin.weight, model.conv2.lin.weight,
                                               model.conv3.bias, model.conv4.lin.weight)])
    for a, b_ in zip(*grads):
        assert torch.equal(a, b_)


def load_into(model, oracle_model):
    model.load_state_dict(oracle_model.state_dict())     # identical key names
    return model


@pytest.mark.parametrize("name,count", [("mutag", 50), ("proteins", 128), ("dd", 64), ("collab", 96)])
def test_model_forward_backward_matches_oracle(name, count):
    cfg = CONFIGS[name]
    batch = make_batch(name, num_graphs=count, tie_free=True)
    torch.manual_seed(324)
    ref = orc.OracleModel(cfg.num_features, cfg.num_classes, cfg.k).double().eval()
    with torch.no_grad():
        for c in (ref.conv1, ref.conv2, ref.conv3, ref.conv4):
            c.bias.uniform_(-0.1, 0.1)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k)
    model.load_state_dict({k_: v.float() for k_, v in ref.state_dict().items()})
    model = model.to(DEV).eval()
    data = batch.to(DEV)
    data.max_nodes = int((batch.ptr[1:] - batch.ptr[:-1]).max())

    g = model.build_graph(data)
    pooled, xcat, perm = model.hot_path(data.x, g)
    rx, rpool = ref.hot_path(batch.x.double(), batch.edge_index, batch.batch, count)
    assert (xcat.detach().cpu().double() - rx.detach()).abs().max().item() <= ATOL
    _, rperm = orc.sort_aggregation(rx.detach(), batch.batch, cfg.k, count, return_perm=True)
    assert_perm_matches(perm.cpu().numpy(), rperm.numpy(), rx.detach().numpy()[:, -1],
                        batch.ptr.numpy(), cfg.k)
    # Near-tied keys (GCN smoothing makes them common) may legitimately swap ranks, so the
    # oracle continues from OUR validated permutation: gather its own x_cat rows with it.
    pcpu = perm.cpu().long()
    rpool = torch.where((pcpu >= 0).unsqueeze(-1), rx[pcpu.clamp(min=0)], rx.new_zeros(()))
    rpool = rpool.reshape(count, cfg.k * 97)
    assert (pooled.detach().cpu().double() - rpool.detach()).abs().max().item() <= ATOL

    out = model(data)
    rout = ref.tail(rpool)
    assert (out.detach().cpu().double() - rout.detach()).abs().max().item() <= 1e-4
    torch.nn.functional.nll_loss(out, data.y).backward()
    torch.nn.functional.nll_loss(rout, batch.y).backward()
    rp = dict(ref.named_parameters())
    for pname, p in model.named_parameters():
        want = rp[pname].grad
        scale = max(1e-3, float(want.abs().max()))
        err = (p.grad.cpu().double() - want).abs().max().item()
        assert err <= 2e-3 * scale, f"{pname}: {err} vs scale {scale}"


def batch_to_double(b):
    return dg.GraphBatch(b.x.double(), b.edge_index, b.batch, b.ptr, b.y, b.num_graphs)


@pytest.mark.parametrize("path", ["fused", "per-layer"])
def test_full_size_collab_properties(path):
    """BASELINE config 4 at full size (512 graphs, ~2.4M edges) on BOTH CUDA paths (KS, and
    K1 x 4 + K2): determinism, sortedness of the pooled keys, batch-of-1 == batched, and
    direct oracle parity.  The batch carries no max_nodes hint (train.py:36-37)."""
    cfg = CONFIGS["collab"]
    batch = make_batch("collab")
    torch.manual_seed(324)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).eval()
    data = batch.to(DEV)
    dg.set_fused(path == "fused")
    try:
        before = ops.LAUNCHES["stack_fwd"]
        g = model.build_graph(data)
        with torch.no_grad():
            pooled, xcat, perm = model.hot_path(data.x, g)
            pooled2, xcat2, perm2 = model.hot_path(data.x, model.build_graph(data))
        assert ops.LAUNCHES["stack_fwd"] - before == (2 if path == "fused" else 0), "wrong kernel path"
    finally:
        dg.set_fused(True)
    assert torch.equal(xcat, xcat2) and torch.equal(perm, perm2) and torch.equal(pooled, pooled2)
    keys = pooled.view(cfg.batch_size, cfg.k, 97)[:, :, -1]
    valid = perm >= 0
    both = valid[:, 1:] & valid[:, :-1]
    assert (keys[:, :-1][both] >= keys[:, 1:][both]).all()
    assert (pooled.view(cfg.batch_size, cfg.k, 97)[~valid] == 0).all()
    # every valid perm entry points into its own graph and is unique
    gp = g.gptr.long()
    lo, hi = gp[:-1].view(-1, 1), gp[1:].view(-1, 1)
    assert ((perm >= lo) & (perm < hi))[valid].all()
    assert valid.sum(1).tolist() == torch.minimum(hi - lo, torch.tensor(cfg.k, device=DEV)).view(-1).tolist()
    # gathered rows equal x_cat rows
    assert torch.equal(pooled.view(-1, 97)[valid.view(-1)], xcat[perm[valid].long()])
    # oracle parity on x_cat
    ws = [c.lin.weight.detach().cpu() for c in (model.conv1, model.conv2, model.conv3, model.conv4)]
    bs = [c.bias.detach().cpu() for c in (model.conv1, model.conv2, model.conv3, model.conv4)]
    ref = orc.graph_conv_stack(batch.x.double(), batch.edge_index, [w.double() for w in ws],
                               [b.double() for b in bs])            # float64: host-independent
    assert (xcat.cpu().double() - ref).abs().max().item() <= ATOL
    # batch-of-1 == batched for a few graphs
    graphs = make_graphs(cfg, cfg.batch_size, 324)
    for gi in (0, 17, 511):
        one = collate([graphs[gi]]).to(DEV)
        with torch.no_grad():
            _, x1, _ = model.hot_path(one.x, model.build_graph(one))
        lo_i, hi_i = int(batch.ptr[gi]), int(batch.ptr[gi + 1])
        assert (x1 - xcat[lo_i:hi_i]).abs().max().item() <= 1e-6


def test_permutation_equivariance():
    """Relabelling nodes inside each graph leaves the pooled output unchanged
    (tie-free features), up to fp32 summation order."""
    cfg = CONFIGS["proteins"]
    graphs = make_graphs(cfg, 16, seed=2, tie_free=True)
    rng = np.random.RandomState(0)
    shuffled = []
    for gr in graphs:
        n = gr["x"].shape[0]
        p = rng.permutation(n)                 # new id of old node i is p[i]
        inv = np.argsort(p)
        ei = p[gr["edge_index"]]
        o = np.lexsort((ei[1], ei[0]))
        shuffled.append({"x": gr["x"][inv], "edge_index": ei[:, o], "y": gr["y"]})
    torch.manual_seed(1)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).eval()
    with torch.no_grad():
        a = collate(graphs).to(DEV)
        b = collate(shuffled).to(DEV)
        pa, _, _ = model.hot_path(a.x, model.build_graph(a))
        pb, _, _ = model.hot_path(b.x, model.build_graph(b))
    assert (pa - pb).abs().max().item() <= 1e-5


# ------------------------------------------------------------------ dense tail + Adam (8f N2/N3)
@pytest.mark.parametrize("b,k,c", [(37, 30, 2), (64, 130, 3), (5, 61, 2), (3, 291, 2)])
def test_dense_tail_matches_torch(b, k, c):
    """KT (model.py:36-43) forward and backward against the stock-torch tail in float64 on
    the CPU (the reference's own op sequence), eval mode so that dropout is the identity."""
    torch.manual_seed(b * 1000 + k)
    ref = orc.OracleModel(7, c, k).double().eval()
    model = dg.Model(7, c, k)
    model.load_state_dict({n_: v.float() for n_, v in ref.state_dict().items()})
    model = model.to(DEV).eval()
    pooled = torch.randn(b, k * 97)
    pooled[:, ::5] = 0.0                                     # zeros like real zero padding
    cot = torch.randn(b, c)
    pd = pooled.to(DEV).requires_grad_(True)
    dg.set_custom_tail(True)
    before = ops.LAUNCHES["tail_fwd"]
    out = model.tail(pd)
    assert ops.LAUNCHES["tail_fwd"] - before == 5
    (out * cot.to(DEV)).sum().backward()
    pr = pooled.double().requires_grad_(True)
    rout = ref.tail(pr)
    (rout * cot.double()).sum().backward()
    assert (out.detach().cpu().double() - rout.detach()).abs().max().item() <= 2e-5
    scale = max(1.0, float(pr.grad.abs().max()))
    assert (pd.grad.cpu().double() - pr.grad).abs().max().item() <= 2e-5 * scale
    rp = dict(ref.named_parameters())
    for name, p_ in model.named_parameters():
        if name.startswith(("conv5", "conv6", "classifier")):
            want = rp[name].grad
            scale = max(1.0, float(want.abs().max()))
            err = (p_.grad.cpu().double() - want).abs().max().item()
            assert err <= 5e-5 * scale, f"{name}: {err:.3e} (scale {scale:.2e})"
    # and against the stock torch tail on the GPU
    dg.set_custom_tail(False)
    try:
        with torch.no_grad():
            tout = model.tail(pd.detach())
    finally:
        dg.set_custom_tail(True)
    assert (out.detach() - tout).abs().max().item() <= 1e-4


def test_dense_tail_dropout_and_graph_replay():
    torch.manual_seed(0)
    model = dg.Model(5, 2, 60).to(DEV).train()
    pooled = torch.randn(256, 60 * 97, device=DEV)
    _, saved = ops.tail_fwd(pooled, 60, (model.conv5.weight, model.conv5.bias, model.conv6.weight,
                                        model.conv6.bias, model.classifier_1.weight, model.classifier_1.bias,
                                        model.classifier_2.weight, model.classifier_2.bias), True, 123,
                            model._tail_rng_offset)
    h3, keep = saved[4], saved[5]
    alive = (keep > 0).float().mean().item()
    assert int(keep.max()) == 2 and set(keep.unique().tolist()) <= {0, 2}
    assert 0.15 < alive < 0.5                      # P(kept) = 0.5 times P(ReLU alive)
    assert torch.equal(h3 > 0, keep > 0)
    a = model.tail(pooled)
    b_ = model.tail(pooled)
    assert not torch.equal(a, b_)                  # the offset advanced: a fresh mask
    assert int(model._tail_rng_offset.item()) == 3
    model.eval()
    assert torch.equal(model.tail(pooled), model.tail(pooled))
    model.train()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        model.tail(pooled)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = model.tail(pooled)
    g.replay()
    first = out.clone()
    g.replay()
    assert not torch.equal(first, out)             # replays draw new masks too


def test_flat_adam_matches_torch_adam():
    torch.manual_seed(1)
    m1 = torch.nn.Sequential(torch.nn.Linear(13, 7), torch.nn.Linear(7, 3)).to(DEV)
    m2 = torch.nn.Sequential(torch.nn.Linear(13, 7), torch.nn.Linear(7, 3)).to(DEV)
    m2.load_state_dict(m1.state_dict())
    ref = torch.optim.Adam(m1.parameters(), lr=1e-3)
    bucket = dg.GradBucket(m2.parameters(), extra=2)
    opt = dg.FlatAdam(m2, bucket, lr=1e-3)
    x = torch.randn(32, 13, device=DEV)
    for it in range(6):
        for m, o in ((m1, ref), (m2, opt)):
            o.zero_grad()
            (m(x) ** 2).sum().backward()
            o.step()
    for p1, p2 in zip(m1.parameters(), m2.parameters()):
        assert (p1 - p2).abs().max().item() <= 2e-6
    assert int(opt.step_count.item()) == 6


def test_fused_trainer_matches_autograd_training():
    """train.py:35-45: three optimisation steps through FusedTrainer (no autograd, flat
    buffers, in-place gradients) equal three steps of Model + F.nll_loss + torch Adam."""
    cfg = CONFIGS["proteins"]
    batches = []
    for i in range(3):
        hb = make_batch("proteins", seed=10 + i, num_graphs=48)
        d = hb.to(DEV)
        d.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
        batches.append(d)
    torch.manual_seed(5)
    m1 = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).train()
    m2 = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).train()
    m2.load_state_dict(m1.state_dict())
    m2._tail_seed = m1._tail_seed
    ref_opt = torch.optim.Adam(m1.parameters(), lr=1e-3)
    trainer = dg.FusedTrainer(m2, lr=1e-3)
    assert set(m2.state_dict()) == set(m1.state_dict())
    for d in batches:
        ref_opt.zero_grad()
        logp = m1(d)
        loss = torch.nn.functional.nll_loss(logp, d.y)          # mean, like train.py:39
        loss.backward()
        ref_opt.step()
        stats = trainer.step(d)
        assert abs(stats[0].item() / d.num_graphs - loss.item()) <= 1e-5 * max(1.0, abs(loss.item()))
        assert stats[1].item() == float((logp.argmax(1) == d.y).sum())
    p1 = dict(m1.named_parameters())
    for name, p_ in m2.named_parameters():
        err = (p_ - p1[name]).abs().max().item()
        assert err <= 1e-5, f"{name}: {err:.3e}"
    assert int(trainer.step_count.item()) == 3


@pytest.mark.parametrize("name,count,compact", [("proteins", 64, False), ("collab", 48, False), ("mutag", 50, True)])
def test_native_train_step_equals_the_python_sequence(name, count, compact, monkeypatch):
    """dgcnn_train_step (one host call, arena-carved buffers) launches the same kernels in the
    same order as FusedTrainer's one-by-one sequence: parameters, Adam state and the
    loss / accuracy scalars must match bit for bit, step after step, also for int32 batches."""
    import copy
    cfg = CONFIGS[name]
    batches = []
    for i in range(3):
        hb = make_batch(name, seed=11 + i, num_graphs=count)
        mx = int((hb.ptr[1:] - hb.ptr[:-1]).max())
        db = (hb.compact() if compact else hb).to(DEV)
        db.max_nodes = mx
        batches.append(db)
    torch.manual_seed(5)
    model_a = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).train()
    model_b = copy.deepcopy(model_a)
    monkeypatch.setenv("DGCNN_NATIVE_STEP", "1")
    tr_a = dg.FusedTrainer(model_a, lr=1e-3)
    monkeypatch.setenv("DGCNN_NATIVE_STEP", "0")
    tr_b = dg.FusedTrainer(model_b, lr=1e-3)
    assert tr_a.native and not tr_b.native
    for step in range(5):
        before = ops.LAUNCHES.get("train_step", 0)
        sa = tr_a.step(batches[step % 3]).clone()
        assert ops.LAUNCHES.get("train_step", 0) > before, "the native entry point was not used"
        sb = tr_b.step(batches[step % 3]).clone()
        assert torch.equal(sa, sb), (step, sa, sb)
        assert torch.equal(tr_a.flat, tr_b.flat), step
        assert torch.equal(tr_a.exp_avg_sq, tr_b.exp_avg_sq), step
        assert torch.equal(tr_a.grad, tr_b.grad), step
    assert int(tr_a._graph_status[0].item() | tr_a._graph_status[1].item()) & ~ops.GRAPH_GENERIC == 0


def test_cuda_graph_capture_replays_bit_identically():
    cfg = CONFIGS["mutag"]
    batch = make_batch("mutag").to(DEV)
    torch.manual_seed(0)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).eval()
    g = model.build_graph(batch)
    with torch.no_grad():
        eager, _, _ = model.hot_path(batch.x, g)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            model.hot_path(batch.x, g)
        torch.cuda.current_stream().wait_stream(s)
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            g2 = model.build_graph(batch)
            captured, _, _ = model.hot_path(batch.x, g2)
        captured.zero_()
        cg.replay()
    torch.cuda.synchronize()
    assert torch.equal(captured, eager)


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.sort_pool_fwd(torch.zeros(3, 4), torch.zeros(2, dtype=torch.int32), 2)
    m = dg.Model(8, 2)
    with pytest.raises((RuntimeError, TypeError)):
        m(make_batch("mutag", num_graphs=3))


# ------------------------------------------------------------------ K1 vectorised / project-first
@pytest.mark.parametrize("staged", [False, True])
@pytest.mark.parametrize("cout", [32, 1, 7])
@pytest.mark.parametrize("norm", [0, 1])
@pytest.mark.parametrize("sizes", [(17, 1, 0, 40, 9), (300, 3), (2, 2, 2, 2, 2, 2, 2), (1200,), (1800, 5, 1729, 1728)])
def test_graph_conv_vectorised_rows(cout, norm, sizes, staged):
    """gc_aggregate_vec32 (four 32-channel rows per warp, 16-byte neighbour loads) and, with the
    batch's graph offsets (`staged`), gc_aggregate_staged (one CTA per graph, rows in shared
    memory; graphs above 1728 nodes fall to the row-parallel kernel in the same call): aligned
    32-wide inputs in a padded buffer, outputs into a column slice, multigraph input with
    loops, duplicates and edge-free rows; forward against the float64 oracle, backward against
    float64 autograd (K3 takes the same kernel for A_hat^T dpre)."""
    rng = np.random.RandomState(len(sizes) * 100 + cout)
    ei, batch, n = random_multigraph(rng, list(sizes), avg_deg=5.0, loops=True)
    b = len(sizes)
    x = rng.standard_normal((n, 32)).astype(np.float32)
    w = (rng.standard_normal((cout, 32)) * 0.3).astype(np.float32)
    bias = rng.uniform(-0.1, 0.1, cout).astype(np.float32)
    g = gpu_graph(ei, batch, n, b)
    xbuf = torch.zeros(n, 100, device=DEV)
    xbuf[:, 32:64] = torch.from_numpy(x).to(DEV)                       # 16-byte aligned slice
    obuf = torch.full((n, 100), 7.0, device=DEV)
    osl = obuf[:, 64:64 + cout]
    ops.graph_conv_fwd(xbuf[:, 32:64], g.rowptr, g.col, g.dis, torch.from_numpy(w).to(DEV),
                       torch.from_numpy(bias).to(DEV), norm, 1, osl, graph=g if staged else None)
    ref64 = torch.tanh(orc.gcn_conv(torch.from_numpy(x).double(), torch.from_numpy(ei), torch.from_numpy(w).double(),
                                    torch.from_numpy(bias).double(), norm))
    assert (osl.cpu().double() - ref64).abs().max().item() <= ATOL
    assert (obuf[:, :64] == 7.0).all() and (obuf[:, 64 + cout:] == 7.0).all()
    if staged:
        return
    # backward through the module-level autograd function
    xt = torch.from_numpy(x).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    bt = torch.from_numpy(bias).double().requires_grad_(True)
    y64 = torch.tanh(orc.gcn_conv(xt, torch.from_numpy(ei), wt, bt, norm))
    dy = rng.standard_normal((n, cout)).astype(np.float32)
    y64.backward(torch.from_numpy(dy).double())
    xd = xbuf[:, 32:64].detach().clone().requires_grad_(True)
    wd = torch.from_numpy(w).to(DEV).requires_grad_(True)
    bd = torch.from_numpy(bias).to(DEV).requires_grad_(True)
    from dgcnn_b200.nn import _GraphConvFn
    yd = _GraphConvFn.apply(xd, wd, bd, g, norm, 1)
    yd.backward(torch.from_numpy(dy).to(DEV))
    for got, want in ((xd.grad, xt.grad), (wd.grad, wt.grad), (bd.grad, bt.grad)):
        assert (got.cpu().double() - want).abs().max().item() <= 3e-5 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("cin", [33, 64, 90, 128])
def test_graph_conv_projects_first_for_wide_inputs(cin):
    """cin > 32 -> 32 (D&D F = 90, power-law F = 64): dgcnn_project_rows + the vectorised
    aggregation with the identity matrix, against the float64 oracle; NaN rows stay local."""
    rng = np.random.RandomState(cin)
    ei, batch, n = random_multigraph(rng, [50, 0, 333, 7], avg_deg=4.0, loops=True)
    x = rng.standard_normal((n, cin)).astype(np.float32)
    w = (rng.standard_normal((32, cin)) * 0.2).astype(np.float32)
    bias = rng.uniform(-0.1, 0.1, 32).astype(np.float32)
    g = gpu_graph(ei, batch, n, 4)
    out = torch.empty(n, 100, device=DEV)[:, :32]
    before = ops.LAUNCHES["graph_conv_fwd"]
    ops.graph_conv_fwd(torch.from_numpy(x).to(DEV), g.rowptr, g.col, g.dis, torch.from_numpy(w).to(DEV),
                       torch.from_numpy(bias).to(DEV), 0, 1, out)
    assert ops.LAUNCHES["graph_conv_fwd"] - before == 2            # project + aggregate
    ref64 = torch.tanh(orc.gcn_conv(torch.from_numpy(x).double(), torch.from_numpy(ei), torch.from_numpy(w).double(),
                                    torch.from_numpy(bias).double(), 0))
    assert (out.cpu().double() - ref64).abs().max().item() <= ATOL
    # a NaN feature poisons exactly the rows the oracle says it does
    x[5, 3] = np.nan
    ops.graph_conv_fwd(torch.from_numpy(x).to(DEV), g.rowptr, g.col, g.dis, torch.from_numpy(w).to(DEV),
                       torch.from_numpy(bias).to(DEV), 0, 1, out)
    refn = torch.tanh(orc.gcn_conv(torch.from_numpy(x).double(), torch.from_numpy(ei), torch.from_numpy(w).double(),
                                   torch.from_numpy(bias).double(), 0))
    assert torch.equal(torch.isnan(out.cpu()).any(1), torch.isnan(refn).any(1))


@pytest.mark.parametrize("n,k", [(1500, 30), (5748, 291), (20000, 130), (4096, 1), (1025, 256)])
@pytest.mark.parametrize("levels", [3, 50, 0])
def test_sort_pool_selection_with_heavy_ties(n, k, levels):
    """K2's radix SELECT path (graphs with at least 4k nodes and more than 1024): keys drawn from
    a handful of distinct values (ties at the k-th key are settled by ascending node index),
    -0.0 / +0.0, a NaN; permutation and pooled rows bit-exact against the oracle."""
    rng = np.random.RandomState(n + k + levels)
    x = rng.randn(n + 7, 5).astype(np.float32)
    if levels:
        x[:, -1] = rng.randint(0, levels, size=n + 7).astype(np.float32) * 0.25 - 0.5
        x[rng.randint(0, n, 20), -1] = -0.0
        x[rng.randint(0, n, 20), -1] = 0.0
    batch = np.concatenate([np.zeros(n, np.int64), np.ones(7, np.int64)])
    gptr = ops.graph_ptr(torch.from_numpy(batch).to(DEV), 2)
    out, perm = ops.sort_pool_fwd(torch.from_numpy(x).to(DEV), gptr, k, n)
    ro, rp = orc.sort_aggregation(torch.from_numpy(x), torch.from_numpy(batch), k, 2, return_perm=True)
    np.testing.assert_array_equal(perm.cpu().numpy(), rp.numpy())
    np.testing.assert_array_equal(out.cpu().numpy(), ro.numpy())
