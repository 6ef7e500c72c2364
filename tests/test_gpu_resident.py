"""GPU parity tests of the resident data set: dgcnn_collate and
dgcnn_train_step_resident through the ctypes binding, against

  * K0 (dgcnn_build_graph, itself checked against the oracle's CSR in test_gpu_parity.py) run on
    the HOST-collated batch of the same graphs -- the restatement of PyG's
    Batch.from_data_list in dgcnn_b200/synth.py::collate: every array bit-exact;
  * the host-fed training step (dgcnn_train_step): parameters, Adam state, loss and accuracy
    bit-identical step after step;
  * the CPU oracle on the host-collated batch (x_cat within 1e-5).
"""
import copy
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import dgcnn_b200 as dg
from dgcnn_b200 import _lib, ops
from dgcnn_b200.synth import CONFIGS, collate, make_graphs
from oracle import dgcnn_oracle as orc

DEV = "cuda:0"
ATOL = 1e-5


def host_graph(graphs, ids, compact=False):
    hb = collate([graphs[int(i)] for i in ids])
    mx = int((hb.ptr[1:] - hb.ptr[:-1]).max())
    db = (hb.compact() if compact else hb).to(DEV)
    db.max_nodes = mx
    return hb, db


def multigraphs(rng, sizes, f=3, avg_deg=3.0):
    """Directed multigraphs with duplicates, self loops, isolated nodes, unsorted edges."""
    out = []
    for n in sizes:
        m = int(avg_deg * n)
        ei = np.stack([rng.randint(0, max(n, 1), size=m), rng.randint(0, max(n, 1), size=m)]).astype(np.int64) \
            if n > 0 else np.zeros((2, 0), np.int64)
        out.append({"x": rng.standard_normal((n, f)).astype(np.float32), "edge_index": ei,
                    "y": int(rng.randint(0, 2))})
    return out


def assert_same_graph(got: ops.Graph, want: ops.Graph, e: int):
    for name in ("rowptr", "dis", "gptr", "gorder"):
        assert torch.equal(getattr(got, name), getattr(want, name)), name
    assert torch.equal(got.col[:e], want.col[:e]), "col"
    assert torch.equal(got.rowptr_t, want.rowptr_t), "rowptr_t"
    assert torch.equal(got.col_t[:e], want.col_t[:e]), "col_t"


ID_CASES = {
    "prefix": lambda g, rng: np.arange(min(g, 50)),
    "reversed": lambda g, rng: np.arange(g)[::-1].copy(),
    "shuffled_with_repeats": lambda g, rng: rng.randint(0, g, size=g + 7),
    "single": lambda g, rng: np.array([g // 2]),
}


@pytest.mark.parametrize("ids_kind", list(ID_CASES))
@pytest.mark.parametrize("name,count", [("mutag", 60), ("proteins", 96), ("collab", 40), ("dd", 12)])
def test_collate_equals_k0_on_the_host_collated_batch(name, count, ids_kind):
    cfg = CONFIGS[name]
    graphs = make_graphs(cfg, count, seed=7)
    ds = dg.DeviceDataset(graphs, DEV, num_classes=cfg.num_classes)
    assert ds.symmetric and ds.rowptr_t is None
    ids = ID_CASES[ids_kind](count, np.random.RandomState(3))
    hb, db = host_graph(graphs, ids)
    want = ops.build_graph(db.edge_index, db.batch, hb.num_nodes, len(ids), transpose=True, max_nodes=db.max_nodes)
    rb = ds.batch(ids)
    n, e, mx = ds.plan(ids)
    assert (n, e, mx) == (hb.num_nodes, hb.num_edges, db.max_nodes)
    assert_same_graph(rb._dgcnn_graph, want, e)
    assert torch.equal(rb.x, db.x) and torch.equal(rb.y, db.y)
    assert torch.equal(rb.batch.long(), db.batch)
    assert int(rb._dgcnn_graph.status.item()) == int(want.status.item()) == 0
    # K0b on the gathered CSR: same bitmaps, fragment maps and work descriptors
    if want.bitmap is not None:
        got = rb._dgcnn_graph
        for name_ in ("bmoff", "gflags", "fgoff", "gdesc"):
            assert torch.equal(getattr(got, name_), getattr(want, name_)), name_
        # the word counts are upper bounds: compare what the graphs own
        bm_end, fg_end = int(want.bmoff[-1]), int(want.fgoff[-1])
        assert torch.equal(got.bitmap[:bm_end], want.bitmap[:bm_end]), "bitmap"
        assert torch.equal(got.fragmap[:fg_end], want.fragmap[:fg_end]), "fragmap"


@pytest.mark.parametrize("sizes", [[5, 0, 7, 1, 0, 9], [40, 3, 3, 17], [1, 1, 1], [300, 2]])
def test_collate_of_a_generic_data_set(sizes):
    """Non-symmetric multigraphs (loops, duplicates, empty graphs): K0 takes the generic path
    for the data set, both CSRs are kept, and a gathered batch equals K0 on that batch."""
    rng = np.random.RandomState(11)
    graphs = multigraphs(rng, sizes)
    ds = dg.DeviceDataset(graphs, DEV, num_classes=2)
    assert not ds.symmetric and ds.rowptr_t is not None
    for ids in (np.arange(len(sizes)), np.arange(len(sizes))[::-1].copy(),
                rng.randint(0, len(sizes), size=2 * len(sizes) + 1)):
        if ds.nodes[ids].sum() == 0:
            continue
        hb, db = host_graph(graphs, ids)
        want = ops.build_graph(db.edge_index, db.batch, hb.num_nodes, len(ids), transpose=True,
                               max_nodes=db.max_nodes)
        rb = ds.batch(ids)
        e = ds.plan(ids)[1]
        assert e == int(want.rowptr[-1].item())                 # loops dropped
        got = rb._dgcnn_graph
        assert_same_graph(got, want, e)
        # K0b's outputs gathered from the data set's cache: both bitmaps, duplicate flags, fragment maps
        for name_ in ("bmoff", "gflags", "gflags_t", "fgoff", "gdesc"):
            assert torch.equal(getattr(got, name_), getattr(want, name_)), name_
        bm_end, fg_end = int(want.bmoff[-1]), int(want.fgoff[-1])
        assert torch.equal(got.bitmap[:bm_end], want.bitmap[:bm_end]), "bitmap"
        assert torch.equal(got.bitmap_t[:bm_end], want.bitmap_t[:bm_end]), "bitmap_t"
        assert torch.equal(got.fragmap[:fg_end], want.fragmap[:fg_end]), "fragmap"
        assert torch.equal(rb.x, db.x) and torch.equal(rb.batch.long(), db.batch)
        st = int(rb._dgcnn_graph.status.item())
        assert st & ops.GRAPH_GENERIC and not st & (ops.GRAPH_BAD_EDGE | ops.GRAPH_BAD_BATCH)


def test_collate_many_graphs_takes_the_global_table_path():
    """More graphs per batch than the shared-memory offset tables (2048) and than the rank
    sort (4096: identity order, as K0) hold."""
    cfg = CONFIGS["mutag"]
    graphs = make_graphs(cfg, 64, seed=5)
    ds = dg.DeviceDataset(graphs, DEV)
    rng = np.random.RandomState(0)
    for b in (2049, 4500):
        ids = rng.randint(0, 64, size=b)
        hb, db = host_graph(graphs, ids)
        want = ops.build_graph(db.edge_index, db.batch, hb.num_nodes, b, transpose=True, max_nodes=0)
        rb = ds.batch(ids, bitmaps=False)
        assert_same_graph(rb._dgcnn_graph, want, hb.num_edges)
        assert torch.equal(rb.x, db.x) and torch.equal(rb.y, db.y)


def test_collate_flags_bad_ids_and_totals():
    cfg = CONFIGS["mutag"]
    graphs = make_graphs(cfg, 20, seed=1)
    ds = dg.DeviceDataset(graphs, DEV)
    with pytest.raises(IndexError):
        ds.plan([0, 20])
    with pytest.raises(IndexError):
        ds.plan([])
    lib = _lib.load_library()
    ids = np.arange(10)
    n, e, _ = ds.plan(ids)
    i32 = dict(dtype=torch.int32, device=DEV)

    def run(ids_np, n_, e_):
        b = len(ids_np)
        ids_dev = torch.from_numpy(np.asarray(ids_np, dtype=np.int32)).to(DEV)
        rowptr, col = torch.full((n_ + 1,), -7, **i32), torch.full((max(e_, 1),), -7, **i32)
        dis = torch.empty(n_, device=DEV)
        gptr, status = torch.empty(b + 1, **i32), torch.zeros(1, **i32)
        ws = torch.empty(int(lib.dgcnn_collate_workspace_bytes(b)), dtype=torch.uint8, device=DEV)
        out = _lib.DgcnnBatchGraph()                             # everything optional stays NULL
        out.rowptr, out.col, out.dis, out.gptr = rowptr.data_ptr(), col.data_ptr(), dis.data_ptr(), gptr.data_ptr()
        rc = lib.dgcnn_collate(ds.c_struct, ids_dev.data_ptr(), b, n_, e_, ctypes.byref(out), status.data_ptr(),
                               ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        return int(status.item()), rowptr, col

    st, rowptr, col = run(ids, n, e)
    assert st == 0 and int(rowptr[-1]) == e and int(col.min()) >= 0
    st, rowptr, col = run(ids, n + 1, e)                        # totals that do not match the ids
    assert st & ops.GRAPH_BAD_BATCH and int(col[0]) == -7       # nothing gathered
    bad = ids.copy()
    bad[3] = 20                                                 # id outside the data set
    st, _, col = run(bad, n, e)
    assert st & ops.GRAPH_BAD_BATCH and int(col[0]) == -7
    # argument errors come back as codes, never as exceptions or crashes
    empty = _lib.DgcnnBatchGraph()
    assert lib.dgcnn_collate(ds.c_struct, None, 1, 1, 0, ctypes.byref(empty), None, None, 0, None) == -1
    ids_dev = torch.zeros(1, **i32)
    assert lib.dgcnn_collate(ds.c_struct, ids_dev.data_ptr(), 1, 1, 0, ctypes.byref(empty), None, None, 0,
                             None) == -1                        # required outputs missing


@pytest.mark.parametrize("name,count,bs", [("proteins", 150, 64), ("collab", 70, 32), ("mutag", 120, 50)])
def test_resident_train_step_is_bit_identical_to_the_host_fed_step(name, count, bs):
    """train.py:35-45 on ids of a resident data set == the same step on the host-collated,
    host-to-device-copied batch: parameters, Adam state, gradients, loss and accuracy bit for
    bit, over a shuffled epoch with a short last batch."""
    cfg = CONFIGS[name]
    graphs = make_graphs(cfg, count, seed=21)
    ds = dg.DeviceDataset(graphs, DEV, num_classes=cfg.num_classes)
    torch.manual_seed(5)
    model_a = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).train()
    model_b = copy.deepcopy(model_a)
    tr_a, tr_b = dg.FusedTrainer(model_a, lr=1e-3), dg.FusedTrainer(model_b, lr=1e-3)
    gen = torch.Generator().manual_seed(0)
    steps = 0
    for ids in dg.epoch_batches(np.arange(count), bs, shuffle=True, generator=gen):
        before = ops.LAUNCHES.get("train_step_resident", 0)
        sa = tr_a.step_resident(ds, ids).clone()
        assert ops.LAUNCHES.get("train_step_resident", 0) > before
        _, db = host_graph(graphs, ids)
        sb = tr_b.step(db).clone()
        assert torch.equal(sa, sb), (steps, sa, sb)
        assert torch.equal(tr_a.flat, tr_b.flat), steps
        assert torch.equal(tr_a.grad, tr_b.grad), steps
        assert torch.equal(tr_a.exp_avg_sq, tr_b.exp_avg_sq), steps
        steps += 1
    assert steps == -(-count // bs) and int(tr_a.step_count.item()) == steps
    assert tr_a._graph_status.tolist() == [0, 0]
    assert torch.isfinite(tr_a.flat).all()


def test_resident_batch_through_the_module_api_matches_host_batch_and_oracle():
    """train.py:60 (evaluation): Model(data) on a gathered batch == Model(data) on the
    host-collated batch (bit-exact), and x_cat is within 1e-5 of the float64 oracle."""
    cfg = CONFIGS["proteins"]
    graphs = make_graphs(cfg, 80, seed=9, tie_free=True)
    ds = dg.DeviceDataset(graphs, DEV, num_classes=cfg.num_classes)
    ids = np.random.RandomState(2).permutation(80)[:48]
    hb, db = host_graph(graphs, ids)
    torch.manual_seed(324)
    ref = orc.OracleModel(cfg.num_features, cfg.num_classes, cfg.k).double().eval()
    with torch.no_grad():
        for c in (ref.conv1, ref.conv2, ref.conv3, ref.conv4):
            c.bias.uniform_(-0.1, 0.1)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k)
    model.load_state_dict({k_: v.float() for k_, v in ref.state_dict().items()})
    model = model.to(DEV).eval()
    rb = ds.batch(ids)
    with torch.no_grad():
        out_r, out_h = model(rb), model(db)
        pooled, xcat, perm = model.hot_path(rb.x, rb._dgcnn_graph)
    assert torch.equal(out_r, out_h)
    rx, _ = ref.hot_path(hb.x.double(), hb.edge_index, hb.batch, len(ids))
    assert (xcat.cpu().double() - rx.detach()).abs().max().item() <= ATOL
    # SortPooling on OUR keys: bit-exact against the oracle's stable descending sort
    _, rperm = orc.sort_aggregation(xcat.cpu(), hb.batch, cfg.k, len(ids), return_perm=True)
    assert torch.equal(perm.cpu().long(), rperm)


def test_reference_driver_two_folds_on_a_tu_format_stand_in(tmp_path):
    """train.py:69-148 through dgcnn_b200.driver: TU-format files on disk -> resident data set
    -> 2 folds x 3 epochs; the reference's output files appear with its names and columns,
    the checkpoints load into Model, and the first epoch's numbers equal a by-hand replay of
    the same protocol through FusedTrainer.step on host-collated batches."""
    from dgcnn_b200 import driver
    argv = ["--data_type", "MUTAG", "--batch_size", "16", "--num_epochs", "3", "--seed", "324",
            "--data_root", str(tmp_path / "data"), "--out_root", str(tmp_path), "--folds", "2",
            "--synthetic", "--synthetic_graphs", "70"]
    over = driver.main(argv)
    assert len(over["train_accuracy"]) == 2 and all(0.0 <= a <= 100.0 for a in over["test_accuracy"])
    for fold in (1, 2):
        lines = (tmp_path / "statistics" / f"MUTAG_results_{fold}.csv").read_text().strip().split("\n")
        assert lines[0] == "epoch,train_loss,test_loss,train_accuracy,test_accuracy" and len(lines) == 4
        assert all(np.isfinite([float(v) for v in ln.split(",")]).all() for ln in lines[1:])
        state = torch.load(tmp_path / "epochs" / f"MUTAG_{fold}.pth", map_location="cpu")
        dg.Model(8, 2).load_state_dict(state)
    overall = (tmp_path / "statistics" / "MUTAG_results_overall.csv").read_text().strip().split("\n")
    assert overall[0] == "fold,train_accuracy,test_accuracy" and len(overall) == 3

    # replay fold 1, epoch 1 by hand: same seeds, same shuffles, host-collated batches
    driver.set_determ(324)
    graphs, f, c = dg.read_tu_dataset(str(tmp_path / "data" / "MUTAG"), "MUTAG")
    assert (f, c, len(graphs)) == (8, 2, 70)
    gen = torch.Generator().manual_seed(324)
    model = dg.Model(f, c, k=30).to(DEV)
    trainer = dg.FusedTrainer(model)
    train_idx, _ = driver.fold_split(str(tmp_path / "data"), "MUTAG", 1, 70, 2, 324)
    model.train()
    loss_sum, correct, batches = 0.0, 0.0, 0
    for ids in dg.epoch_batches(train_idx, 16, True, gen):
        _, db = host_graph(graphs, ids)
        st = trainer.step(db)
        loss_sum += float(st[0]) / len(ids)
        correct += float(st[1])
        batches += 1
    first = (tmp_path / "statistics" / "MUTAG_results_1.csv").read_text().strip().split("\n")[1].split(",")
    assert abs(float(first[1]) - loss_sum / batches) <= 1e-5 * max(1.0, abs(loss_sum / batches))
    assert abs(float(first[3]) - correct / len(train_idx) * 100.0) <= 1e-4


def test_collate_against_the_committed_golden_vectors():
    """tests/golden/collate_case.npz (generated from the oracle's Batch.from_data_list + CSR
    restatement, pinned on the CPU by tests/test_data_host.py): six hand-sized graphs incl. an
    empty one, a duplicated edge, a self loop and a directed cycle; a batch with a repeated id."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "collate_case.npz"))
    graphs = [{"x": z[f"g{i}_x"], "edge_index": z[f"g{i}_edge_index"], "y": int(z["ys"][i])}
              for i in range(len(z["sizes"]))]
    ds = dg.DeviceDataset(graphs, DEV, num_classes=2)
    assert not ds.symmetric and ds.num_edges == 4 + 7 + 3 + 8
    rb = ds.batch(z["ids"])
    g = rb._dgcnn_graph
    e = int(z["rowptr"][-1])
    np.testing.assert_array_equal(g.rowptr.cpu().numpy(), z["rowptr"])
    np.testing.assert_array_equal(g.col.cpu().numpy()[:e], z["col"])
    np.testing.assert_array_equal(g.rowptr_t.cpu().numpy(), z["rowptr_t"])
    np.testing.assert_array_equal(g.col_t.cpu().numpy()[:e], z["col_t"])
    np.testing.assert_allclose(g.dis.cpu().numpy(), z["dis"], rtol=5e-7, atol=0)   # 1/sqrtf vs torch pow(-0.5)
    np.testing.assert_array_equal(g.gptr.cpu().numpy(), z["ptr"])
    np.testing.assert_array_equal(g.gorder.cpu().numpy(), z["gorder"])
    np.testing.assert_array_equal(rb.x.cpu().numpy(), z["x"])
    np.testing.assert_array_equal(rb.batch.cpu().numpy(), z["batch"])
    np.testing.assert_array_equal(rb.y.cpu().numpy(), z["y"])
    st = int(g.status.item())
    assert st & ops.GRAPH_GENERIC and not st & (ops.GRAPH_BAD_EDGE | ops.GRAPH_BAD_BATCH)


def test_driver_epoch_falls_back_for_graphs_beyond_the_fused_kernels():
    """ADVICE r01 (driver.py:129): a data set with graphs of 1100 and 650 nodes (D&D / PROTEINS
    tails) through driver.train_epoch -- batches that hold them take FusedTrainer.step_autograd
    (per-layer kernels), the others the one-call resident step; the epoch finishes, the loss is
    finite and equals a by-hand replay that uses Model(data) + torch Adam for every batch."""
    from dgcnn_b200 import driver
    cfg = CONFIGS["proteins"]
    graphs = make_graphs(cfg, 30, seed=8, tie_free=True)
    rng = np.random.RandomState(4)
    for n in (1100, 650):
        m = 3 * n
        a, b_ = rng.randint(0, n, m), rng.randint(0, n, m)
        keep = a != b_
        lo, hi = np.minimum(a, b_)[keep], np.maximum(a, b_)[keep]
        pairs = np.unique(np.stack([lo, hi], 1), axis=0)
        src = np.concatenate([pairs[:, 0], pairs[:, 1]])
        dst = np.concatenate([pairs[:, 1], pairs[:, 0]])
        order = np.argsort(src * n + dst, kind="stable")
        graphs.append({"x": rng.standard_normal((n, cfg.num_features)).astype(np.float32),
                       "edge_index": np.stack([src[order], dst[order]]), "y": int(rng.randint(0, 2))})
    ds = dg.DeviceDataset(graphs, DEV, num_classes=2)
    ids = np.arange(len(graphs), dtype=np.int64)
    torch.manual_seed(1)
    model = dg.Model(cfg.num_features, 2, k=cfg.k).to(DEV)
    ref_model = copy.deepcopy(model)
    trainer = dg.FusedTrainer(model)
    assert not trainer.resident_supported(ds, ids) and trainer.resident_supported(ds, ids[:8])
    gen = torch.Generator().manual_seed(5)
    loss, acc = driver.train_epoch(trainer, ds, ids, 8, gen)
    assert np.isfinite(loss) and 0.0 <= acc <= 100.0
    assert int(trainer.step_count.item()) == 4
    # replay: same shuffles, every batch through Model(data) + autograd + torch Adam (dropout uses
    # the model's own counter-hash stream: same seed, same offsets)
    ref_model._tail_seed = model._tail_seed
    ref_model.train()
    opt = torch.optim.Adam(ref_model.parameters(), lr=1e-3)
    gen = torch.Generator().manual_seed(5)
    total = 0.0
    for b in dg.epoch_batches(ids, 8, True, gen):
        data = ds.batch(b)
        opt.zero_grad()
        out = ref_model(data)
        l = torch.nn.functional.nll_loss(out, data.y)
        l.backward()
        opt.step()
        total += float(l)
    assert abs(loss - total / 4) <= 1e-4 * max(1.0, abs(total / 4))
    for (n_, p_), (_, q_) in zip(model.named_parameters(), ref_model.named_parameters()):
        assert (p_ - q_).abs().max().item() <= 5e-5, n_


def test_graph_replayed_resident_step_is_bit_identical_to_the_eager_one():
    """dgcnn_train_step_resident_graphed (the step captured on every call, the executable graph
    updated in place, then launched) against the eagerly launched resident step: statistics and
    parameters bit for bit over batches of changing size -- including a change of topology (a batch
    whose largest graph takes the conv5 fusion away re-instantiates the graph) -- and the library's
    own counters say that the graph path really ran."""
    cfg = CONFIGS["collab"]
    graphs = make_graphs(cfg, 160, seed=12)
    ds = dg.DeviceDataset(graphs, DEV, num_classes=cfg.num_classes)
    rng = np.random.RandomState(0)
    batches = [np.sort(rng.choice(len(graphs), size=sz, replace=False)) for sz in (64, 17, 100, 64, 3)]
    torch.manual_seed(2)
    base = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).train()
    lib = _lib.load_library()
    before = (ctypes.c_int64 * 4)()
    lib.dgcnn_train_step_graph_counts(ctypes.cast(before, ctypes.c_void_p))
    results = {}
    for graphed in (False, True):
        tr = dg.FusedTrainer(copy.deepcopy(base))
        stream = tr.graph_stream() if graphed else torch.cuda.current_stream()
        stream.wait_stream(torch.cuda.current_stream())
        stats = []
        with torch.cuda.stream(stream):
            for step, ids in enumerate(batches):
                if step == 3:
                    ops.set_fuse_conv5(False)                 # another launch sequence: new topology
                try:
                    stats.append(tr.step_resident(ds, ids, graphed=graphed).clone())
                finally:
                    ops.set_fuse_conv5(True)
            tr.check_status()
        torch.cuda.current_stream().wait_stream(stream)
        torch.cuda.synchronize()
        results[graphed] = (stats, tr.flat.clone())
    for a, b_ in zip(results[False][0], results[True][0]):
        assert torch.equal(a, b_)
    assert torch.equal(results[False][1], results[True][1])
    after = (ctypes.c_int64 * 4)()
    lib.dgcnn_train_step_graph_counts(ctypes.cast(after, ctypes.c_void_p))
    updated, instantiated, eager, failed = (int(after[i] - before[i]) for i in range(4))
    assert failed == 0 and updated + instantiated >= 4 and instantiated >= 1
