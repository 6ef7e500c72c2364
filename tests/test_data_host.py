"""CPU tests of the host side of the data path: TU raw-format reader,
Indegree, fold files, DataLoader-order batching, CSV writers, the collate algebra that
csrc/collate.cu implements.  No GPU, no compute call into the library."""
import os

import numpy as np
import pytest
import torch

from dgcnn_b200 import data as dd
from dgcnn_b200 import driver
from dgcnn_b200.synth import CONFIGS, collate, make_graphs
from oracle import dgcnn_oracle as orc


def write(path, text):
    with open(path, "w") as fh:
        fh.write(text)


def test_tu_reader_hand_fixture(tmp_path):
    """Two graphs written by hand in the TU raw format; expected tensors follow PyG's
    read_tu_data: 1-based ids, self loops removed, duplicates merged, edges sorted by
    (row, col); x = [node_attributes | one-hot(node_labels - min)] then Indegree appended;
    graph labels {-1, 1} -> {0, 1}."""
    root = tmp_path / "TOY" / "raw"
    root.mkdir(parents=True)
    # graph 1: path 1-2-3 (both directions) + a self loop on 2 + a duplicate of (1,2)
    # graph 2: single edge 4-5, node 6 isolated
    write(root / "TOY_A.txt", "2, 1\n1, 2\n2, 3\n3, 2\n2, 2\n1, 2\n4, 5\n5, 4\n")
    write(root / "TOY_graph_indicator.txt", "1\n1\n1\n2\n2\n2\n")
    write(root / "TOY_graph_labels.txt", "1\n-1\n")
    write(root / "TOY_node_labels.txt", "3\n5\n3\n4\n3\n5\n")          # min 3 -> classes 0..2
    write(root / "TOY_node_attributes.txt", "0.5\n1.5\n2.5\n3.5\n4.5\n5.5\n")
    graphs, f, c = dd.read_tu_dataset(str(tmp_path / "TOY"), "TOY")
    assert (f, c, len(graphs)) == (5, 2, 2)
    g0, g1 = graphs
    assert g0["edge_index"].tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]
    assert g1["edge_index"].tolist() == [[0, 1], [1, 0]]
    assert (g0["y"], g1["y"]) == (1, 0)
    want0 = np.array([[0.5, 1, 0, 0, 0.5], [1.5, 0, 0, 1, 1.0], [2.5, 1, 0, 0, 0.5]], np.float32)
    want1 = np.array([[3.5, 0, 1, 0, 1.0], [4.5, 1, 0, 0, 1.0], [5.5, 0, 0, 1, 0.0]], np.float32)
    assert np.array_equal(g0["x"], want0) and np.array_equal(g1["x"], want1)
    # use_node_attr=False drops the attribute column (PyG TUDataset default)
    graphs, f, _ = dd.read_tu_dataset(str(tmp_path / "TOY"), "TOY", use_node_attr=False)
    assert f == 4 and np.array_equal(graphs[0]["x"], want0[:, 1:])


def test_tu_reader_rejects_cross_graph_edges_and_missing_files(tmp_path):
    root = tmp_path / "BAD"
    root.mkdir()
    with pytest.raises(FileNotFoundError):
        dd.read_tu_dataset(str(root), "BAD")
    write(root / "BAD_A.txt", "1, 3\n3, 1\n")
    write(root / "BAD_graph_indicator.txt", "1\n1\n2\n")
    write(root / "BAD_graph_labels.txt", "0\n1\n")
    with pytest.raises(ValueError, match="different graphs"):
        dd.read_tu_dataset(str(root), "BAD")


@pytest.mark.parametrize("name", ["mutag", "proteins", "collab"])
def test_tu_round_trip_of_synthetic_graphs(tmp_path, name):
    """write_tu_dataset -> read_tu_dataset reproduces the synthetic graphs bit for bit:
    edges, labels, one-hot features and the Indegree column (utils.py:18-33)."""
    cfg = CONFIGS[name]
    graphs = make_graphs(cfg, 40, seed=3)
    dd.write_tu_dataset(str(tmp_path / "raw"), "SYN", graphs)
    back, f, c = dd.read_tu_dataset(str(tmp_path), "SYN")
    assert len(back) == len(graphs) and f == cfg.num_features
    assert c == len({g["y"] for g in graphs})
    for a, b in zip(graphs, back):
        assert np.array_equal(a["edge_index"], b["edge_index"])
        assert np.array_equal(a["x"], b["x"], equal_nan=True)
    ys = sorted({g["y"] for g in graphs})
    assert [ys.index(g["y"]) for g in graphs] == [g["y"] for g in back]


def test_indegree_matches_the_oracle_restatement():
    rng = np.random.RandomState(0)
    for n, m in [(7, 12), (1, 0), (5, 0), (30, 200)]:
        ei = np.stack([rng.randint(0, n, m), rng.randint(0, n, m)]).astype(np.int64)
        x = rng.standard_normal((n, 3)).astype(np.float32)
        want = orc.indegree_feature(torch.from_numpy(ei), n, torch.from_numpy(x)).numpy()
        got = dd.indegree(x, ei, n)
        assert np.array_equal(got, want, equal_nan=True)
        want = orc.indegree_feature(torch.from_numpy(ei), n, None).numpy()
        assert np.array_equal(dd.indegree(None, ei, n), want, equal_nan=True)
    # max_value / norm=False / cat=False branches of utils.py:11-33
    ei = np.array([[0, 1, 2], [1, 1, 0]])
    assert dd.indegree(None, ei, 3, norm=False).ravel().tolist() == [1.0, 2.0, 0.0]
    assert dd.indegree(None, ei, 3, max_value=4).ravel().tolist() == [0.25, 0.5, 0.0]
    assert dd.indegree(np.ones(3, np.float32), ei, 3, cat=False).shape == (3, 1)
    assert dd.indegree(np.ones(3, np.float32), ei, 3).shape == (3, 2)


def test_fold_files_and_fallback_split(tmp_path):
    base = tmp_path / "MUTAG" / "10fold_idx"
    base.mkdir(parents=True)
    write(base / "train_idx-3.txt", "5\n1\n9\n")
    write(base / "test_idx-3.txt", "2\n")
    tr, te = dd.load_fold(str(tmp_path), "MUTAG", 3)
    assert tr.tolist() == [5, 1, 9] and te.tolist() == [2] and tr.dtype == np.int64
    tr2, te2 = driver.fold_split(str(tmp_path), "MUTAG", 3, 10)
    assert tr2.tolist() == [5, 1, 9] and te2.tolist() == [2]
    # no files: ten disjoint test sets that cover the data set, train = the rest
    seen = []
    for fold in range(1, 11):
        tr, te = driver.fold_split(str(tmp_path), "NCI1", fold, 103)
        assert len(tr) + len(te) == 103 and not set(tr) & set(te)
        seen += te.tolist()
    assert sorted(seen) == list(range(103))


def test_epoch_batches_follow_the_dataloader_contract():
    ids = np.arange(100, 123)
    got = list(dd.epoch_batches(ids, 10, shuffle=False))
    assert [len(b) for b in got] == [10, 10, 3] and np.concatenate(got).tolist() == ids.tolist()
    g1 = torch.Generator().manual_seed(7)
    g2 = torch.Generator().manual_seed(7)
    a = np.concatenate(list(dd.epoch_batches(ids, 10, True, g1)))
    want = ids[torch.randperm(23, generator=g2).numpy()]           # RandomSampler's draw
    assert a.tolist() == want.tolist() and sorted(a.tolist()) == ids.tolist()
    b = np.concatenate(list(dd.epoch_batches(ids, 10, True, g1)))  # the next epoch reshuffles
    assert a.tolist() != b.tolist()


def test_driver_arguments_and_csv_layout(tmp_path):
    opt = driver.get_args([])
    assert (opt.data_type, opt.batch_size, opt.num_epochs, opt.seed) == ("DD", 50, 100, 324)   # train.py:18-23
    with pytest.raises(SystemExit):
        driver.get_args(["--data_type", "CORA"])
    res = {"train_loss": [0.7, 0.5], "test_loss": [0.8, 0.6], "train_accuracy": [50.0, 75.0],
           "test_accuracy": [40.0, 60.0]}
    path = tmp_path / "MUTAG_results_1.csv"
    driver.write_fold_csv(str(path), res, "epoch")
    lines = path.read_text().strip().split("\n")
    assert lines[0] == "epoch,train_loss,test_loss,train_accuracy,test_accuracy"      # train.py:130-131
    assert lines[1].split(",")[0] == "1" and float(lines[2].split(",")[3]) == 75.0
    pd = pytest.importorskip("pandas")
    frame = pd.read_csv(path, index_col="epoch")
    assert frame.index.tolist() == [1, 2] and frame["test_accuracy"].tolist() == [40.0, 60.0]


def test_collate_algebra_of_the_device_gather():
    """The formulas in csrc/collate.cu's header, in numpy: the CSR of a batch is the
    concatenation of the per-graph CSR segments of the data set with the node and edge
    offsets fixed up -- equal to the CSR built from the host-collated batch."""
    cfg = CONFIGS["proteins"]
    graphs = make_graphs(cfg, 30, seed=4)

    def csr(batch):
        src, dst = batch.edge_index.numpy()
        n = batch.num_nodes
        order = np.lexsort((src, dst))
        rowptr = np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=n))])
        return rowptr, src[order], batch.ptr.numpy()

    ds_rowptr, ds_col, ds_gptr = csr(collate(graphs))
    ids = np.array([7, 3, 3, 29, 0, 11])
    want_rowptr, want_col, want_gptr = csr(collate([graphs[i] for i in ids]))
    nodes = ds_gptr[ids + 1] - ds_gptr[ids]
    gptr = np.concatenate([[0], np.cumsum(nodes)])
    sn0 = ds_gptr[ids]
    se0 = ds_rowptr[sn0]
    eoff = np.concatenate([[0], np.cumsum(ds_rowptr[ds_gptr[ids + 1]] - se0)])
    rowptr = np.zeros(gptr[-1] + 1, dtype=np.int64)
    col = np.zeros(eoff[-1], dtype=np.int64)
    for b in range(len(ids)):
        i = np.arange(nodes[b])
        rowptr[gptr[b] + i] = ds_rowptr[sn0[b] + i] - se0[b] + eoff[b]
        j = np.arange(eoff[b + 1] - eoff[b])
        col[eoff[b] + j] = ds_col[se0[b] + j] - sn0[b] + gptr[b]
    rowptr[-1] = eoff[-1]
    assert np.array_equal(gptr, want_gptr)
    assert np.array_equal(rowptr, want_rowptr) and np.array_equal(col, want_col)


def test_device_dataset_refuses_cpu():
    graphs = make_graphs(CONFIGS["mutag"], 3, seed=0)
    with pytest.raises(RuntimeError, match="CUDA"):
        dd.DeviceDataset(graphs, "cpu")


def test_shard_ids_splits_a_batch_of_graph_ids_across_ranks():
    """SURVEY 8e on the resident path: contiguous slices in batch order, disjoint, covering,
    balanced by nodes + edges; identical on every rank without communication."""
    from dgcnn_b200 import shard_ids
    rng = np.random.RandomState(0)
    nodes = rng.randint(5, 500, size=300)
    edges = nodes * rng.randint(2, 60, size=300)
    ids = rng.permutation(300)[:128]
    for world in (1, 2, 4, 8):
        parts = [shard_ids(ids, nodes, edges, world, r) for r in range(world)]
        assert np.concatenate(parts).tolist() == ids.tolist()
        cost = [float((nodes[p] + edges[p]).sum()) for p in parts]
        assert max(cost) <= sum(cost) / world + float((nodes[ids] + edges[ids]).max())
    assert shard_ids(ids[:1], nodes, edges, 2, 1).tolist() == ids[:1].tolist()   # the last rank takes the rest


GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def golden_collate_case():
    z = np.load(os.path.join(GOLDEN, "collate_case.npz"))
    graphs = [{"x": z[f"g{i}_x"], "edge_index": z[f"g{i}_edge_index"], "y": int(z["ys"][i])}
              for i in range(len(z["sizes"]))]
    return z, graphs


def test_collate_golden_pins_the_oracle_and_the_host_loader():
    """tests/golden/collate_case.npz: Batch.from_data_list (train.py:108-109) and the batch's
    graph structure, restated in the oracle, against the committed vectors, against the
    product's host loader (synth.collate) and against an independent numpy CSR; a few entries
    by hand."""
    z, graphs = golden_collate_case()
    ids = z["ids"]
    x, ei, batch, ptr, y = orc.from_data_list([(graphs[i]["x"], graphs[i]["edge_index"], graphs[i]["y"]) for i in ids])
    for name, t in (("x", x), ("edge_index", ei), ("batch", batch), ("ptr", ptr), ("y", y)):
        assert np.array_equal(t.numpy(), z[name]), name
    assert ptr.tolist() == [0, 5, 9, 9, 12, 16, 17, 20] and y.tolist() == [0, 1, 1, 1, 1, 0, 0]
    hb = collate([graphs[i] for i in ids])                       # the product's host loader
    assert torch.equal(hb.x, x) and torch.equal(hb.edge_index, ei) and torch.equal(hb.batch, batch)
    assert torch.equal(hb.ptr, ptr) and torch.equal(hb.y, y)
    rowptr, col, rowptr_t, col_t, dis = orc.batch_csr(ei, 20)
    for name, t in (("rowptr", rowptr), ("col", col), ("rowptr_t", rowptr_t), ("col_t", col_t), ("dis", dis)):
        assert np.array_equal(t.numpy(), z[name]), name
    # independent numpy restatement (lexsort), and hand-checked entries
    src, dst = ei.numpy()
    keep = src != dst
    src, dst = src[keep], dst[keep]
    assert src.size == 29 == int(rowptr[-1])                     # 8 + 7 + 0 + 3 + 7 + 0 + 4, the self loops dropped
    o = np.lexsort((src, dst))
    assert np.array_equal(col.numpy(), src[o])
    assert np.array_equal(rowptr.numpy(), np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=20))]))
    ot = np.lexsort((dst, src))
    assert np.array_equal(col_t.numpy(), dst[ot])
    assert rowptr[:6].tolist() == [0, 4, 5, 6, 7, 8]             # the star: hub 0 has in-degree 4
    assert np.allclose(dis[:5].numpy(), [5 ** -0.5] + [2 ** -0.5] * 4)
    assert col[8:11].tolist() == [6, 7, 5]                       # triangle node 5: sources 6 (0->1), 7 (2->1); node 6: 5
    assert z["gorder"].tolist() == [0, 1, 4, 3, 6, 5, 2]         # sizes 5,4,0,3,4,1,3: descending, ties by index


def test_indegree_transform_object_has_the_reference_interface():
    """utils.py:5-36: Indegree(norm, max_value, cat)(data) -> data, on tensors, against the oracle."""
    from types import SimpleNamespace
    rng = np.random.RandomState(1)
    ei = torch.from_numpy(np.stack([rng.randint(0, 9, 30), rng.randint(0, 9, 30)]))
    x = torch.from_numpy(rng.standard_normal((9, 4)).astype(np.float32))
    data = SimpleNamespace(x=x.clone(), edge_index=ei, num_nodes=9)
    out = dd.Indegree()(data)
    assert out is data and isinstance(data.x, torch.Tensor) and data.x.dtype == torch.float32
    assert torch.equal(data.x, orc.indegree_feature(ei, 9, x))
    data = SimpleNamespace(x=None, edge_index=ei, num_nodes=9)                 # COLLAB / IMDB: no node features
    assert torch.equal(dd.Indegree()(data).x, orc.indegree_feature(ei, 9, None))
    data = SimpleNamespace(x=x[:, 0].clone(), edge_index=ei, num_nodes=9)      # 1-D x is viewed as a column
    assert dd.Indegree()(data).x.shape == (9, 2)
    data = SimpleNamespace(x=x.clone(), edge_index=ei, num_nodes=9)
    assert dd.Indegree(cat=False)(data).x.shape == (9, 1)
    raw = dd.Indegree(norm=False)(SimpleNamespace(x=None, edge_index=ei, num_nodes=9)).x
    assert torch.equal(raw.ravel(), torch.bincount(ei[1], minlength=9).float())
    assert repr(dd.Indegree(max_value=3)) == "Indegree(norm=True, max_value=3)"


def test_resident_loader_iterates_like_the_reference_dataloader():
    """train.py:108-109 / 31-32: len(loader) batches, len(loader.dataset) graphs, DataLoader order;
    the gather itself is the data set's business (a stand-in records the ids it is asked for)."""
    class Recorder:
        def __init__(self):
            self.seen = []

        def batch(self, ids):
            self.seen.append(np.asarray(ids).tolist())
            return len(ids)

    rec = Recorder()
    ids = np.arange(50, 73)
    loader = dd.ResidentLoader(rec, ids, batch_size=10, shuffle=False)
    assert len(loader) == 3 and len(loader.dataset) == 23
    assert list(loader) == [10, 10, 3] and sum(rec.seen, []) == ids.tolist()
    g1, g2 = torch.Generator().manual_seed(3), torch.Generator().manual_seed(3)
    rec2 = Recorder()
    list(dd.ResidentLoader(rec2, ids, batch_size=8, shuffle=True, generator=g1))
    assert sum(rec2.seen, []) == ids[torch.randperm(23, generator=g2).numpy()].tolist()
    with pytest.raises(ValueError):
        dd.ResidentLoader(rec, ids, batch_size=0)


def test_epoch_plan_equals_the_per_batch_sums():
    """driver.epoch_plan (one vectorised pass for a whole epoch) against the per-batch arithmetic of
    DeviceDataset.plan: nodes, edges and the largest graph of every batch, short last batch included;
    the batches are the ones epoch_batches yields for the same permutation."""
    rng = np.random.RandomState(3)
    nodes = rng.randint(1, 500, size=1037).astype(np.int64)
    edges = (nodes * rng.randint(0, 30, size=nodes.size)).astype(np.int64)
    for batch_size in (1, 7, 512, 1037, 5000):
        gen = torch.Generator().manual_seed(batch_size)
        ids = np.arange(nodes.size, dtype=np.int64)
        order = ids[torch.randperm(ids.size, generator=gen).numpy()]
        starts, n_b, e_b, mx_b = driver.epoch_plan(order, nodes, edges, batch_size)
        gen = torch.Generator().manual_seed(batch_size)
        batches = list(dd.epoch_batches(ids, batch_size, True, gen))
        assert len(batches) == len(starts)
        for i, b in enumerate(batches):
            np.testing.assert_array_equal(b, order[starts[i]:starts[i] + batch_size])
            assert (int(n_b[i]), int(e_b[i]), int(mx_b[i])) == \
                (int(nodes[b].sum()), int(edges[b].sum()), int(nodes[b].max()))


def test_epoch_stats_reproduce_the_reference_running_means():
    """EpochStats (train.py:33, 44-45): mean of the per-batch MEAN losses and accuracy over all
    samples, from per-step [sum of NLL, #correct] pairs; unequal last batch."""
    st = driver.EpochStats(torch.device("cpu"), steps=2)
    sums = [(10.0, 3.0, 8), (4.0, 5.0, 8), (1.5, 1.0, 3)]          # more steps than announced: grows
    for s, c, g in sums:
        st.add(torch.tensor([s, c, 99.0]), g)
    loss, acc = st.result()
    assert abs(loss - (10.0 / 8 + 4.0 / 8 + 1.5 / 3) / 3) < 1e-12
    assert abs(acc - 9.0 / 19 * 100.0) < 1e-12
    assert driver.EpochStats(torch.device("cpu")).result() == (0.0, 0.0)
