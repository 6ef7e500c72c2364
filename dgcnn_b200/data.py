"""Data side of the hot path (SURVEY.md 8f N1 + N4): what feeds ``model(data)``.

Reference call sites replaced here:

    train.py:81-86    TUDataset(root, name, pre_transform=Indegree(), use_node_attr=True)
                      -> ``read_tu_dataset`` (PyG's raw TU text format) + ``indegree``
    utils.py:5-36     Indegree                       -> ``indegree``
    train.py:102-107  10fold_idx/{train,test}_idx-N.txt, ``data_set[idx]``  -> ``load_fold``
    train.py:108-109  DataLoader(batch_size, shuffle) -> ``epoch_batches``
    train.py:36       Batch.from_data_list + ``.to(device)``
                      -> ``DeviceDataset``: the data set lives in HBM as one canonical CSR
                      (K0 over all graphs, once); a batch is a list of graph ids and is
                      gathered on the device by ``dgcnn_collate`` (csrc/collate.cu).

Graphs are plain dicts ``{"x": f32 [n,F], "edge_index": i64 [2,e] (local ids), "y": int}``
-- the format ``synth.make_graphs`` also produces.  Host code is numpy only; everything
that touches the GPU goes through the C ABI (there is no CPU collate for the product path:
``synth.collate`` is the host loader's restatement used by the tests and the host-fed
bench leg).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from .synth import GraphBatch, collate

__all__ = ["read_tu_dataset", "write_tu_dataset", "indegree", "Indegree", "load_fold", "epoch_batches",
           "DeviceDataset", "ResidentBatch", "ResidentLoader"]


# ----------------------------------------------------------------------------------------
# TU raw format (what PyG's TUDataset downloads and parses; train.py:81-86)
# ----------------------------------------------------------------------------------------
def _read_table(path: str, dtype) -> Optional[np.ndarray]:
    if not os.path.exists(path):
        return None
    try:                                                    # numpy's C parser: COLLAB's _A.txt has 24.6 M lines
        return np.loadtxt(path, delimiter=",", dtype=np.float64, ndmin=2).astype(dtype)
    except ValueError:
        pass
    with open(path) as fh:
        rows = [line.replace(",", " ").split() for line in fh if line.strip()]
    if not rows:
        return np.zeros((0, 1), dtype=dtype)
    return np.asarray(rows, dtype=np.float64).astype(dtype)


def _one_hot_columns(labels: np.ndarray) -> np.ndarray:
    """PyG read_tu_data: every label column is shifted to start at 0 and one-hot encoded;
    the encodings are concatenated."""
    parts = []
    for c in range(labels.shape[1]):
        v = labels[:, c] - labels[:, c].min() if labels.shape[0] else labels[:, c]
        width = int(v.max()) + 1 if v.size else 0
        hot = np.zeros((v.size, width), dtype=np.float32)
        hot[np.arange(v.size), v] = 1.0
        parts.append(hot)
    return np.concatenate(parts, axis=1) if parts else np.zeros((labels.shape[0], 0), np.float32)


def indegree(x: Optional[np.ndarray], edge_index: np.ndarray, num_nodes: int, norm: bool = True,
             max_value: Optional[float] = None, cat: bool = True) -> np.ndarray:
    """utils.py:18-33 ``Indegree.__call__`` on one graph: float32 in-degree (count of
    ``edge_index[1]``), divided by the graph's own maximum (NaN when the graph has no edge:
    0/0, as in the reference) and appended as the LAST feature column."""
    deg = np.bincount(edge_index[1], minlength=num_nodes).astype(np.float32)      # utils.py:20
    if norm:
        top = np.float32(deg.max() if deg.size else 0.0) if max_value is None else np.float32(max_value)
        with np.errstate(invalid="ignore", divide="ignore"):
            deg = deg / top                                                      # utils.py:23
    deg = deg.reshape(-1, 1)
    if x is not None and cat:
        x = x.reshape(-1, 1) if x.ndim == 1 else x
        return np.concatenate([x, deg.astype(x.dtype)], axis=1)                  # utils.py:29
    return deg


class Indegree:
    """The reference's ``utils.Indegree`` transform object (utils.py:5-36), same constructor and
    call convention: ``Indegree(norm=True, max_value=None, cat=True)(data)`` appends (or, with
    ``cat=False``, substitutes) the normalised in-degree column on any object with
    ``edge_index`` / ``x`` / ``num_nodes`` and returns it.  A pre-transform: it runs on the host,
    once per graph, before the data set goes to HBM; the arithmetic is ``indegree`` above."""

    def __init__(self, norm: bool = True, max_value: Optional[float] = None, cat: bool = True):
        self.norm, self.max, self.cat = norm, max_value, cat

    def __call__(self, data):
        ei = data.edge_index
        ei_np = ei.detach().cpu().numpy() if isinstance(ei, torch.Tensor) else np.asarray(ei)
        x = getattr(data, "x", None)
        as_tensor = isinstance(x, torch.Tensor) or (x is None and isinstance(ei, torch.Tensor))
        x_np = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x
        n = int(data.num_nodes) if getattr(data, "num_nodes", None) is not None else int(x_np.shape[0])
        out = indegree(x_np, ei_np, n, self.norm, self.max, self.cat)
        data.x = torch.from_numpy(np.ascontiguousarray(out)) if as_tensor else out
        return data

    def __repr__(self):
        return f"{self.__class__.__name__}(norm={self.norm}, max_value={self.max})"


def read_tu_dataset(root: str, name: str, use_node_attr: bool = True,
                    pre_transform: bool = True) -> Tuple[List[dict], int, int]:
    """Parse ``{root}/{name}_A.txt`` & co. the way PyG's ``read_tu_data`` does and apply the
    reference's ``pre_transform=Indegree()`` per graph.  Returns (graphs, num_features,
    num_classes).  Looks in ``root`` and in ``root/raw`` and ``root/{name}/raw`` (PyG's layout
    under train.py:82's ``data/{name}``)."""
    for cand in (root, os.path.join(root, "raw"), os.path.join(root, name, "raw"), os.path.join(root, name)):
        if os.path.exists(os.path.join(cand, f"{name}_A.txt")):
            root = cand
            break
    else:
        raise FileNotFoundError(f"{name}_A.txt not found under {root} (TU raw files; the reference "
                                "downloads them through PyG, this build has no network)")
    pre = os.path.join(root, name)
    edges = _read_table(pre + "_A.txt", np.int64) - 1                  # 1-based (row, col) pairs
    indicator = _read_table(pre + "_graph_indicator.txt", np.int64).reshape(-1) - 1
    glabels = _read_table(pre + "_graph_labels.txt", np.int64)
    nlabels = _read_table(pre + "_node_labels.txt", np.int64)
    nattr = _read_table(pre + "_node_attributes.txt", np.float32) if use_node_attr else None
    num_nodes = indicator.size
    num_graphs = int(indicator.max()) + 1 if num_nodes else 0
    if np.any(np.diff(indicator) < 0):
        raise ValueError("graph_indicator must be non-decreasing")
    parts = []
    if nattr is not None:
        parts.append(nattr.reshape(num_nodes, -1).astype(np.float32))  # attributes first, then labels
    if nlabels is not None:
        parts.append(_one_hot_columns(nlabels.reshape(num_nodes, -1)))
    x_all = np.concatenate(parts, axis=1) if parts else None
    # graph labels -> 0..C-1 in sorted order (y.unique(sorted=True, return_inverse=True))
    classes, y_all = np.unique(glabels.reshape(-1), return_inverse=True)
    # remove_self_loops + coalesce: sorted by (row, col), duplicates merged
    edges = edges[edges[:, 0] != edges[:, 1]]
    key = np.unique(edges[:, 0] * max(num_nodes, 1) + edges[:, 1])
    row, col = key // max(num_nodes, 1), key % max(num_nodes, 1)
    if row.size and (indicator[row] != indicator[col]).any():
        raise ValueError("an edge connects two different graphs")
    ptr = np.searchsorted(indicator, np.arange(num_graphs + 1))
    eptr = np.searchsorted(row, ptr)
    graphs = []
    for g in range(num_graphs):
        lo, hi = int(ptr[g]), int(ptr[g + 1])
        ei = np.stack([row[eptr[g]:eptr[g + 1]] - lo, col[eptr[g]:eptr[g + 1]] - lo]).astype(np.int64)
        x = x_all[lo:hi] if x_all is not None else None
        if pre_transform:
            x = indegree(x, ei, hi - lo)
        elif x is None:
            x = np.ones((hi - lo, 1), dtype=np.float32)
        graphs.append({"x": np.ascontiguousarray(x, dtype=np.float32), "edge_index": ei, "y": int(y_all[g])})
    num_features = graphs[0]["x"].shape[1] if graphs else 0
    return graphs, num_features, int(classes.size)


def write_tu_dataset(root: str, name: str, graphs: Sequence[dict], node_labels: bool = True) -> None:
    """Inverse of ``read_tu_dataset`` for one-hot features (tests, synthetic stand-ins for the
    absent TU downloads): writes ``_A``, ``_graph_indicator``, ``_graph_labels`` and, when the
    graphs carry more than the in-degree column, ``_node_labels`` (argmax of x[:, :-1])."""
    os.makedirs(root, exist_ok=True)
    pre = os.path.join(root, name)
    off, a_lines, ind, nl = 0, [], [], []
    for g, gr in enumerate(graphs):
        ei, n = gr["edge_index"], gr["x"].shape[0]
        a_lines += [f"{s + off + 1}, {d + off + 1}" for s, d in zip(ei[0].tolist(), ei[1].tolist())]
        ind += [str(g + 1)] * n
        if node_labels and gr["x"].shape[1] > 1:
            nl += [str(int(v)) for v in np.argmax(gr["x"][:, :-1], axis=1)]
        off += n
    with open(pre + "_A.txt", "w") as fh:
        fh.write("\n".join(a_lines) + "\n")
    with open(pre + "_graph_indicator.txt", "w") as fh:
        fh.write("\n".join(ind) + "\n")
    with open(pre + "_graph_labels.txt", "w") as fh:
        fh.write("\n".join(str(int(gr["y"])) for gr in graphs) + "\n")
    if nl:
        with open(pre + "_node_labels.txt", "w") as fh:
            fh.write("\n".join(nl) + "\n")


def load_fold(root: str, data_type: str, fold_number: int) -> Tuple[np.ndarray, np.ndarray]:
    """train.py:102-105: the 0-based graph ids of ``10fold_idx/{train,test}_idx-N.txt``."""
    base = os.path.join(root, data_type, "10fold_idx")
    tr = np.loadtxt(os.path.join(base, f"train_idx-{fold_number}.txt"), dtype=np.int32)
    te = np.loadtxt(os.path.join(base, f"test_idx-{fold_number}.txt"), dtype=np.int32)
    return np.atleast_1d(tr).astype(np.int64), np.atleast_1d(te).astype(np.int64)


def epoch_batches(ids: np.ndarray, batch_size: int, shuffle: bool,
                  generator: Optional[torch.Generator] = None) -> Iterator[np.ndarray]:
    """train.py:108-109 ``DataLoader(dataset[ids], batch_size, shuffle)``: one pass over ``ids``
    in (optionally shuffled: ``torch.randperm``, as the DataLoader's RandomSampler draws it)
    order, last batch short (drop_last=False)."""
    ids = np.asarray(ids, dtype=np.int64)
    if shuffle:
        ids = ids[torch.randperm(ids.size, generator=generator).numpy()]
    for lo in range(0, ids.size, batch_size):
        yield ids[lo:lo + batch_size]


# ----------------------------------------------------------------------------------------
# resident data set
# ----------------------------------------------------------------------------------------
class ResidentBatch:
    """What ``DeviceDataset.batch`` returns: the attributes ``Model.forward`` reads
    (``x, batch, y, num_graphs, max_nodes``) plus the prebuilt ``Graph`` (no edge_index is
    materialised: the CSR is what every kernel consumes)."""

    def __init__(self, x, batch, y, gptr, graph: ops.Graph, ids):
        self.x, self.batch, self.y, self.ptr = x, batch, y, gptr
        self.edge_index = None
        self.num_graphs = graph.num_graphs
        self.max_nodes = graph.max_nodes
        self._dgcnn_graph = graph
        self.ids = ids

    @property
    def num_nodes(self) -> int:
        return int(self.x.size(0))

    def to(self, device, non_blocking: bool = False) -> "ResidentBatch":
        if torch.device(device).type != "cuda":
            raise RuntimeError("ResidentBatch lives on the GPU (there is no CPU path)")
        return self


class DeviceDataset:
    """All graphs of a data set in HBM, as the outputs of ONE K0 pass over the whole set.

    ``plan(ids)`` is host arithmetic on the per-graph sizes (no sync); ``batch(ids)`` gathers a
    batch for ``Model(data)`` (evaluation, train.py:60); ``FusedTrainer.step_resident`` runs a
    whole training step on ``ids`` (train.py:35-45)."""

    def __init__(self, graphs: Sequence[dict], device, num_classes: Optional[int] = None,
                 maps: bool = True):
        if not graphs:
            raise ValueError("DeviceDataset needs at least one graph")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DeviceDataset: the data set must live on a CUDA device")
        for g, gr in enumerate(graphs):
            n, ei = gr["x"].shape[0], gr["edge_index"]
            if ei.size and (ei.min() < 0 or ei.max() >= n):
                raise ValueError(f"graph {g}: edge_index outside [0, {n})")
        host = collate(graphs)                        # one batch of everything (host, once)
        self.device = dev
        self.num_graphs = len(graphs)
        self.num_features = int(host.x.size(1))
        self.num_classes = int(num_classes if num_classes is not None else int(host.y.max()) + 1)
        self.nodes = np.diff(host.ptr.numpy()).astype(np.int64)         # per-graph node counts (host)
        self.x = host.x.to(dev).contiguous()
        self.y = host.y.to(dev).contiguous()
        n = int(host.x.size(0))
        graph = ops.build_graph(host.edge_index.to(dev), host.batch.to(dev), n, self.num_graphs,
                                transpose=True, max_nodes=0)
        graph.check()                                 # set-up: one sync is fine here
        self.symmetric = not (int(graph.status.item()) & ops.GRAPH_GENERIC)
        self.gptr, self.rowptr, self.dis = graph.gptr, graph.rowptr, graph.dis
        self.num_nodes = n
        self.num_edges = int(graph.rowptr[-1].item())                    # loops dropped by K0
        self.col = graph.col[:max(self.num_edges, 1)].contiguous()
        self.rowptr_t = None if self.symmetric else graph.rowptr_t
        self.col_t = None if self.symmetric else graph.col_t[:max(self.num_edges, 1)].contiguous()
        first = self.rowptr[self.gptr.long()]
        self.edges = (first[1:] - first[:-1]).cpu().numpy().astype(np.int64)   # per-graph edge counts
        # K0b over the whole data set, once: per-graph bitmaps / fragment maps are relative to the
        # graph's first node, so a batch's maps are copies of these blocks
        self.maps = None
        if maps and int(self.nodes.max()) > 0:
            graph.max_nodes = int(min(int(self.nodes.max()), ops.BITMAP_MAX_NODES))
            graph.gorder = None                                          # descriptors are per batch
            ops._build_bitmaps(graph, not self.symmetric, host.batch.to(dev))
            self.maps = graph
        m = self.maps
        self._struct = _lib.DgcnnDataset(
            self.num_graphs, self.num_nodes, self.num_edges, self.num_features, int(self.symmetric),
            self.x.data_ptr(), int(self.x.stride(0)) if n > 1 else self.num_features, self.y.data_ptr(),
            self.gptr.data_ptr(), self.rowptr.data_ptr(), self.col.data_ptr(),
            None if self.symmetric else self.rowptr_t.data_ptr(),
            None if self.symmetric else self.col_t.data_ptr(), self.dis.data_ptr(),
            None if m is None else m.bitmap.data_ptr(),
            None if m is None or self.symmetric else m.bitmap_t.data_ptr(),
            None if m is None else m.bmoff.data_ptr(), None if m is None else m.gflags.data_ptr(),
            None if m is None or self.symmetric else m.gflags_t.data_ptr(),
            None if m is None else m.fragmap.data_ptr(), None if m is None else m.fgoff.data_ptr(), None)
        # one validating pass (offsets closed, CSRs consistent, no edge leaves its graph) that also
        # writes the per-graph extent records the gather starts from
        self.gext = torch.empty(self.num_graphs, 4, dtype=torch.int32, device=dev)
        check = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load_library().dgcnn_dataset_prepare(self.c_struct, self.gext.data_ptr(), check.data_ptr(),
                                                           torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "dataset_prepare")
        if int(check.item()) != 0:
            raise ValueError(f"DeviceDataset: inconsistent data set (status {int(check.item())}): an edge "
                             "leaves its graph or the offsets are not monotone")
        self._struct.gext = self.gext.data_ptr()

    def __len__(self) -> int:
        return self.num_graphs

    @property
    def c_struct(self):
        return ctypes.byref(self._struct)

    def nbytes(self) -> int:
        ts = [self.x, self.y, self.gptr, self.rowptr, self.col, self.dis, self.rowptr_t, self.col_t, self.gext]
        if self.maps is not None:
            m = self.maps
            ts += [m.bitmap, m.bitmap_t, m.bmoff, m.gflags, m.gflags_t, m.fragmap, m.fgoff]
        return sum(t.numel() * t.element_size() for t in ts if t is not None)

    def plan(self, ids) -> Tuple[int, int, int]:
        """(nodes, edges, largest graph) of the batch ``ids``: what sizes the step's buffers."""
        ids = np.asarray(ids, dtype=np.int64)
        if ids.size == 0 or ids.min() < 0 or ids.max() >= self.num_graphs:
            raise IndexError("DeviceDataset: graph id outside the data set (or empty batch)")
        nn = self.nodes[ids]
        return int(nn.sum()), int(self.edges[ids].sum()), int(nn.max())

    def shard(self, ids, world_size: int, rank: int) -> np.ndarray:
        """This rank's slice of the batch ``ids`` (graph-sharded data parallelism;
        every rank holds the whole data set, the split is balanced by nodes + edges)."""
        from .dp import shard_ids
        return shard_ids(np.asarray(ids, dtype=np.int64), self.nodes, self.edges, world_size, rank)

    def ids_to_device_pinned(self, ids) -> torch.Tensor:
        """int32 copy of ``ids`` on the device through a pinned staging buffer that is allocated once
        and reused (a whole epoch's shuffled ids in one transfer).  The caller must have synchronised
        with the previous use of the returned tensor (driver.train_epoch reads the epoch's statistics
        before the next epoch starts)."""
        count = int(len(ids))
        if getattr(self, "_ids_pinned", None) is None or self._ids_pinned.numel() < count:
            self._ids_pinned = torch.empty(max(count, 1024), dtype=torch.int32).pin_memory()
            self._ids_device = torch.empty(self._ids_pinned.numel(), dtype=torch.int32, device=self.device)
        self._ids_pinned.numpy()[:count] = np.asarray(ids)   # (casts to int32 in place)
        self._ids_device[:count].copy_(self._ids_pinned[:count], non_blocking=True)
        return self._ids_device[:count]

    def ids_to_device(self, ids) -> torch.Tensor:
        t = ids if isinstance(ids, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int32))
        return t.to(dtype=torch.int32).to(self.device, non_blocking=True)

    def batch(self, ids, ids_device: Optional[torch.Tensor] = None, need_x: bool = True,
              bitmaps: bool = True) -> ResidentBatch:
        """Gather the batch ``ids`` on the device (dgcnn_collate + K0b) for ``Model(data)``."""
        lib = _lib.load_library()
        n, e, mx = self.plan(ids)
        b = int(len(ids))
        dev = self.device
        if ids_device is None:
            ids_device = self.ids_to_device(ids)
        i32 = dict(dtype=torch.int32, device=dev)
        x = torch.empty(n, self.num_features, dtype=torch.float32, device=dev) if need_x else None
        batch32 = torch.empty(n, **i32)
        y = torch.empty(b, dtype=torch.int64, device=dev)
        rowptr, col = torch.empty(n + 1, **i32), torch.empty(max(e, 1), **i32)
        if self.symmetric:
            rowptr_t, col_t = rowptr, col
        else:
            rowptr_t, col_t = torch.empty(n + 1, **i32), torch.empty(max(e, 1), **i32)
        dis = torch.empty(n, dtype=torch.float32, device=dev)
        gptr, gorder = torch.empty(b + 1, **i32), torch.empty(b, **i32)
        status = torch.zeros(1, **i32)
        ws = torch.empty(int(lib.dgcnn_collate_workspace_bytes(b)), dtype=torch.uint8, device=dev)
        graph = ops.Graph(rowptr, col, rowptr_t, col_t, dis, gptr, gorder, status, n, b, mx)
        want_maps = bitmaps and 0 < mx <= ops.BITMAP_MAX_NODES
        gathered = want_maps and self.maps is not None
        p = lambda t: None if t is None else t.data_ptr()
        if gathered:                                   # K0b's outputs come from the data set's cache
            words = int(lib.dgcnn_graph_bitmap_words(n, b, mx))
            fwords = int(lib.dgcnn_graph_fragmap_words(n, b, mx))
            both = torch.empty((1 if self.symmetric else 2) * words, **i32)
            graph.bitmap = both[:words]
            graph.bitmap_t = None if self.symmetric else both[words:]
            graph.bmoff, graph.gflags = torch.empty(b + 1, **i32), torch.empty(b, **i32)
            graph.bmoff_t = None if self.symmetric else graph.bmoff
            graph.gflags_t = None if self.symmetric else torch.empty(b, **i32)
            graph.fragmap, graph.fgoff = torch.empty(fwords, **i32), torch.empty(b + 1, **i32)
            graph.gdesc = torch.empty(b, 4, **i32)
        out = _lib.DgcnnBatchGraph(
            p(x), self.num_features, p(batch32), p(y), p(rowptr), p(col), p(rowptr_t), p(col_t), p(dis),
            p(gptr), p(gorder), p(graph.bitmap), p(graph.bitmap_t), p(graph.bmoff), p(graph.gflags),
            p(graph.gflags_t), p(graph.fragmap), p(graph.fgoff), p(graph.gdesc))
        with torch.cuda.device(dev):
            rc = lib.dgcnn_collate(self.c_struct, ids_device.data_ptr(), b, n, e, ctypes.byref(out),
                                   status.data_ptr(), ws.data_ptr(), ws.numel(),
                                   torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "collate")
        ops.LAUNCHES["collate"] = ops.LAUNCHES.get("collate", 0) + (1 if b <= 1024 else 2)
        if want_maps and not gathered:
            ops._build_bitmaps(graph, True, batch32)
        return ResidentBatch(x, batch32, y, gptr, graph, ids)

    def host_batch(self, graphs: Sequence[dict], ids) -> GraphBatch:
        """The host loader's batch of the same ids (tests / the host-fed comparison)."""
        b = collate([graphs[int(i)] for i in ids])
        b.max_nodes = int(self.nodes[np.asarray(ids, dtype=np.int64)].max())
        return b


class ResidentLoader:
    """``DataLoader(dataset=data_set[ids], batch_size=..., shuffle=...)`` of train.py:108-109 over a
    resident data set: iterating yields gathered batches (``ResidentBatch``: ``.x .batch .y
    .num_graphs``, ``.to(device)`` is a no-op) in the DataLoader's order, so the loops of
    train.py:35-45 / 57-64 (``for sample in dataloader: data, y = sample.to(device),
    sample.y.to(device); pred = model(data) ...``) run as they are.  ``len(loader)`` = number of
    batches, ``len(loader.dataset)`` = number of graphs, as the reference's averages need."""

    def __init__(self, dataset: "DeviceDataset", ids, batch_size: int = 1, shuffle: bool = False,
                 generator: Optional[torch.Generator] = None):
        self.source = dataset
        self.dataset = np.asarray(ids, dtype=np.int64)          # what len(dataloader.dataset) counts
        self.batch_size, self.shuffle, self.generator = int(batch_size), bool(shuffle), generator
        if self.batch_size < 1:
            raise ValueError("batch_size must be positive")

    def __len__(self) -> int:
        return -(-len(self.dataset) // self.batch_size)

    def __iter__(self):
        for ids in epoch_batches(self.dataset, self.batch_size, self.shuffle, self.generator):
            yield self.source.batch(ids)
