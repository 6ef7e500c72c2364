"""Synthetic TU-shaped graph batches and the batch container.

The reference reads real TU datasets through PyG (train.py:81-86) and collates
them with PyG's DataLoader (train.py:108-109).  Neither the datasets nor PyG
exist here, so the benchmark / parity inputs are synthetic graphs with the same
layout the hot path consumes (model.py:27):

    x          f32 [N, F]   one-hot(label over F-1 classes) || Indegree column
                            (README.md:44-45, utils.py:18-33)
    edge_index i64 [2, E]   row 0 = source, row 1 = target; symmetric, no loops,
                            no duplicates, graph-contiguous, sorted by (src,dst)
    batch      i64 [N]      graph id per node, non-decreasing
    ptr        i64 [B+1]    node offsets per graph
    y          i64 [B]

Everything is drawn from ``numpy.random.RandomState(seed)`` (seed 324 is the
reference's default, train.py:24), so the same batch is rebuilt bit-identically
on any host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

__all__ = ["GraphBatch", "CONFIGS", "make_graphs", "collate", "make_batch"]


class GraphBatch:
    """Minimal stand-in for PyG's ``Batch``: the attributes model.py:27 and
    train.py:36 touch (``x, edge_index, batch, y, ptr, num_graphs, to()``)."""

    def __init__(self, x, edge_index, batch, ptr, y, num_graphs: Optional[int] = None):
        self.x, self.edge_index, self.batch, self.ptr, self.y = x, edge_index, batch, ptr, y
        self.num_graphs = int(num_graphs if num_graphs is not None else ptr.numel() - 1)

    @property
    def num_nodes(self) -> int:
        return int(self.x.size(0))

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.size(1))

    def _map(self, fn) -> "GraphBatch":
        return GraphBatch(fn(self.x), fn(self.edge_index), fn(self.batch), fn(self.ptr),
                          fn(self.y), self.num_graphs)

    def to(self, device, non_blocking: bool = False) -> "GraphBatch":
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self) -> "GraphBatch":
        return self._map(lambda t: t.pin_memory())

    def compact(self) -> "GraphBatch":
        """The same batch with int32 indices (edge_index, batch, ptr): what a loader that
        collates int32 would hand over; halves the host-to-device copy of train.py:36
        (dgcnn_build_graph_i32 takes it as is).  Labels stay int64."""
        i32 = lambda t: t.to(torch.int32)
        return GraphBatch(self.x, i32(self.edge_index), i32(self.batch), i32(self.ptr), self.y,
                          self.num_graphs)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size()
                   for t in (self.x, self.edge_index, self.batch, self.ptr, self.y))

    def __repr__(self):
        return (f"GraphBatch(graphs={self.num_graphs}, nodes={self.num_nodes}, "
                f"edges={self.num_edges}, F={self.x.size(1)})")


@dataclass(frozen=True)
class SynthConfig:
    name: str
    num_graphs: int        # graphs per batch (cfg1: the whole 188-graph data set)
    num_features: int
    num_classes: int
    k: int
    batch_size: int
    kind: str              # gnm | collab | powerlaw
    mean_nodes: float = 0.0
    sigma: float = 0.0
    min_nodes: int = 0
    max_nodes: int = 0
    mean_degree: float = 0.0
    force_max: bool = False


CONFIGS: Dict[str, SynthConfig] = {
    # BASELINE.json configs[0..4]
    "mutag": SynthConfig("mutag", 188, 8, 2, 30, 50, "gnm_uniform", min_nodes=10, max_nodes=28,
                         mean_degree=2.2),
    "proteins": SynthConfig("proteins", 128, 5, 2, 60, 128, "gnm", mean_nodes=39.0, sigma=0.9,
                            min_nodes=4, max_nodes=620, mean_degree=3.73),
    "dd": SynthConfig("dd", 64, 90, 2, 291, 64, "gnm", mean_nodes=284.0, sigma=0.8,
                      min_nodes=30, max_nodes=5748, mean_degree=5.03, force_max=True),
    "collab": SynthConfig("collab", 512, 1, 3, 130, 512, "collab", mean_nodes=74.5, sigma=0.5,
                          min_nodes=32, max_nodes=492),
    "powerlaw": SynthConfig("powerlaw", 256, 64, 2, 512, 256, "powerlaw", mean_nodes=1000,
                            mean_degree=10),
}


def _lognormal_sizes(rng, cfg: SynthConfig, count: int) -> np.ndarray:
    mu = np.log(cfg.mean_nodes) - 0.5 * cfg.sigma ** 2      # so that E[n] = mean_nodes
    n = np.rint(rng.lognormal(mu, cfg.sigma, size=count)).astype(np.int64)
    n = np.clip(n, cfg.min_nodes, cfg.max_nodes)
    if cfg.force_max and count > 0:
        n[int(np.argmax(n))] = cfg.max_nodes                # keep the heavy tail (D&D max 5748)
    return n


def _gnm_pairs(rng, n: int, m: int) -> np.ndarray:
    """m distinct undirected pairs (i<j) of an n-node graph, [m,2]."""
    total = n * (n - 1) // 2
    m = int(min(m, total))
    if m <= 0:
        return np.zeros((0, 2), dtype=np.int64)
    if m * 4 >= total:                                       # dense: sample pair ids exactly
        iu, ju = np.triu_indices(n, 1)
        pick = rng.choice(total, size=m, replace=False)
        return np.stack([iu[pick], ju[pick]], 1).astype(np.int64)
    seen = np.zeros(0, dtype=np.int64)
    while seen.size < m:                                     # sparse: rejection
        a = rng.randint(0, n, size=2 * (m - seen.size) + 8)
        b = rng.randint(0, n, size=a.size)
        ok = a != b
        lo, hi = np.minimum(a, b)[ok], np.maximum(a, b)[ok]
        seen = np.concatenate([seen, lo * n + hi])
        _, first = np.unique(seen, return_index=True)
        seen = seen[np.sort(first)]                          # dedup, first occurrence order
    seen = seen[:m]
    return np.stack([seen // n, seen % n], 1)


def _powerlaw_pairs(rng, n: int, m: int) -> np.ndarray:
    """Barabasi-Albert preferential attachment: (n-m)*m undirected edges."""
    src, dst = [], []
    targets = np.arange(m)
    repeated = np.zeros(2 * (n - m) * m, dtype=np.int64)
    fill = 0
    for v in range(m, n):
        src.append(np.full(m, v, dtype=np.int64))
        dst.append(targets.astype(np.int64))
        repeated[fill:fill + m] = targets
        repeated[fill + m:fill + 2 * m] = v
        fill += 2 * m
        chosen = set()
        while len(chosen) < m:                               # m distinct, degree-proportional
            for c in repeated[rng.randint(0, fill, size=m - len(chosen))]:
                chosen.add(int(c))
        targets = np.fromiter(chosen, dtype=np.int64, count=m)
    a, b = np.concatenate(src), np.concatenate(dst)
    return np.stack([np.minimum(a, b), np.maximum(a, b)], 1)


def _symmetrise_sorted(pairs: np.ndarray, n: int) -> np.ndarray:
    src = np.concatenate([pairs[:, 0], pairs[:, 1]])
    dst = np.concatenate([pairs[:, 1], pairs[:, 0]])
    order = np.argsort(src * n + dst, kind="stable")
    return np.stack([src[order], dst[order]])


def make_graphs(cfg: SynthConfig, count: Optional[int] = None, seed: int = 324,
                tie_free: bool = False) -> List[dict]:
    """``count`` independent graphs: dicts of numpy ``x [n,F] f32``,
    ``edge_index [2,e] i64`` (local ids), ``y`` int."""
    count = cfg.num_graphs if count is None else count
    rng = np.random.RandomState(seed)
    if cfg.kind == "gnm_uniform":
        sizes = rng.randint(cfg.min_nodes, cfg.max_nodes + 1, size=count)
    elif cfg.kind == "powerlaw":
        sizes = np.full(count, int(cfg.mean_nodes), dtype=np.int64)
    else:
        sizes = _lognormal_sizes(rng, cfg, count)
    graphs = []
    for g in range(count):
        n = int(sizes[g])
        if cfg.kind == "collab":
            pairs = _gnm_pairs(rng, n, min(n * (n - 1) // 2, 33 * n))
        elif cfg.kind == "powerlaw":
            pairs = _powerlaw_pairs(rng, n, int(cfg.mean_degree))
        else:
            pairs = _gnm_pairs(rng, n, int(round(cfg.mean_degree * n / 2.0)))
        ei = _symmetrise_sorted(pairs, n)
        f = cfg.num_features
        if tie_free:
            x = rng.standard_normal((n, f)).astype(np.float32)
        else:
            x = np.zeros((n, f), dtype=np.float32)
            if f > 1:
                x[np.arange(n), rng.randint(0, f - 1, size=n)] = 1.0
            deg = np.bincount(ei[1], minlength=n).astype(np.float32)     # utils.py:20
            x[:, f - 1] = deg / deg.max() if deg.max() > 0 else np.float32("nan")
        graphs.append({"x": x, "edge_index": ei, "y": int(rng.randint(0, cfg.num_classes))})
    return graphs


def collate(graphs: Sequence[dict]) -> GraphBatch:
    """What PyG's ``Batch.from_data_list`` does for the fields the hot path reads:
    concatenate, shift edge ids by the node offset, emit ``batch`` and ``ptr``."""
    sizes = np.array([g["x"].shape[0] for g in graphs], dtype=np.int64)
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    f = graphs[0]["x"].shape[1] if graphs else 0
    x = np.concatenate([g["x"] for g in graphs]) if graphs else np.zeros((0, f), np.float32)
    ei = (np.concatenate([g["edge_index"] + ptr[i] for i, g in enumerate(graphs)], axis=1)
          if graphs else np.zeros((2, 0), np.int64))
    batch = np.repeat(np.arange(len(graphs), dtype=np.int64), sizes)
    y = np.array([g["y"] for g in graphs], dtype=np.int64)
    return GraphBatch(torch.from_numpy(x), torch.from_numpy(np.ascontiguousarray(ei)),
                      torch.from_numpy(batch), torch.from_numpy(ptr.astype(np.int64)),
                      torch.from_numpy(y), len(graphs))


def make_batch(name: str, seed: int = 324, num_graphs: Optional[int] = None,
               tie_free: bool = False) -> GraphBatch:
    """One batch of BASELINE.json's configuration ``name`` (``batch_size`` graphs)."""
    cfg = CONFIGS[name]
    count = cfg.batch_size if num_graphs is None else num_graphs
    return collate(make_graphs(cfg, count, seed, tie_free))
