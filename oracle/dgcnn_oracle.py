"""CPU oracle for the DGCNN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
package ``dgcnn_b200`` never does: its operators fail loudly without the CUDA
library.

PARITY UNPINNED.  The arithmetic of the reference's hot path lives in a
third-party dependency, PyTorch Geometric (``torch_geometric``, version
unpinned by the reference: no requirements file; ``SortAggregation`` implies
>= 2.1.0), which is neither vendored under /root/reference nor installable in
this image (no network, not in /opt/wheelhouse), and the reference ships no
tests or golden vectors.  This file therefore restates the *published* PyG
algorithms at the reference's own call sites:

    model.py:28      remove_self_loops            -> remove_self_loops()
    model.py:13-16   GCNConv ctor  (PyG nn/conv/gcn_conv.py)
    model.py:30-33   GCNConv.forward + torch.tanh -> gcn_norm(), gcn_conv()
    model.py:34      torch.cat                     -> graph_conv_stack()
    model.py:17,35   SortAggregation(k).forward (PyG nn/aggr/sort.py,
                     utils/to_dense_batch.py)      -> sort_aggregation()
    model.py:36-43   dense tail                    -> OracleModel.forward
    utils.py:18-33   Indegree transform            -> indegree_feature()

The only numeric anchors the reference offers are the README.md:96-104
parameter counts; ``tests/test_oracle.py`` pins those, plus closed-form dense
``D^-1/2 (A+I) D^-1/2`` evaluations in float64 and hand-computed fixtures.

Everything here is plain torch on CPU, the same ATen op sequence PyG's pure
torch fallback issues (index_select -> mul -> scatter_add_; to_dense_batch ->
sort -> gather -> pad), which is also why it doubles as the CPU baseline.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

NORM_SYM = 0  # D^-1/2 (A+I) D^-1/2  -- what GCNConv (hence the reference) computes
NORM_RW = 1   # D^-1 (A+I)           -- the AAAI-18 paper's formula (north_star D1)


# ----------------------------------------------------------------------------
# model.py:28  (PyG utils/loop.py remove_self_loops)
# ----------------------------------------------------------------------------
def remove_self_loops(edge_index: Tensor) -> Tuple[Tensor, None]:
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], None


# ----------------------------------------------------------------------------
# model.py:30-33 -> GCNConv.forward -> gcn_norm  (PyG nn/conv/gcn_conv.py)
# improved=False, add_self_loops=True, flow='source_to_target', edge_weight=None
# ----------------------------------------------------------------------------
def gcn_norm(edge_index: Tensor, num_nodes: int, dtype=torch.float32,
             norm: int = NORM_SYM) -> Tuple[Tensor, Tensor]:
    """Returns (edge_index' [2,E+N], weight [E+N]).

    add_remaining_self_loops: existing loops are dropped, the surviving edges
    keep their order, one loop per node (weight 1) is appended AFTER them.
    deg[i] counts incoming edges at the TARGET i (multi-edges counted) + 1.
    """
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    loops = torch.arange(num_nodes, dtype=edge_index.dtype)
    src = torch.cat([src[keep], loops])
    dst = torch.cat([dst[keep], loops])
    w = torch.ones(src.numel(), dtype=dtype)
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, dst, w)
    if norm == NORM_SYM:
        dis = deg.pow(-0.5)
        dis = dis.masked_fill(dis == float("inf"), 0.0)
        w = dis[src] * w * dis[dst]
    elif norm == NORM_RW:
        dinv = deg.pow(-1.0)
        dinv = dinv.masked_fill(dinv == float("inf"), 0.0)
        w = w * dinv[dst]
    else:
        raise ValueError(f"unknown norm {norm}")
    return torch.stack([src, dst]), w


def gcn_conv(x: Tensor, edge_index: Tensor, weight: Tensor, bias: Optional[Tensor],
             norm: int = NORM_SYM) -> Tensor:
    """GCNConv.forward: lin (no bias) -> propagate (gather, scale, scatter-add at
    target) -> + bias.  weight is [Cout, Cin] like PyG's ``lin.weight``."""
    n = x.size(0)
    ei, w = gcn_norm(edge_index, n, x.dtype, norm)
    h = x @ weight.t()
    msg = w.unsqueeze(-1) * h.index_select(0, ei[0])
    out = torch.zeros(n, weight.size(0), dtype=x.dtype)
    out.scatter_add_(0, ei[1].unsqueeze(-1).expand_as(msg), msg)
    if bias is not None:
        out = out + bias
    return out


def graph_conv_stack(x: Tensor, edge_index: Tensor, weights, biases,
                     norm: int = NORM_SYM) -> Tensor:
    """model.py:28-34: remove_self_loops, 4x tanh(GCNConv), channel concat."""
    edge_index, _ = remove_self_loops(edge_index)
    outs = []
    h = x
    for w, b in zip(weights, biases):
        h = torch.tanh(gcn_conv(h, edge_index, w, b, norm))
        outs.append(h)
    return torch.cat(outs, dim=-1)


# ----------------------------------------------------------------------------
# model.py:35 -> SortAggregation.forward (PyG nn/aggr/sort.py) on top of
# to_dense_batch (PyG utils/to_dense_batch.py)
# ----------------------------------------------------------------------------
def to_dense_batch(x: Tensor, batch: Tensor, fill_value, batch_size: Optional[int] = None
                   ) -> Tuple[Tensor, Tensor, int]:
    if batch_size is None:
        batch_size = int(batch.max()) + 1 if batch.numel() else 0
    counts = torch.zeros(batch_size, dtype=torch.long).scatter_add_(
        0, batch, torch.ones_like(batch))
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    nmax = int(counts.max()) if batch_size else 0
    pos = torch.arange(batch.numel()) - ptr[batch]          # position inside own graph
    flat = batch * nmax + pos
    dense = x.new_full((batch_size * nmax, x.size(-1)), fill_value)
    dense[flat] = x
    return dense.view(batch_size, nmax, x.size(-1)), ptr, nmax


def sort_aggregation(x: Tensor, batch: Tensor, k: int, batch_size: Optional[int] = None,
                     return_perm: bool = False):
    """SortAggregation(k)(x, batch) -> [B, k*D].

    Sort key = LAST channel, descending.  The reference leaves ``stable``
    unset, so its tie order is implementation-defined; the build pins it to
    ``stable=True`` (ties by ascending node index, -0.0 == +0.0, NaN first).
    ``perm`` (int64 [B,k]) holds the GLOBAL node index feeding each output row,
    or -1 for zero-padded rows -- the contract the CUDA kernel is bit-exact to.
    """
    d = x.size(-1)
    fill = x.detach().min() - 1 if x.numel() else x.new_zeros(())
    dense, ptr, nmax = to_dense_batch(x, batch, fill, batch_size)
    b = dense.size(0)
    _, order = dense[:, :, -1].sort(dim=-1, descending=True, stable=True)
    gathered = dense.view(b * nmax, d)[(order + torch.arange(b).view(-1, 1) * nmax).view(-1)]
    gathered = gathered.view(b, nmax, d)
    if nmax >= k:
        gathered = gathered[:, :k].contiguous()
    else:
        pad = gathered.new_full((b, k - nmax, d), fill)
        gathered = torch.cat([gathered, pad], dim=1)
    gathered[gathered == fill] = 0
    out = gathered.view(b, k * d)
    if not return_perm:
        return out
    counts = ptr[1:] - ptr[:-1]
    r = torch.arange(k).view(1, -1)
    valid = r < counts.view(-1, 1).clamp(max=k)
    ordk = order[:, :k] if nmax >= k else torch.cat(
        [order, order.new_zeros(b, k - nmax)], dim=1)
    perm = torch.where(valid, ordk + ptr[:-1].view(-1, 1), torch.full_like(ordk, -1))
    return out, perm


# ----------------------------------------------------------------------------
# train.py:108-109 -> DataLoader -> Batch.from_data_list (PyG data/batch.py + collate.py)
# ----------------------------------------------------------------------------
def from_data_list(graphs):
    """What PyG's collate does for the attributes model.py:27 / train.py:36 read: ``x`` and
    ``y`` concatenated along dim 0, ``edge_index`` concatenated along dim 1 with every graph's
    ids incremented by the number of nodes before it (``__inc__`` = num_nodes), ``batch`` =
    graph id repeated per node, ``ptr`` = cumulative node counts.  ``graphs``: sequence of
    (x [n,F], edge_index [2,e] local ids, y) -> (x, edge_index, batch, ptr, y)."""
    xs, eis, ys, batch, ptr = [], [], [], [], [0]
    for g, (x, ei, y) in enumerate(graphs):
        x, ei = torch.as_tensor(x), torch.as_tensor(ei, dtype=torch.long)
        xs.append(x)
        eis.append(ei + ptr[-1])
        ys.append(int(y))
        batch.append(torch.full((x.size(0),), g, dtype=torch.long))
        ptr.append(ptr[-1] + x.size(0))
    return (torch.cat(xs, 0), torch.cat(eis, 1), torch.cat(batch), torch.tensor(ptr, dtype=torch.long),
            torch.tensor(ys, dtype=torch.long))


def batch_csr(edge_index: Tensor, num_nodes: int):
    """The graph structure every kernel consumes, from the reference's edge list: self loops
    dropped (model.py:28), CSR by target with ascending sources, CSR by source with ascending
    targets, dis = (1 + in-degree)^-1/2 (gcn_norm with the appended self loop), multi-edges kept."""
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    src, dst = src[keep], dst[keep]
    by_t = torch.argsort(dst * num_nodes + src, stable=True)
    by_s = torch.argsort(src * num_nodes + dst, stable=True)
    indeg = torch.bincount(dst, minlength=num_nodes)
    outdeg = torch.bincount(src, minlength=num_nodes)
    rowptr = torch.cat([indeg.new_zeros(1), indeg.cumsum(0)])
    rowptr_t = torch.cat([outdeg.new_zeros(1), outdeg.cumsum(0)])
    dis = (indeg + 1).to(torch.float32).pow(-0.5)
    return rowptr, src[by_t], rowptr_t, dst[by_s], dis


# ----------------------------------------------------------------------------
# utils.py:18-33  Indegree (norm=True, max_value=None, cat=True)
# ----------------------------------------------------------------------------
def indegree_feature(edge_index: Tensor, num_nodes: int, x: Optional[Tensor]) -> Tensor:
    """Per-graph transform: in-degree (float32 scatter of ones over edge_index[1])
    divided by the graph's max in-degree, appended as the last feature column.
    A graph without edges yields NaN (0/0), like the reference."""
    deg = torch.zeros(num_nodes, dtype=torch.float32).scatter_add_(
        0, edge_index[1], torch.ones(edge_index.size(1), dtype=torch.float32))
    deg = (deg / deg.max()).view(-1, 1)
    if x is None:
        return deg
    x = x.view(-1, 1) if x.dim() == 1 else x
    return torch.cat([x, deg.to(x.dtype)], dim=-1)


# ----------------------------------------------------------------------------
# model.py:9-45  Model, with k made a parameter and PyG's
# parameter names (conv{1-4}.lin.weight [Cout,Cin], conv{1-4}.bias)
# ----------------------------------------------------------------------------
class _OracleGCN(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.lin = nn.Linear(cin, cout, bias=False)
        self.bias = nn.Parameter(torch.zeros(cout))
        a = math.sqrt(6.0 / (cin + cout))           # PyG glorot()
        with torch.no_grad():
            self.lin.weight.uniform_(-a, a)

    def forward(self, x, edge_index, norm=NORM_SYM):
        return gcn_conv(x, edge_index, self.lin.weight, self.bias, norm)


def classifier_in_features(k: int) -> int:
    return 32 * (k // 2 - 4)


class OracleModel(nn.Module):
    def __init__(self, num_features: int, num_classes: int, k: int = 30, norm: int = NORM_SYM):
        super().__init__()
        self.k, self.norm = k, norm
        self.conv1 = _OracleGCN(num_features, 32)
        self.conv2 = _OracleGCN(32, 32)
        self.conv3 = _OracleGCN(32, 32)
        self.conv4 = _OracleGCN(32, 1)
        self.conv5 = nn.Conv1d(1, 16, 97, 97)
        self.conv6 = nn.Conv1d(16, 32, 5, 1)
        self.pool = nn.MaxPool1d(2, 2)
        self.classifier_1 = nn.Linear(classifier_in_features(k), 128)
        self.drop_out = nn.Dropout(0.5)
        self.classifier_2 = nn.Linear(128, num_classes)

    def hot_path(self, x, edge_index, batch, batch_size=None):
        """model.py:27-35 -> (x_cat [N,97], pooled [B,k*97])."""
        convs = (self.conv1, self.conv2, self.conv3, self.conv4)
        x_cat = graph_conv_stack(x, edge_index, [c.lin.weight for c in convs],
                                 [c.bias for c in convs], self.norm)
        return x_cat, sort_aggregation(x_cat, batch, self.k, batch_size)

    def tail(self, pooled):
        """model.py:36-43."""
        h = pooled.view(pooled.size(0), 1, pooled.size(-1))
        h = self.pool(F.relu(self.conv5(h)))
        h = F.relu(self.conv6(h)).flatten(1)
        h = self.drop_out(F.relu(self.classifier_1(h)))
        return F.log_softmax(self.classifier_2(h), dim=-1)

    def forward(self, data):
        bs = getattr(data, "num_graphs", None)
        _, pooled = self.hot_path(data.x, data.edge_index, data.batch, bs)
        return self.tail(pooled)
