"""Graph-sharded data parallelism: one process per GPU, every rank
runs the hot path on its own slice of the batch's graphs with replicated
parameters, and ONE all-reduce on a flat fp32 gradient bucket closes the step.

The reference has no distributed code (single device, train.py:75-79); this is
the one collective the build adds.  Graphs are independent block-diagonal
components in both GraphConv and SortPool (model.py:30-35 never mixes graphs),
so no data-path collective exists -- only the gradient sum.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_ids", "balanced_shards", "GradBucket", "PeerExchange"]


def shard_bounds(costs: Sequence[float], world_size: int) -> List[Tuple[int, int]]:
    """Split graphs ``0..len(costs)`` into ``world_size`` CONTIGUOUS slices of nearly equal
    total cost (cost ~ nodes + edges of a graph; counts alone skew badly on D&D-like size
    tails).  Every rank gets at least one graph whenever there are at least ``world_size``
    graphs; with fewer, the LAST ranks take them and the first slices are empty (such a rank
    still joins the gradient exchange: ``FusedTrainer.step_autograd(None, global_batch)``).
    Returns [(lo, hi)] per rank."""
    n = len(costs)
    total = float(sum(costs))
    bounds, lo, acc = [], 0, 0.0
    for r in range(world_size):
        later = world_size - r - 1                        # ranks still to be served after this one
        if r == world_size - 1:
            hi = n
        elif n < world_size:
            hi = lo + (1 if n - lo > later else 0)        # the last `n` ranks take one graph each
        else:
            target = total * (r + 1) / world_size
            cap = n - later                               # leave one graph for every later rank
            hi = lo
            while hi < cap and (hi == lo or acc + 0.5 * costs[hi] <= target):
                acc += costs[hi]
                hi += 1
        bounds.append((lo, hi))
        lo = hi
    return bounds


def balanced_shards(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Partition graphs ``0..len(costs)`` into ``world_size`` shards of (nearly) equal COUNT whose
    cost PROFILES match: the graphs are ranked by cost and dealt boustrophedon-wise (rank 0, 1, ..,
    W-1, W-1, .., 1, 0, ...), so every rank gets one graph of each size class -- including its share
    of the largest ones, which bound the fused kernels' run time.  The per-step barrier of the
    gradient exchange then waits for ranks that finish together, not for whoever drew the biggest
    graph.  Deterministic: every rank computes the same partition without communication.  Each
    shard lists its graphs in ascending id (the batch order of the loader is kept)."""
    n = len(costs)
    order = sorted(range(n), key=lambda i: (-float(costs[i]), i))
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for pos, i in enumerate(order):
        rnd, k = divmod(pos, world_size)
        shards[k if rnd % 2 == 0 else world_size - 1 - k].append(i)
    return [sorted(s) for s in shards]


def shard_ids(ids: Sequence[int], nodes: Sequence[int], edges: Sequence[int], world_size: int,
              rank: int):
    """The slice of a batch's graph ids that ``rank`` trains on when the data set is resident in
    HBM on every rank (``DeviceDataset``): contiguous in batch order, balanced by
    nodes + edges of the graphs (``nodes`` / ``edges``: per-graph sizes of the DATA SET, indexed
    by id).  Every rank computes the same split from the same ids: no communication."""
    costs = [float(nodes[int(i)]) + float(edges[int(i)]) for i in ids]
    lo, hi = shard_bounds(costs, world_size)[rank]
    return ids[lo:hi]


class GradBucket:
    """All parameter gradients as views into ONE flat fp32 buffer, plus ``extra``
    trailing slots for step scalars (loss sum, correct count) so that they ride
    the same collective.  ``p.grad`` aliases the bucket: autograd accumulates in
    place, the all-reduce needs no gather/scatter copies."""

    def __init__(self, params: Iterable[torch.nn.Parameter], extra: int = 2):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket needs at least one trainable parameter")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total + extra, dtype=torch.float32, device=dev)
        self.extra = self.flat[total:]
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self) -> None:
        self.flat.zero_()

    def all_reduce(self, global_batch: int, group=None) -> None:
        """Sum over ranks, then turn the summed (not averaged) local NLL gradients into
        the gradient of the GLOBAL-batch mean loss.  Scalars in ``extra`` stay sums."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat[: self.flat.numel() - self.extra.numel()].mul_(1.0 / float(global_batch))


class PeerExchange:
    """Exchange buffers of the fused all-reduce + Adam kernel (csrc/allreduce_adam.cu): every
    rank of ONE node allocates a small device buffer, shares it through CUDA IPC (torch's own
    storage-sharing handles, exchanged with all_gather_object) and maps the buffers of all its
    peers, so that the kernel can read the other ranks' gradient sums directly over
    NVLink / NVSwitch.  `ptrs[r]` is the device pointer of rank r's buffer as seen from this
    process.  Raises if the ranks cannot map each other's memory (e.g. several nodes): the
    caller then keeps the NCCL all-reduce."""

    def __init__(self, n_total: int, device: torch.device, group=None):
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised process group")
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        import ctypes
        lib = _lib.load_library()
        self._lib, self.device = lib, device
        local, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        with torch.cuda.device(device):
            _lib.check(lib.dgcnn_exchange_create(int(n_total), self.world, ctypes.byref(local), handle), "exchange_create")
        self.local_ptr = int(local.value)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, bytes(handle), group=group)
        self.ptrs, self.peer_ptrs = [], []
        for r, h in enumerate(gathered):
            if r == self.rank:
                self.ptrs.append(self.local_ptr)
                continue
            peer = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            with torch.cuda.device(device):          # OUR device current: mapped for OUR kernels
                _lib.check(lib.dgcnn_exchange_open(buf, ctypes.byref(peer)), "exchange_open")
            self.peer_ptrs.append(int(peer.value))
            self.ptrs.append(int(peer.value))
        self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
        dist.barrier(group=group)

    def close(self) -> None:
        """Unmap the peers' buffers and free the own one (after a barrier: nobody may still poll)."""
        if self.local_ptr is None:
            return
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for p in self.peer_ptrs:
                self._lib.dgcnn_exchange_close(p)
            self._lib.dgcnn_exchange_destroy(self.local_ptr)
        self.local_ptr, self.peer_ptrs = None, []
