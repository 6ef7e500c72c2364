"""One training step of the reference loop, train.py:35-45, as a straight launch sequence.

    data.to(device) -> model(data) -> NLL -> backward -> Adam.step -> zero_grad -> loss/acc

The reference (and the autograd path of this package) pays for an autograd graph, ~16
gradient-accumulation kernels into ``p.grad`` and a multi-launch optimizer per step.
``FusedTrainer`` calls the hand-written kernels directly, with every gradient written in
place into ONE flat buffer that the single all-reduce and the flat Adam consume:

    K0 (+K0b) -> KS -> KT forward -> NLL -> KT backward -> KSB -> [all-reduce] -> Adam

14 parameter tensors stay ordinary ``nn.Parameter``s (views of the flat buffers), so
``state_dict()``/checkpoints (train.py:129) are unchanged.  CUDA-graph capturable: no host
sync, step counter and dropout offset live on the device.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

from . import ops
from .nn import Model

__all__ = ["FusedTrainer"]


class FusedTrainer:
    def __init__(self, model: Model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 group=None, distributed: bool = True):
        convs = (model.conv1, model.conv2, model.conv3, model.conv4)
        self.model = model
        self.group = group
        self.distributed = bool(distributed)       # False: never exchange gradients (single-process reference)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        # flat order = KSB's gradient layout for the graph convolutions, then the dense tail
        self.stack_params = []
        for c in convs:
            if c.bias is None:
                raise ValueError("FusedTrainer needs GCNConv layers with bias (the reference's)")
            self.stack_params += [c.lin.weight, c.bias]
        self.tail_params = [model.conv5.weight, model.conv5.bias, model.conv6.weight, model.conv6.bias,
                            model.classifier_1.weight, model.classifier_1.bias,
                            model.classifier_2.weight, model.classifier_2.bias]
        self.params = self.stack_params + self.tail_params
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedTrainer: the model must live on a CUDA device")
        n = sum(p.numel() for p in self.params)
        self.num_params = n
        self.num_stack = sum(p.numel() for p in self.stack_params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n + 2, dtype=torch.float32, device=dev)   # + loss sum, #correct
        self.stats = self.grad[n:]
        off = 0
        self.tail_grad_views = []
        with torch.no_grad():
            for i, p in enumerate(self.params):
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                gview = self.grad[off:off + p.numel()].view_as(p)
                p.grad = gview
                if i >= len(self.stack_params):
                    self.tail_grad_views.append(gview)
                off += p.numel()
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        # dgcnn_train_step: the whole step as ONE call into the library (DGCNN_NATIVE_STEP=0 keeps
        # the one-by-one Python sequence below, which is also the fallback)
        self.native = os.environ.get("DGCNN_NATIVE_STEP", "1") != "0"
        self._arena = None
        self._resident_checked = None                 # (F, k, C) the data set was checked against
        self._fusable_cache = {}                      # (F, largest graph, fusion switch) -> conv5 fused
        # [0] flags of the last step, [1] sticky OR of the error flags of the earlier ones
        self._graph_status = torch.zeros(2, dtype=torch.int32, device=dev)
        # multi-GPU: gradients are summed by the fused peer-memory all-reduce + Adam kernel when
        # the ranks (one node) can map each other's memory, else by NCCL (DGCNN_ALLREDUCE=nccl)
        self.exchange = None
        self.comm_status = torch.zeros(1, dtype=torch.int32, device=dev)
        if (self.distributed and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
                and os.environ.get("DGCNN_ALLREDUCE", "p2p").lower() == "p2p"):
            try:
                from .dp import PeerExchange
                self.exchange = PeerExchange(n + 2, dev, group)
            except Exception as exc:                                       # noqa: BLE001
                print(f"[dgcnn_b200] peer-memory all-reduce unavailable ({exc!r}); using NCCL", flush=True)
            # all ranks must agree, or the collectives no longer match
            flag = torch.tensor([1.0 if self.exchange is not None else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if float(flag.item()) < 1.0:
                self.exchange = None

    def _world(self) -> int:
        if not self.distributed:
            return 1
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _global_batch(self, local_graphs: int, global_batch: Optional[int], world: int) -> int:
        """Number of graphs the mean NLL of train.py:39 runs over.  With several ranks it MUST be
        given: cost-balanced shards (dp.shard_ids) hold different numbers of graphs, and every
        rank has to scale the summed gradient by the same 1 / global_batch."""
        if global_batch is not None:
            if int(global_batch) < 1:
                raise ValueError("FusedTrainer: global_batch must be positive")
            return int(global_batch)
        if world > 1:
            raise ValueError("FusedTrainer: pass global_batch (graphs over ALL ranks) when training on "
                             "several GPUs; the ranks' shards need not hold the same number of graphs")
        return int(local_graphs)

    def status_words(self) -> torch.Tensor:
        """[comm_status, graph_status[0], graph_status[1]] as one float32 device tensor (flags are
        small integers), for callers that fold the check into a host read they do anyway."""
        return torch.cat([self.comm_status.reshape(-1)[:1], self._graph_status.reshape(-1)[:2]]).float()

    def check_status(self, values=None) -> None:
        """Check of the device status words (call once per epoch; host-syncing unless ``values`` =
        status_words() already on the host): a gradient exchange that timed out leaves the
        parameters untouched and raises here; bad input flags (BAD_EDGE / BAD_BATCH / fp16 RANGE)
        accumulate across steps until cleared."""
        if values is None:
            values = self.status_words().tolist()
        comm, graph = int(values[0]), int(values[1]) | int(values[2])
        if comm:
            raise RuntimeError("FusedTrainer: the peer-memory gradient exchange timed out (a rank is gone or "
                               "stuck); parameters were NOT updated by the affected steps")
        if graph & ~ops.GRAPH_GENERIC:
            self._graph_status.zero_()
            raise RuntimeError(f"FusedTrainer: the kernels flagged bad input (status {graph}): "
                               "1 = edge outside the batch, 2 = batch vector, 4 = fp16 split range")

    def supported(self, data) -> bool:
        """The fused kernels need the largest graph of the batch (host knowledge)."""
        from .nn import batch_max_nodes
        mx = batch_max_nodes(data)                 # data.max_nodes, or one cached device read
        f = data.x.size(1)
        return (ops.stack_fwd_supported(f, mx) and ops.stack_bwd_supported(f, mx)
                and self.model.classifier_2.out_features <= 32)

    def step(self, data, global_batch: Optional[int] = None) -> torch.Tensor:
        """One optimisation step on `data`; returns the device tensor
        [sum of NLL over the (global) batch, number of correct predictions]."""
        m = self.model
        if not self.supported(data):
            # a graph beyond the fused kernels (more than ~600 nodes): same step on the per-layer
            # kernels through autograd; the exchange and the optimiser are the same, so ranks may
            # take different branches for their shards of one global batch
            return self.step_autograd(data, global_batch)
        world = self._world()
        global_batch = self._global_batch(int(data.num_graphs), global_batch, world)
        if self.native and self._native_step(data, global_batch, world):
            return self.stats
        graph = m.build_graph(data)
        weights = self.stack_params[0::2]
        biases = self.stack_params[1::2]
        k, norm = m.sort_pool.k, m.conv1.norm
        if ops.conv5_fusable(data.x.size(1), graph.max_nodes):
            # no pooled / dpooled; KSB writes the GraphConv + conv5 gradients
            h1, arg, xcat, perm, _ = ops.stack_fwd_conv5(data.x, graph, weights, biases, self.tail_params[0],
                                                         self.tail_params[1], k, norm)
            logp, saved, tctx = ops.tail_fwd_loss(None, k, self.tail_params, data.y, m.training, m._tail_seed,
                                                  m._tail_rng_offset, h1=h1, arg=arg)
            dh1, _, pending = ops.tail_bwd_after_loss(tctx, logp, saved, k, self.tail_params, self.stats, True,
                                                      out_grads=self.tail_grad_views[2:], defer_join=True)
            ops.stack_bwd_conv5(dh1, arg, perm, xcat, data.x, graph, weights, self.tail_params[0], k, norm,
                                out=self.grad[:self.num_stack + 16 * 97 + 16])
        else:
            pooled, xcat, perm = ops.stack_fwd(data.x, graph, weights, biases, k, norm)
            logp, saved, tctx = ops.tail_fwd_loss(pooled, k, self.tail_params, data.y, m.training, m._tail_seed,
                                                  m._tail_rng_offset)
            # the tail's parameter gradients run on a side stream underneath KSB; joined before
            # the all-reduce / Adam read them
            dpooled, _, pending = ops.tail_bwd_after_loss(tctx, logp, saved, k, self.tail_params, self.stats, False,
                                                          out_grads=self.tail_grad_views, defer_join=True)
            ops.stack_bwd(dpooled, perm, xcat, data.x, graph, weights, k, norm,
                          out=self.grad[:self.num_stack])
        pending.join()
        return self._finish(global_batch, world)

    def _finish(self, global_batch: int, world: int) -> torch.Tensor:
        """Gradient sum over the ranks (one exchange of the flat buffer) + flat Adam."""
        if world > 1 and self.exchange is not None:
            ops.allreduce_adam(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count,
                               self.exchange.epoch, self.lr, self.betas[0], self.betas[1], self.eps,
                               1.0 / float(global_batch), self.exchange.ptrs, self.exchange.rank,
                               self.comm_status)
            return self.stats
        if world > 1 and os.environ.get("DGCNN_SKIP_ALLREDUCE", "0") != "1":   # (skip: timing experiments only)
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group)
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr,
                      self.betas[0], self.betas[1], self.eps, grad_scale=1.0 / float(global_batch))
        return self.stats

    def step_autograd(self, data, global_batch: Optional[int] = None) -> torch.Tensor:
        """The same optimisation step for ANY batch (graphs of any size: D&D's 5748-node graph,
        PROTEINS' 620): ``Model(data)`` through torch autograd -- the per-layer kernels K1 / K3
        where a graph exceeds one SM's shared memory -- with the gradients accumulated in place in
        the flat bucket, then the same exchange + flat Adam.  An empty shard (``data`` None or
        without graphs) still joins the exchange with zero gradients."""
        world = self._world()
        local = 0 if data is None else int(data.num_graphs)
        global_batch = self._global_batch(local, global_batch, world)
        self.grad.zero_()
        if local > 0:
            logp = self.model(data)
            loss = torch.nn.functional.nll_loss(logp, data.y, reduction="sum")
            loss.backward()
            with torch.no_grad():
                self.stats[0] = loss.detach()
                self.stats[1] = (logp.argmax(1) == data.y).sum()
        return self._finish(global_batch, world)

    def resident_supported(self, dataset, ids, max_nodes: Optional[int] = None) -> bool:
        """Can the one-call fused step run this batch of a DeviceDataset (every graph must fit the
        fused kernels)?  Otherwise use ``step_autograd(dataset.batch(ids))``.  ``max_nodes``: the
        batch's largest graph when the caller already knows it."""
        mx = int(max_nodes) if max_nodes is not None else dataset.plan(ids)[2]
        f = dataset.num_features
        return (ops.stack_fwd_supported(f, mx) and ops.stack_bwd_supported(f, mx)
                and self.model.classifier_2.out_features <= 32)

    def graph_stream(self) -> torch.cuda.Stream:
        """A stream of the trainer's own for graph-replayed steps (CUDA graphs cannot be captured on
        the legacy default stream)."""
        if getattr(self, "_graph_stream", None) is None:
            with torch.cuda.device(self.flat.device):
                self._graph_stream = torch.cuda.Stream()
        return self._graph_stream

    def step_resident(self, dataset, ids, ids_device: Optional[torch.Tensor] = None,
                      global_batch: Optional[int] = None, plan=None, graphed: bool = False) -> torch.Tensor:
        """One optimisation step on the graphs ``ids`` of a ``DeviceDataset``:
        dgcnn_collate gathers the batch from the resident data set, so neither the host
        collate (train.py:108-109) nor the host-to-device copy (train.py:36) nor K0 runs.
        Bit-identical to ``step()`` on the host-collated batch of the same graphs.  Returns
        the device tensor [sum of NLL, number of correct predictions].  ``plan``: (nodes, edges,
        largest graph) of the batch when the caller computed them already (driver.train_epoch does,
        for the whole epoch at once); ``ids_device``: the ids already on the device; ``graphed``:
        replay the step as a CUDA graph that is updated in place from call to call
        (dgcnn_train_step_resident_graphed; needs a non-default current stream, see graph_stream())."""
        import ctypes
        from . import _lib
        m = self.model
        lib = _lib.load_library()
        n, e, mx = plan if plan is not None else dataset.plan(ids)
        b, f = int(len(ids)), dataset.num_features
        k, c = m.sort_pool.k, m.classifier_2.out_features
        world = self._world()
        global_batch = self._global_batch(b, global_batch, world)
        if dataset.device != self.flat.device:
            raise RuntimeError("FusedTrainer.step_resident: data set and model live on different devices")
        if self._resident_checked != (f, k, c):
            if self.num_params != int(lib.dgcnn_train_step_num_params(f, k, c)):
                raise RuntimeError("FusedTrainer.step_resident: the model does not match the data set "
                                   "(num_features / k / num_classes)")
            self._resident_checked = (f, k, c)
        if world > 1 and self.exchange is None:
            raise RuntimeError("FusedTrainer.step_resident: multi-GPU needs the peer-memory exchange "
                               "(DGCNN_ALLREDUCE=p2p)")
        if ids_device is None:
            ids_device = dataset.ids_to_device(ids)
        need = int(lib.dgcnn_train_step_resident_workspace_bytes(n, e, b, f, k, c, mx))
        if need == 0:
            raise ValueError("FusedTrainer.step_resident: bad sizes")
        if self._arena is None or self._arena.numel() < need:
            self._arena = ops._empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=self.flat.device)
        table, epoch, rank = None, None, 0
        if world > 1:
            table = (ctypes.c_void_p * world)(*[ctypes.c_void_p(int(p_)) for p_ in self.exchange.ptrs])
            epoch, rank = self.exchange.epoch.data_ptr(), self.exchange.rank
        with torch.cuda.device(self.flat.device):
            entry = lib.dgcnn_train_step_resident_graphed if graphed else lib.dgcnn_train_step_resident
            rc = entry(
                dataset.c_struct, ids_device.data_ptr(), n, e, b, k, c, mx, int(m.conv1.norm),
                self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                self.step_count.data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps, int(global_batch),
                int(bool(m.training)), int(m._tail_seed) & 0xFFFFFFFFFFFFFFFF, m._tail_rng_offset.data_ptr(),
                table, world, rank, epoch, self.comm_status.data_ptr(), self._graph_status.data_ptr(),
                self._arena.data_ptr(), self._arena.numel(), torch.cuda.current_stream().cuda_stream)
        if rc == -2:
            raise RuntimeError("FusedTrainer.step_resident: the batch holds graphs too large for the fused "
                               "kernels (dgcnn_stack_fwd_supported); feed it through step()")
        _lib.check(rc, "train_step_resident")
        # n1_gather + KS + 5 tail fwd + NLL + 12 tail bwd + 2 KSB + 2 Adam; with conv5 fused into KS / KSB
        #  the tail loses its conv5 forward kernel and three backward ones
        fused = self._fusable_cache.get((f, mx, ops.FUSE_CONV5))
        if fused is None:
            fused = self._fusable_cache[(f, mx, ops.FUSE_CONV5)] = bool(ops.conv5_fusable(f, mx))
        launches = (16 if fused else 20) + (0 if b <= 1024 else 1)
        ops.LAUNCHES["train_step_resident"] = ops.LAUNCHES.get("train_step_resident", 0) + launches
        return self.stats

    def _native_step(self, data, global_batch, world) -> bool:
        """The same step through dgcnn_train_step (one ctypes call, buffers from a cached arena);
        False when the configuration is outside what that entry point covers."""
        import ctypes
        from . import _lib
        m = self.model
        x, ei, bt, y = data.x, data.edge_index, data.batch, data.y
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and
                ei.dtype in (torch.int32, torch.int64) and bt.dtype == ei.dtype and y.dtype == torch.int64 and
                ei.is_contiguous() and bt.is_contiguous() and y.is_contiguous() and
                ops.STACK_VARIANT == ops.STACK_MMA and ops.TAIL_OVERLAP and not ops.EXACT_SYMMETRY_CHECK):
            return False
        if world > 1 and self.exchange is None:
            return False                                   # NCCL path: keep the Python sequence
        lib = _lib.load_library()
        n, f = x.shape
        e, b = ei.size(1), int(data.num_graphs)
        k, c = m.sort_pool.k, m.classifier_2.out_features
        mx = int(data.max_nodes)
        if self.num_params != int(lib.dgcnn_train_step_num_params(f, k, c)) or e < 1:
            return False
        need = int(lib.dgcnn_train_step_workspace_bytes(n, e, b, f, k, c, mx))
        if self._arena is None or self._arena.numel() < need:
            self._arena = ops._empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=x.device)
        table, epoch, rank = None, None, 0
        if world > 1:
            table = (ctypes.c_void_p * world)(*[ctypes.c_void_p(int(p_)) for p_ in self.exchange.ptrs])
            epoch, rank = self.exchange.epoch.data_ptr(), self.exchange.rank
        with torch.cuda.device(x.device):
            rc = lib.dgcnn_train_step(
                x.data_ptr(), int(x.stride(0)) if n > 1 else max(int(x.stride(0)), f), ei.data_ptr(),
                int(ei.dtype == torch.int32), bt.data_ptr(), y.data_ptr(), n, e, b, f, k, c, mx, int(m.conv1.norm),
                self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                self.step_count.data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps, int(global_batch),
                int(bool(m.training)), int(m._tail_seed) & 0xFFFFFFFFFFFFFFFF, m._tail_rng_offset.data_ptr(),
                table, world, rank, epoch, self.comm_status.data_ptr(), self._graph_status.data_ptr(),
                self._arena.data_ptr(), self._arena.numel(), torch.cuda.current_stream().cuda_stream)
        if rc == -2:                                       # DGCNN_ERR_UNSUPPORTED: graphs too large for KS / KSB
            return False
        _lib.check(rc, "train_step")
        # status, K0 (3), K0b (3; lazy maps: 2), KS, tail forward (3) and backward (7), KSB (2), Adam (2);
        # without the conv5 fusion four more (conv5 forward / backward, SortPool gradient)
        lazy = ops.LAZY_MAPS and int(lib.dgcnn_stack_bwd_supported(int(f), int(mx))) == 1
        ops.LAUNCHES["train_step"] = ops.LAUNCHES.get("train_step", 0) + \
            (22 if ops.conv5_fusable(f, mx) else 26) - (1 if lazy else 0)
        return True
