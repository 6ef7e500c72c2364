#!/usr/bin/env python
"""Benchmark of the DGCNN hot path (BASELINE.json: graphs/sec DGCNN fwd+bwd on
COLLAB-synth at 1/2/4/8 B200; GraphConv HBM GB/s as % of the measured peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one training step of the reference's loop (train.py:35-42) on one
COLLAB-shaped synthetic batch of 512 graphs per GPU: graph build (K0), 4x fused
GraphConv + SortPool forward, dense tail, NLL, backward through everything, the
single gradient all-reduce when N > 1, Adam.  Prints ONE JSON line on rank 0.

`value`      graphs/s with the batch (x, edge_index, batch) already resident in HBM
`e2e`        the same step through the public API from pinned HOST buffers, the H2D
             copy and the D2H loss read inside the timed region
`e2e_resident_dataset`   the same step fed from a data set resident in HBM: per
             step only the graph ids cross PCIe; reported next to `e2e`, never instead of it
`roofline`   the dominant kernel, timed alone with CUDA events, against MEASURED_PEAKS
`cpu_baseline` / `--impl reference`   the CPU oracle restatement of the reference's
             path (PyG itself is not installable here), timed on this host's cores

N > 1 (torchrun, one rank per GPU): weak scaling, 512 graphs per rank and step; `--shards balanced`
(default) deals a global batch of 512 N graphs with dp.balanced_shards.  The ranks' timed steps
are queued behind a device-side rendezvous right after the host barrier; `first_timed_step_ms` /
`ms_per_step_after_first` show what start skew is left, `per_rank` each rank's own view,
`params_equal_across_ranks` / `comm_status_per_rank` that the replicas stayed identical;
`--trace-exchange FILE` records the exchange kernel's %globaltimer timeline per step and rank.
Other BASELINE configs: `--workload dd|powerlaw|proteins|mutag`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F

METRIC = "graphs_per_sec_fwd_bwd"
UNIT = "graphs/s"
WORKLOAD = "collab"          # BASELINE.json configs[3]: the config the metric is quoted on
RING = 4                     # distinct pre-built batches cycled through the timed steps
L2_FLUSH_BYTES = 256 << 20   # > 126 MB L2
# DRAM bytes (read + write) of one stack_fwd_mma_kernel launch on COLLAB-synth bs512 cannot be
# measured live; they are parsed from the committed ncu --set full summary of the CURRENT kernel
KS_PROFILE = os.path.join(ROOT, "profiles", "r02_stack_fwd_mma.md")


def ks_traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the first launch listed in
    profiles/r02_stack_fwd_mma.md (written by scripts/summarize_ncu.py), or None."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        total, seen = 0.0, 0
        for line in open(KS_PROFILE):
            cells = [c.strip() for c in line.split("|")]
            if len(cells) >= 4 and cells[1] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(cells[2].replace(",", "")) * unit.get(cells[3], 1.0)
                seen += 1
                if seen == 2:
                    return int(total)
    except (OSError, ValueError):
        pass
    return None


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-resident", action="store_true", help="skip the resident-data-set leg")
    ap.add_argument("--no-fuse-conv5", action="store_true",
                    help="keep conv5 + ReLU + max-pool out of KS / KSB (SURVEY 8f N2 off; A/B timing)")
    ap.add_argument("--autograd", action="store_true",
                    help="step through torch autograd (Model + GradBucket + FlatAdam) instead of FusedTrainer")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--shards", default="balanced", choices=["balanced", "independent", "same"],
                    help="N > 1: how the global batch of bs*N graphs is split (balanced, the default: "
                         "dp.balanced_shards, every rank gets the same count and the same size profile; "
                         "independent: every rank draws its own bs graphs; same: every rank the same batches, "
                         "a diagnostic).  Measured at N = 8 after the cluster-pair split of large graphs: "
                         "0.317 ms balanced, 0.318 ms independent (profiles/r02_scaling.md)")
    ap.add_argument("--seed-rank", type=int, default=-1,
                    help="diagnostic, N = 1: draw the batches rank R of a multi-GPU run would draw")
    ap.add_argument("--trace-exchange", default="",
                    help="N > 1: write the gradient-exchange kernel's per-step timeline of every rank to this JSON")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# --------------------------------------------------------------------------------------
# clocks: sampled on a thread DURING the timed region (pynvml; nvidia-smi as a fallback)
# --------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                getter = getattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    self.nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = getter(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------
# algorithmic bytes (fp32 features, int32 CSR)
# --------------------------------------------------------------------------------------
def layer_bytes(n, e, cin, cout):
    return 4 * (n + 1) + 4 * e + 4 * n + 4 * n * cin + 4 * cin * cout + 4 * cout + 4 * n * cout


def sortpool_bytes(n, b, k, d=97):
    return 4 * n * d + 4 * (b + 1) + 4 * b * k * d + 4 * b * k


def forward_bytes(n, e, b, f, k):
    dims = [(f, 32), (32, 32), (32, 32), (32, 1)]
    return sum(layer_bytes(n, e, ci, co) for ci, co in dims) + sortpool_bytes(n, b, k)


# --------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference's own path, all host threads
# --------------------------------------------------------------------------------------
def cpu_reference_step_fn(cfg, batch):
    from oracle import dgcnn_oracle as orc      # the one place bench.py touches oracle/
    torch.manual_seed(324)
    model = orc.OracleModel(cfg.num_features, cfg.num_classes, cfg.k).train()
    opt = torch.optim.Adam(model.parameters())

    def step():
        logp = model(batch)
        loss = F.nll_loss(logp, batch.y)
        loss.backward()
        opt.step()
        opt.zero_grad()
        return float(loss.detach())
    return step


def time_cpu(cfg, batch, seconds, warmup=1):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_reference_step_fn(cfg, batch)
    for _ in range(warmup):
        step()
    t0, n = time.perf_counter(), 0
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds or n >= 200:
            break
    return batch.num_graphs * n / el, n, el


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dgcnn_b200.synth import CONFIGS, make_batch
    cfg = CONFIGS[args.workload]
    batch = make_batch(args.workload, seed=324)
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_reference_step_fn(cfg, batch)
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    budget = 150.0                                  # keep the whole arm within a few minutes
    if first * (args.steps + args.warmup) > budget:
        keep = max(16, int(batch.num_graphs * budget / (first * (args.steps + args.warmup))))
        batch = make_batch(args.workload, seed=324, num_graphs=keep)   # bounded sample
        step = cpu_reference_step_fn(cfg, batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    value = batch.num_graphs * args.steps / el
    sample = (f"{args.steps} train steps (fwd+NLL+bwd+Adam) on one {args.workload}-synth batch of "
              f"{batch.num_graphs} graphs, N={batch.num_nodes}, E={batch.num_edges}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}-synth bs{cfg.batch_size}/GPU k{cfg.k} "
                               f"F{cfg.num_features} (BASELINE.json configs[3])",
                   "sample_graphs": batch.num_graphs,
                   "nodes_per_batch": batch.num_nodes, "edges_per_batch": batch.num_edges,
                   "global_batch": batch.num_graphs,
                   "note": "CPU oracle restatement of model.py:26-45 + train.py:35-42 "
                           "(torch_geometric is not installable here); rank 0 only"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# device time of a launch sequence
# --------------------------------------------------------------------------------------
def timed(fn, flush, reps=20):
    """Mean device time of fn(): captured once in a CUDA graph and replayed, so that
    Python/ctypes launch overhead never sits between the two events; L2 is flushed
    (untimed) before every replay."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # the two timing events are nodes of the captured graph (external events), right
    # before and after the kernels of fn(): the graph-launch latency stays outside
    try:
        a = torch.cuda.Event(enable_timing=True, external=True)
        b_ = torch.cuda.Event(enable_timing=True, external=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            a.record()
            fn()
            b_.record()
        inside = True
    except Exception:                          # noqa: BLE001  (older torch: events around the replay)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        inside = False
    times = []
    for _ in range(reps):
        flush.zero_()
        if inside:
            g.replay()
        else:
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b_.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b_))
    return statistics.mean(times) * 1e-3


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    import dgcnn_b200 as dg
    from dgcnn_b200 import ops
    from dgcnn_b200.synth import CONFIGS, make_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.no_fuse_conv5:
        ops.set_fuse_conv5(False)
    cfg = CONFIGS[args.workload]
    # weak scaling: every rank owns RING distinct batches of cfg.batch_size graphs.  With N > 1 the
    # step's GLOBAL batch (bs * N graphs, the same on every rank) is split with dp.balanced_shards:
    # equal counts, matching size profiles -- the per-step barrier of the gradient exchange then
    # waits for ranks that finish together (power-law graphs all have 1000 nodes: nothing to balance)
    from dgcnn_b200.synth import collate, make_graphs
    balanced = world > 1 and args.shards == "balanced" and cfg.kind != "powerlaw"
    ring_graphs = []                                   # this rank's graphs per ring slot
    for i in range(RING):
        if balanced:
            pool = make_graphs(cfg, cfg.batch_size * world, seed=324 + i)
            costs = [g["x"].shape[0] + g["edge_index"].shape[1] for g in pool]
            mine = dg.balanced_shards(costs, world)[rank]
            ring_graphs.append([pool[j] for j in mine])
        else:
            ring_graphs.append(make_graphs(cfg, cfg.batch_size, seed=324 + (0 if args.shards == "same" else 1000 * (args.seed_rank if args.seed_rank >= 0 else rank)) + i))
    host_batches = [collate(gs).pin_memory() for gs in ring_graphs]
    for hb in host_batches:
        hb.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
    dev_batches = []
    for hb in host_batches:
        db = hb.to(dev)
        db.max_nodes = hb.max_nodes
        dev_batches.append(db)
    graphs_per_step = cfg.batch_size
    global_batch = graphs_per_step * world

    torch.manual_seed(324)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).train()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    fused_step = not args.autograd
    if fused_step:
        # train.py:35-45 as a straight launch sequence (dgcnn_b200/trainer.py): no autograd,
        # gradients written in place into one flat buffer, one all-reduce, flat Adam
        trainer = dg.FusedTrainer(model, lr=1e-3)
        if not trainer.supported(dev_batches[0]):
            fused_step = False
    if fused_step:
        def train_step(data):
            return trainer.step(data, global_batch)[0]
    else:
        bucket = dg.GradBucket(model.parameters(), extra=2)
        opt = dg.FlatAdam(model, bucket, lr=1e-3)    # train.py:99 Adam defaults, one flat kernel

        def train_step(data):
            bucket.zero_()
            logp = model(data)
            loss = F.nll_loss(logp, data.y, reduction="sum")
            loss.backward()
            bucket.extra[0].copy_(loss.detach())
            bucket.all_reduce(global_batch)
            opt.step()
            return loss

    trace_buf = None
    if world > 1 and args.trace_exchange and fused_step and getattr(trainer, "exchange", None) is not None:
        from dgcnn_b200 import _lib as _dl
        trace_buf = torch.zeros(1024, 4, dtype=torch.int64, device=dev)
        _dl.load_library().dgcnn_allreduce_set_trace(trace_buf.data_ptr())

    # ---- launch census on one eager step ------------------------------------------
    before = ops.launches_total()
    train_step(dev_batches[0])
    torch.cuda.synchronize()
    launches_per_step = ops.launches_total() - before

    # ---- optional CUDA-graph capture: one graph per ring slot ---------------------
    use_graph = not args.no_graph
    graphs, static_loss = [], []
    if use_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for db in dev_batches:
                    train_step(db)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            pool = None
            for db in dev_batches:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    static_loss.append(train_step(db))
                pool = g.pool()
                graphs.append(g)
        except Exception as exc:                       # noqa: BLE001
            if rank == 0:
                print(f"[bench] CUDA-graph capture failed ({exc!r}); timing eagerly", file=sys.stderr)
            use_graph, graphs = False, []
            torch.cuda.synchronize()

    def run_slot(i):
        if use_graph:
            graphs[i % RING].replay()
        else:
            train_step(dev_batches[i % RING])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, L2 flushed before each, CUDA events ----
    for i in range(max(args.warmup, 3)):
        flush.zero_()
        run_slot(i)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    # The sampler (NVML initialisation: milliseconds, different on every rank) is set up and started
    # BEFORE the barrier: whatever a rank's host does between the barrier and its first launch is
    # skew that the other ranks wait out inside the first timed step's gradient exchange (measured
    # at N = 8: 1.5-3 ms in step 1, i.e. 75-150 us per step of a 20-40 step run;
    # profiles/r02_scaling.md).
    sync_token = torch.zeros(1, device=dev)
    with ClockSampler(local_rank) as clocks:
        barrier()
        if world > 1:
            # device-side rendezvous on top of the host barrier: the collective completes on every
            # GPU at the same moment, and each rank's timed steps are queued behind it -- the hosts
            # leave dist.barrier() up to milliseconds apart, the GPUs now start together
            dist.all_reduce(sync_token)
        wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()                              # evict the batch / weights from L2
            starts[i].record()
            run_slot(i)
            stops[i].record()
        barrier()
        wall = time.perf_counter() - wall0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    # where the time of a multi-rank run sits: the first timed step absorbs the ranks' skew after the
    # barrier (every rank waits for the last one inside the exchange kernel); the rest is steady state
    first_rest = torch.tensor([step_ms[0], sum(step_ms[1:]) / max(1, args.steps - 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(first_rest, op=dist.ReduceOp.MAX)
    first_step_ms, steady_ms = (float(v) for v in first_rest.tolist())
    ranks_report = None
    if world > 1:
        # each rank's own view: its device time per step (the wait for the slowest rank inside the
        # exchange kernel included), its SM clock under load, the largest graph of each ring batch
        rank_report = {"rank": rank, "ms_per_step": round(sum(step_ms) / args.steps, 4),
                "first_step_ms": round(step_ms[0], 4),
                "sm_mhz": clocks.summary().get("sm_mhz"), "reasons": clocks.summary().get("reasons"),
                "largest_graph_per_ring_batch": [hb.max_nodes for hb in host_batches],
                "launches_per_step": launches_per_step}
        ranks_report = [None] * world
        dist.all_gather_object(ranks_report, rank_report)
    value = global_batch * args.steps / (total_ms / 1e3)

    # ---- every rank must hold the same parameters after the same updates ------------
    params_equal, comm_status_all = None, None
    if world > 1 and fused_step:
        flat = trainer.flat.double()
        digest = torch.stack([flat.sum(), flat.abs().sum(), (flat * torch.arange(1, flat.numel() + 1, device=dev,
                                                                                dtype=torch.float64)).sum(),
                              trainer.comm_status.double().sum()])
        gathered = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        params_equal = all(torch.equal(g_[:3], gathered[0][:3]) for g_ in gathered)
        comm_status_all = [int(g_[3].item()) for g_ in gathered]
    if trace_buf is not None:
        t_all = [torch.zeros_like(trace_buf) for _ in range(world)]
        dist.all_gather(t_all, trace_buf)
        if rank == 0:
            steps_done = int(trainer.exchange.epoch.item())
            rows = []
            for e in range(max(1, steps_done - args.steps), steps_done):
                per_rank = [t_all[r][e % 1024].tolist() for r in range(world)]
                prev = [t_all[r][(e - 1) % 1024].tolist() for r in range(world)]
                rows.append({"step": e,
                             # the rank's own work between two exchanges: L2 flush + graph launch + K0 .. KSB
                             "compute_us": [round((p_[0] - q_[3]) / 1e3, 2) for p_, q_ in zip(per_rank, prev)],
                             "wait_us": [round((p_[2] - p_[1]) / 1e3, 2) for p_ in per_rank],
                             "push_us": [round((p_[1] - p_[0]) / 1e3, 2) for p_ in per_rank],
                             "sum_adam_us": [round((p_[3] - p_[2]) / 1e3, 2) for p_ in per_rank],
                             "enter_skew_us": round((max(p_[0] for p_ in per_rank) - min(p_[0] for p_ in per_rank)) / 1e3, 2)})
            import statistics as _st
            summary = {"world": world, "steps": len(rows),
                       "mean_wait_us_per_rank": [round(_st.mean(r_["wait_us"][k] for r_ in rows), 2) for k in range(world)],
                       "mean_compute_us_per_rank": [round(_st.median(r_["compute_us"][k] for r_ in rows), 2) for k in range(world)],
                       "mean_of_max_compute_us": round(_st.mean(max(r_["compute_us"]) for r_ in rows), 2),
                       "mean_of_mean_compute_us": round(_st.mean(_st.mean(r_["compute_us"]) for r_ in rows), 2),
                       "mean_push_us": round(_st.mean(_st.mean(r_["push_us"]) for r_ in rows), 2),
                       "mean_sum_adam_us": round(_st.mean(_st.mean(r_["sum_adam_us"]) for r_ in rows), 2),
                       "mean_enter_skew_us": round(_st.mean(r_["enter_skew_us"] for r_ in rows), 2),
                       "note": "per step and rank, %globaltimer inside allreduce_adam_kernel: push = entered -> sums "
                               "pushed; wait = pushed -> every rank arrived (the straggler cost); sum_adam = the rest. "
                               "enter_skew = latest minus earliest kernel entry over the ranks (GPU timers are only "
                               "loosely synchronised across devices)", "rows": rows}
            os.makedirs(os.path.dirname(os.path.abspath(args.trace_exchange)), exist_ok=True)
            with open(args.trace_exchange, "w") as fh:
                json.dump(summary, fh, indent=1)

    # ---- end to end from pinned host buffers (public API, eager) ------------------
    # Every step's batch starts in pinned HOST memory.  Like a DataLoader with pin_memory +
    # non_blocking copies, the H2D copy of step i+1 is issued on a copy stream while step i
    # computes; both the copies and the D2H loss read are inside the timed region.
    copy_stream = torch.cuda.Stream()

    def device_slots(hbs):
        """One preallocated device batch per ring slot (a prefetching loader's staging buffers):
        the H2D copy lands in place, no allocation in the loop."""
        slots = []
        for hb in hbs:
            slot = hb._map(lambda t: torch.empty_like(t, device=dev))
            slot.max_nodes = hb.max_nodes
            slots.append(slot)
        return slots

    def h2d(hb, slot):
        with torch.cuda.stream(copy_stream):
            for src, dst in ((hb.x, slot.x), (hb.edge_index, slot.edge_index), (hb.batch, slot.batch),
                             (hb.ptr, slot.ptr), (hb.y, slot.y)):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return slot, ev

    def e2e_run(nsteps, hbs=None, slots=None):
        hbs = host_batches if hbs is None else hbs
        slots = e2e_slots if slots is None else slots
        done = [None] * RING                            # compute of the step that last used the slot
        nxt = h2d(hbs[0], slots[0])
        last = 0.0
        for i in range(nsteps):
            data, ev = nxt
            if i + 1 < nsteps:
                j = (i + 1) % RING
                if done[j] is not None:
                    copy_stream.wait_event(done[j])     # do not overwrite a batch still being read
                nxt = h2d(hbs[j], slots[j])
            torch.cuda.current_stream().wait_event(ev)
            loss = train_step(data)
            done[i % RING] = torch.cuda.Event()
            done[i % RING].record()
            last = float(loss.item())                  # D2H read of the step's result
        return last

    e2e_slots = device_slots(host_batches)
    e2e_steps = max(5, min(args.steps, 20))
    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = global_batch * e2e_steps / float(e2e_s.item())
    h2d_bytes = host_batches[0].nbytes()

    # the same loop on COMPACT host batches (int32 indices, dgcnn_build_graph_i32): not the
    # reference's tensor format, reported next to `e2e`, never instead of it
    compact_batches = []
    for hb in host_batches:
        cb = hb.compact().pin_memory()
        cb.max_nodes = hb.max_nodes
        compact_batches.append(cb)
    compact_slots = device_slots(compact_batches)
    e2e_run(3, compact_batches, compact_slots)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps, compact_batches, compact_slots)
    barrier()
    e2e_c = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_c, op=dist.ReduceOp.MAX)
    e2e_compact_value = global_batch * e2e_steps / float(e2e_c.item())
    h2d_compact_bytes = compact_batches[0].nbytes()

    # ---- N > 1: the same training step fed from data sets resident in HBM (every rank holds the
    # graphs of ITS shards; per step only the graph ids cross PCIe, the exchange is unchanged) ----
    resident_multi = None
    if world > 1 and fused_step and not args.no_resident and getattr(trainer, "exchange", None) is not None:
        try:
            import numpy as np
            ds_m = dg.DeviceDataset([g_ for gs in ring_graphs for g_ in gs], dev, num_classes=cfg.num_classes)
            bs_m = cfg.batch_size
            id_sets_m = [np.arange(i * bs_m, (i + 1) * bs_m, dtype=np.int32) for i in range(RING)]
            id_pinned_m = [torch.from_numpy(a_).pin_memory() for a_ in id_sets_m]
            model.train()

            def resident_run_m(nsteps):
                last = 0.0
                for i in range(nsteps):
                    ids_dev = id_pinned_m[i % RING].to(dev, non_blocking=True)
                    stats = trainer.step_resident(ds_m, id_sets_m[i % RING], ids_dev, global_batch)
                    last = float(stats[0].item())
                return last

            resident_run_m(3)
            barrier()
            t0 = time.perf_counter()
            resident_run_m(e2e_steps)
            barrier()
            res_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(res_s, op=dist.ReduceOp.MAX)
            ids0_m = id_pinned_m[0].to(dev)
            t_dev = torch.tensor([timed(lambda: trainer.step_resident(ds_m, id_sets_m[0], ids0_m, global_batch), flush)],
                                 dtype=torch.float64, device=dev)
            dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
            resident_multi = {"value": global_batch * e2e_steps / float(res_s.item()), "unit": UNIT,
                              "h2d_bytes_per_step": 4 * bs_m, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                              "device_step_us": float(t_dev.item()) * 1e6,
                              "device_value": global_batch / float(t_dev.item()),
                              "note": "dgcnn_train_step_resident on every rank (its shard's graphs resident in HBM), "
                                      "fused peer-memory exchange + Adam; wall clock and device time = max over ranks"}
        except Exception as exc:                       # noqa: BLE001
            resident_multi = {"error": repr(exc)}
            if rank == 0:
                print(f"[bench] multi-GPU resident leg failed: {exc!r}", file=sys.stderr)

    if rank != 0:
        # Nothing collective happens after this point.  Tearing down an NCCL communicator whose
        # kernels live inside captured CUDA graphs can block forever: leave without ceremony.
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)

    comm_kind = "none (1 GPU)"
    if world > 1:
        p2p = fused_step and getattr(trainer, "exchange", None) is not None
        comm_kind = ("one-shot NVLink peer-memory all-reduce fused with Adam (csrc/allreduce_adam.cu)" if p2p
                     else "NCCL all-reduce + flat Adam")
        if p2p and int(trainer.comm_status.item()) != 0:
            comm_kind += " -- COMM TIMEOUT RAISED, numbers invalid"

    # ---- roofline: forward hot path and its dominant kernel, each timed alone ----
    peak, peak_kind = measured_peaks()
    db0 = dev_batches[0]
    n, e = db0.num_nodes, db0.num_edges
    model.eval()
    with torch.no_grad():
        g0 = model.build_graph(db0)
        convs = (model.conv1, model.conv2, model.conv3, model.conv4)

        t_fwd = timed(lambda: model.hot_path(db0.x, g0), flush)
        xcat = torch.empty(n, ops.XCAT_LD, device=dev)[:, :97]       # 16-byte aligned rows, like the model's
        ops.graph_conv_fwd(db0.x, g0.rowptr, g0.col, g0.dis, convs[0].lin.weight, convs[0].bias,
                           0, 1, xcat[:, 0:32])
        t_l2 = timed(lambda: ops.graph_conv_fwd(xcat[:, 0:32], g0.rowptr, g0.col, g0.dis,
                                                convs[1].lin.weight, convs[1].bias, 0, 1,
                                                xcat[:, 32:64], graph=g0), flush)
        t_k0 = timed(lambda: model.build_graph(db0), flush)
        t_k0_only = timed(lambda: ops.build_graph(db0.edge_index, db0.batch, n, db0.num_graphs,
                                                  transpose=False, max_nodes=0), flush)
    # SURVEY 8f N2 variants of the two fused kernels (what the training step runs): KS that emits
    # h1 / arg instead of `pooled`, KSB fed with d(h1)
    n2 = None
    if fused_step and ops.conv5_fusable(cfg.num_features, db0.max_nodes):
        with torch.enable_grad():
            g_t = model.build_graph(db0)               # with A_hat^T / transposed maps for the backward
        with torch.no_grad():
            ws_ = [c.lin.weight for c in convs]
            bs_ = [c.bias for c in convs]
            w5_, b5_ = model.conv5.weight, model.conv5.bias
            t_ks5 = timed(lambda: ops.stack_fwd_conv5(db0.x, g_t, ws_, bs_, w5_, b5_, cfg.k, 0), flush)
            h1_, arg_, xcat_, perm_, _ = ops.stack_fwd_conv5(db0.x, g_t, ws_, bs_, w5_, b5_, cfg.k, 0)
            dh1_ = torch.randn_like(h1_)
            t_ksb5 = timed(lambda: ops.stack_bwd_conv5(dh1_, arg_, perm_, xcat_, db0.x, g_t, ws_, w5_, cfg.k, 0), flush)
        bsz, kk = cfg.batch_size, cfg.k
        a_ks5 = (forward_bytes(n, e, bsz, cfg.num_features, kk) - 4 * bsz * kk * 97       # no pooled
                 + bsz * 16 * (kk // 2) * 5 + 4 * 16 * 98)                                # + h1, arg, W5, b5
        n2 = {"what": "SURVEY 8f N2: conv5 + ReLU + MaxPool1d(2,2) inside KS, their backward inside KSB; "
                      "SortPooling's [B, k*97] output and its gradient are never materialised",
              "stack_fwd_conv5_us": t_ks5 * 1e6, "algorithmic_bytes": a_ks5,
              "achieved_GBps": a_ks5 / t_ks5 / 1e9, "frac_of_peak": a_ks5 / t_ks5 / 1e9 / peak,
              "stack_bwd_conv5_us": t_ksb5 * 1e6}
    a_fwd = forward_bytes(n, e, cfg.batch_size, cfg.num_features, cfg.k)
    a_l2 = layer_bytes(n, e, 32, 32)
    # stricter figure when the layers are fused (x_1..x_3 never re-read from HBM)
    a_stack = (4 * (n + 1) + 4 * e + 4 * n + 4 * n * cfg.num_features + 4 * n * 97
               + 4 * (cfg.batch_size + 1) + 4 * cfg.batch_size * cfg.k * 98)
    fused = ops.stack_fwd_supported(cfg.num_features, db0.max_nodes) and dg.fused_enabled()
    per_layer = {"kernel": "gc_aggregate_staged / gc_aggregate_vec32 (GraphConv 32->32 forward, per-layer path)",
                 "algorithmic_bytes": a_l2, "launch_us": t_l2 * 1e6,
                 "achieved_GBps": a_l2 / t_l2 / 1e9, "frac_of_peak": a_l2 / t_l2 / 1e9 / peak}
    if fused:
        roofline = {"bound": "hbm",
                    "kernel": "stack_fwd_kernel (GraphConv x4 + SortPool forward, one launch)",
                    "achieved": a_fwd / t_fwd / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": a_fwd / t_fwd / 1e9 / peak, "traffic": ks_traffic_from_profile(), "peak_source": peak_kind,
                    # K0b (bitmaps / fragment maps, once per batch) does the adjacency read that A_fwd
                    # charges to every layer: the same fraction with its time added to the launch
                    "frac_with_k0b": a_fwd / (t_fwd + max(t_k0 - t_k0_only, 0.0)) / 1e9 / peak,
                    "k0b_us": max(t_k0 - t_k0_only, 0.0) * 1e6,
                    "algorithmic_bytes": a_fwd, "launch_us": t_fwd * 1e6,
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
                                      "capture of this kernel on this workload, parsed from profiles/r02_stack_fwd_mma.md: "
                                      "adjacency arrives as K0b bitmaps and the 41 MB of outputs stay in the "
                                      "126 MB L2, so DRAM traffic is far BELOW the algorithmic bytes",
                    "note": "effective figure on A_fwd = sum of per-layer algorithmic bytes + SortPool "
                            "(SURVEY 8d); the fused kernel's stricter A_stack figure is in hot_path_fwd"}
    else:
        roofline = {"bound": "hbm", "kernel": per_layer["kernel"], "achieved": per_layer["achieved_GBps"],
                    "peak": peak, "unit": "GB/s", "frac": per_layer["frac_of_peak"], "traffic": None,
                    "peak_source": peak_kind, "algorithmic_bytes": a_l2, "launch_us": t_l2 * 1e6}
    hot_fwd = {"what": "GraphConv x4 + SortPool forward (A_fwd, SURVEY 8d)", "fused_one_launch": fused,
               "algorithmic_bytes": a_fwd, "us": t_fwd * 1e6, "achieved_GBps": a_fwd / t_fwd / 1e9,
               "frac_of_peak": a_fwd / t_fwd / 1e9 / peak, "a_stack_bytes": a_stack,
               "a_stack_GBps": a_stack / t_fwd / 1e9, "a_stack_frac_of_peak": a_stack / t_fwd / 1e9 / peak,
               "graph_build_us": t_k0 * 1e6,
               "graph_build": {"what": "K0 + K0b: int64 COO -> int32 CSR, dis, bitmaps, fragment maps, work "
                                       "descriptors (once per batch, shared by 4 layers fwd + bwd)",
                               "algorithmic_bytes": 24 * e, "us": t_k0 * 1e6,
                               "achieved_GBps": 24 * e / t_k0 / 1e9,
                               "frac_of_peak": 24 * e / t_k0 / 1e9 / peak},
               "per_layer_kernel": per_layer, "conv5_fused_variant": n2}

    # ---- resident data set: the ring's graphs live in HBM, a step's input is
    # its list of graph ids; dgcnn_collate replaces the host collate, the H2D copy and K0 ----
    resident = None
    if fused_step and world == 1 and not args.no_resident:
        try:
            import numpy as np
            from dgcnn_b200.synth import make_graphs
            ds_graphs = []
            # the same graphs as host_batches[i], then more up to nine steps per epoch: COLLAB's
            # training fold (train.py:89-95, 4500 of 5000 graphs) is nine batches of 512
            for i in range(max(RING, 9 if args.workload in ("collab", "proteins", "mutag") else RING)):
                ds_graphs += make_graphs(cfg, cfg.batch_size, seed=324 + 1000 * rank + i)
            ds = dg.DeviceDataset(ds_graphs, dev, num_classes=cfg.num_classes)
            bs = cfg.batch_size
            id_sets = [np.arange(i * bs, (i + 1) * bs, dtype=np.int32) for i in range(RING)]
            id_pinned = [torch.from_numpy(a).pin_memory() for a in id_sets]
            model.train()

            def resident_run(nsteps):
                last = 0.0
                for i in range(nsteps):
                    ids_dev = id_pinned[i % RING].to(dev, non_blocking=True)     # the step's H2D: 4 B / graph
                    stats = trainer.step_resident(ds, id_sets[i % RING], ids_dev, global_batch)
                    last = float(stats[0].item())          # D2H read of the step's result
                return last

            resident_run(3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            resident_run(e2e_steps)
            torch.cuda.synchronize()
            res_s = time.perf_counter() - t0
            # the product's own loop (dgcnn_b200.driver.train_epoch = train.py:27-47): shuffled ids,
            # one library call per step, loss / accuracy accumulated on the device and read ONCE
            # per epoch (the reference syncs twice per batch, train.py:44-45)
            from dgcnn_b200 import driver as drv
            all_ids = np.arange(len(ds), dtype=np.int64)
            gen = torch.Generator().manual_seed(324)
            drv.train_epoch(trainer, ds, all_ids, bs, gen)
            torch.cuda.synchronize()
            epochs = max(3, e2e_steps // max(1, len(ds) // bs))
            t0 = time.perf_counter()
            for _ in range(epochs):
                ep_loss, ep_acc = drv.train_epoch(trainer, ds, all_ids, bs, gen)
            torch.cuda.synchronize()
            epoch_s = time.perf_counter() - t0
            # device time of the gather (+ K0b, like graph_build_us) and of one whole resident step
            ids0 = id_pinned[0].to(dev)
            with torch.no_grad():
                t_gather = timed(lambda: ds.batch(id_sets[0], ids0), flush)
                t_collate = timed(lambda: ds.batch(id_sets[0], ids0, bitmaps=False), flush)
            t_res_step = timed(lambda: trainer.step_resident(ds, id_sets[0], ids0, global_batch), flush)
            gather_bytes = 8 * e + 4 * (n + 1) * 2 + 8 * n + 8 * n * cfg.num_features   # read + write
            resident = {"value": global_batch * e2e_steps / res_s, "unit": UNIT,
                        "h2d_bytes_per_step": 4 * bs, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                        "driver_epoch_value": len(ds) * epochs / epoch_s, "driver_epochs": epochs,
                        "driver_epoch_note": "dgcnn_b200.driver.train_epoch over the resident data set "
                                             f"({len(ds)} graphs, {len(ds) // bs} steps per epoch, shuffled): "
                                             "one host sync per epoch",
                        "device_step_us": t_res_step * 1e6,
                        "device_value": global_batch / t_res_step,
                        "collate_us": t_collate * 1e6, "collate_plus_bitmaps_us": t_gather * 1e6,
                        "collate_algorithmic_bytes": gather_bytes,
                        "collate_GBps": gather_bytes / t_collate / 1e9,
                        "collate_frac_of_peak": gather_bytes / t_collate / 1e9 / peak,
                        "dataset_bytes_in_hbm": ds.nbytes(), "dataset_graphs": len(ds),
                        "note": "data set resident in HBM (DeviceDataset); per step: pinned int32 graph ids -> H2D "
                                "-> FusedTrainer.step_resident (dgcnn_collate gather instead of host collate + "
                                "H2D of the batch + K0, then K0b .. Adam) -> loss.item(); bit-identical to the "
                                "host-fed step (tests/test_gpu_resident.py).  NOT `e2e`: the batch tensors never "
                                "cross PCIe"}
        except Exception as exc:                       # noqa: BLE001  (never lose the main line)
            resident = {"error": repr(exc)}
            print(f"[bench] resident-data-set leg failed: {exc!r}", file=sys.stderr)

    cpu = None
    if not args.no_cpu_baseline:
        cb = make_batch(args.workload, seed=324)
        v, nsteps, el = time_cpu(cfg, cb, args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{nsteps} train steps in {el:.1f} s on one {args.workload}-synth batch "
                         f"({cb.num_graphs} graphs, N={cb.num_nodes}, E={cb.num_edges}), "
                         "CPU oracle restatement, all host threads"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}-synth bs{cfg.batch_size}/GPU k{cfg.k} "
                               f"F{cfg.num_features} (BASELINE.json configs[3])",
                   "nodes_per_batch": n, "edges_per_batch": e, "global_batch": global_batch,
                   "step": "K0 graph build + GraphConv x4 + SortPool + dense tail + NLL + backward "
                           "+ grad all-reduce (N>1) + Adam (all hand-written kernels except NLL)",
                   "l2": f"flushed ({L2_FLUSH_BYTES >> 20} MiB write) before every timed step; "
                         f"ring of {RING} distinct batches",
                   "cuda_graph": use_graph, "fused_trainer": fused_step,
                   "conv5_fused_into_graph_kernels": bool(fused_step and ops.conv5_fusable(cfg.num_features, db0.max_nodes)),
                   "parallelism": f"dp{world} (graph-sharded)",
                   "gradient_exchange": comm_kind,
                   "wall_s_incl_flush": wall},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "note": "pinned host batch -> H2D into a preallocated device slot (copy stream, overlapped with the previous step) -> FusedTrainer.step (K0 .. Adam) -> loss.item()"},
        "e2e_int32_indices": {"value": e2e_compact_value, "unit": UNIT, "h2d_bytes_per_step": h2d_compact_bytes,
                              "d2h_bytes_per_step": 4, "steps": e2e_steps,
                              "note": "same loop, host batch collated with int32 edge_index/batch "
                                      "(dgcnn_build_graph_i32): NOT the reference's int64 format; `e2e` is"},
        "e2e_resident_dataset": resident if world == 1 else resident_multi,
        "first_timed_step_ms": first_step_ms, "ms_per_step_after_first": steady_ms,
        "per_rank": ranks_report,
        "params_equal_across_ranks": params_equal, "comm_status_per_rank": comm_status_all,
        "shards": ("balanced (dp.balanced_shards of a global batch of bs*N graphs)" if balanced else
                   ("every rank the same batches (diagnostic: separates batch variance from system effects)"
                    if args.shards == "same" else "independent draws per rank")) if world > 1 else "n/a",
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": roofline,
        "hot_path_fwd": hot_fwd,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)                                   # see the note at the rank != 0 exit


if __name__ == "__main__":
    main()
