"""2-GPU check of the fused peer-memory all-reduce + Adam kernel (csrc/allreduce_adam.cu)
against NCCL all-reduce + flat Adam.  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 scripts/check_p2p_allreduce.py

Every step starts both trainers from the same state.  With two ranks both paths add the same two
numbers, so the reduced gradients must match bit for bit; the two Adam kernels may contract
their FMAs differently, so the updated parameters are compared to 1e-6.  Across RANKS the fused
kernel must give bit-identical parameters (fixed summation order)."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import dgcnn_b200 as dg
from dgcnn_b200.synth import CONFIGS, make_batch

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

name = "collab"                      # (every graph fits the fused kernels: both trainers take the same path)
cfg = CONFIGS[name]
batches = []
for i in range(4):
    hb = make_batch(name, seed=324 + 1000 * rank + i, num_graphs=64)
    db = hb.to(dev)
    db.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
    batches.append(db)

torch.manual_seed(324)
model_a = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).train()
model_b = copy.deepcopy(model_a)
os.environ["DGCNN_ALLREDUCE"] = "p2p"
tr_a = dg.FusedTrainer(model_a, lr=1e-3)
os.environ["DGCNN_ALLREDUCE"] = "nccl"
tr_b = dg.FusedTrainer(model_b, lr=1e-3)
assert tr_a.exchange is not None, "peer mapping failed"
assert tr_b.exchange is None
gb = 64 * world


def sync_state():
    """Trainer b continues from trainer a's exact state: the comparison is per step (Adam turns a
    1e-9 difference in a near-zero gradient into an lr-sized difference in the parameter, so
    trajectories of several steps are not comparable to a tight tolerance)."""
    tr_b.flat.copy_(tr_a.flat)
    tr_b.exp_avg.copy_(tr_a.exp_avg)
    tr_b.exp_avg_sq.copy_(tr_a.exp_avg_sq)
    tr_b.step_count.copy_(tr_a.step_count)
    model_b._tail_rng_offset.copy_(model_a._tail_rng_offset)


def compare(tag):
    assert int(tr_a.comm_status.item()) == 0, "peer all-reduce timed out"
    # the two sums of W = 2 numbers are the same additions: bit-identical reduced gradients
    gd = (tr_a.grad - tr_b.grad).abs().max().item()
    assert (gd == 0.0) if world == 2 else gd <= 1e-5 * tr_b.grad.abs().max().item(), (tag, gd)
    # one Adam step from identical state and (near-)identical gradients
    pd = (tr_a.flat - tr_b.flat).abs().max().item()
    assert pd <= 1e-6, (tag, pd)


for step in range(6):
    sync_state()
    tr_a.step(batches[step % 4], gb)
    tr_b.step(batches[step % 4], gb)
    torch.cuda.synchronize()
    compare(step)
# ranks hold identical parameters
mine = tr_a.flat.clone()
ref = mine.clone()
dist.broadcast(ref, src=0)
assert torch.equal(mine, ref)
# CUDA-graph capture of a step with the fused kernel
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    tr_a.step(batches[0], gb)
    tr_b.step(batches[0], gb)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    tr_a.step(batches[1], gb)
for it in range(3):
    sync_state()
    g.replay()
    tr_b.step(batches[1], gb)
    torch.cuda.synchronize()
    compare(f"graph replay {it}")
diff = (tr_a.flat - tr_b.flat).abs().max().item()
if rank == 0:
    print(f"p2p all-reduce + Adam == NCCL all-reduce + Adam on {world} GPUs (max diff {diff})", flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
