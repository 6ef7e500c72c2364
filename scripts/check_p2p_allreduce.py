"""2-GPU check of the fused peer-memory all-reduce + Adam kernel (csrc/allreduce_adam.cu)
against NCCL all-reduce + flat Adam.  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 scripts/check_p2p_allreduce.py

With two ranks both paths add the same two numbers; the two Adam kernels may contract their
FMAs differently, so parameters are compared to 1e-7.  Across RANKS the fused kernel must give
bit-identical parameters (fixed summation order)."""
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

name = "proteins"
cfg = CONFIGS[name]
batches = []
for i in range(4):
    hb = make_batch(name, seed=324 + 1000 * rank + i)
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
gb = cfg.batch_size * world
for step in range(6):
    sa = tr_a.step(batches[step % 4], gb).clone()
    sb = tr_b.step(batches[step % 4], gb).clone()
    torch.cuda.synchronize()
    assert int(tr_a.comm_status.item()) == 0, "peer all-reduce timed out"
    assert torch.equal(sa, sb), (step, sa, sb)
    diff = (tr_a.flat - tr_b.flat).abs().max().item()
    assert diff < 1e-7, (step, diff)
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
for _ in range(3):
    g.replay()
    tr_b.step(batches[1], gb)
torch.cuda.synchronize()
assert int(tr_a.comm_status.item()) == 0
diff = (tr_a.flat - tr_b.flat).abs().max().item()
assert diff < 1e-6, diff
if rank == 0:
    print(f"p2p all-reduce + Adam == NCCL all-reduce + Adam on {world} GPUs (max diff {diff})", flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
