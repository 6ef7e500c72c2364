"""Multi-GPU check of graph-sharded data parallelism on the resident path: every rank
holds the data set in HBM, a step's global batch of graph ids is split with DeviceDataset.shard
(cost-balanced: the ranks get DIFFERENT numbers of graphs) and every rank runs
dgcnn_train_step_resident on its shard with the fused peer-memory exchange + Adam.  After every
step the parameters must (a) be bit-identical on all ranks and (b) equal a single-process trainer
that ran the whole batch, up to fp32 summation order.  An oversize batch goes through
step_autograd on every rank; a rank with an empty shard still joins the exchange.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29513 scripts/check_dp_resident.py"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import dgcnn_b200 as dg
from dgcnn_b200.synth import CONFIGS, make_graphs

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

cfg = CONFIGS["proteins"]
graphs = make_graphs(cfg, 96, seed=7)
ds = dg.DeviceDataset(graphs, dev, num_classes=cfg.num_classes)
torch.manual_seed(324)
model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).eval()      # eval: no dropout streams to align
ref_model = copy.deepcopy(model)
trainer = dg.FusedTrainer(model)
ref = dg.FusedTrainer(ref_model, distributed=False)
assert trainer.exchange is not None, "peer mapping failed"
rng = np.random.RandomState(1)
worst = 0.0
for step in range(6):
    ids = rng.permutation(96)[:40] if step != 4 else rng.permutation(96)[:max(1, world - 1)]   # step 4: fewer graphs than ranks
    mine = ds.shard(ids, world, rank)
    counts = [len(ds.shard(ids, world, r)) for r in range(world)]
    assert sum(counts) == len(ids)
    if len(mine) == 0:
        trainer.step_autograd(None, global_batch=len(ids))        # empty shard: zero gradients, same exchange
    elif step % 2 == 0:
        trainer.step_resident(ds, mine, global_batch=len(ids))
    else:
        trainer.step_autograd(ds.batch(mine), global_batch=len(ids))
    ref.step_resident(ds, ids, global_batch=len(ids))
    torch.cuda.synchronize()
    trainer.check_status()
    first = trainer.flat.clone()
    dist.broadcast(first, src=0)
    assert torch.equal(first, trainer.flat), f"step {step}: ranks hold different parameters"
    diff = (trainer.flat - ref.flat).abs().max().item()
    worst = max(worst, diff)
    assert diff <= 2e-5, f"step {step}: sharded {counts} vs single process: {diff}"
    ref.flat.copy_(trainer.flat); ref.exp_avg.copy_(trainer.exp_avg); ref.exp_avg_sq.copy_(trainer.exp_avg_sq)
    loss_all = float(trainer.stats[0]) / len(ids)
    assert abs(loss_all - float(ref.stats[0]) / len(ids)) <= 1e-4 * max(1.0, abs(loss_all))
# global_batch is mandatory on several ranks
try:
    trainer.step_resident(ds, ds.shard(np.arange(40), world, rank))
    raise SystemExit("global_batch=None was accepted on several ranks")
except ValueError:
    pass
if rank == 0:
    print(f"graph-sharded resident training on {world} GPUs == single process (max parameter diff {worst:.2e}); "
          f"ranks bit-identical; unequal shards, empty shard and autograd fallback covered", flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
