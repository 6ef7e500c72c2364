"""N>1 host logic on CPU (gloo, world_size 2): graph-sharding + the single flat
gradient all-reduce (dgcnn_b200/dp.py) reproduce the single-process gradient of the
global-batch mean NLL.  The per-rank compute here is the ORACLE model (tests may use
it); on the GPU box the same GradBucket wraps dgcnn_b200.Model (bench.py --gpus N)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import dgcnn_oracle as orc
        from dgcnn_b200 import GradBucket, shard_bounds
        from dgcnn_b200.synth import CONFIGS, collate, make_graphs
        torch.set_num_threads(1)
        cfg = CONFIGS["mutag"]
        graphs = make_graphs(cfg, 12, seed=5)
        costs = [g["x"].shape[0] + g["edge_index"].shape[1] for g in graphs]
        lo, hi = shard_bounds(costs, world)[rank]
        torch.manual_seed(324)                                   # replicated parameters
        model = orc.OracleModel(cfg.num_features, cfg.num_classes, cfg.k).eval()
        bucket = GradBucket(model.parameters(), extra=2)
        bucket.zero_()
        local = collate(graphs[lo:hi])
        logp = model(local)
        loss_sum = torch.nn.functional.nll_loss(logp, local.y, reduction="sum")
        loss_sum.backward()
        bucket.extra[0] = loss_sum.detach()
        bucket.extra[1] = (logp.argmax(1) == local.y).sum()
        bucket.all_reduce(global_batch=len(graphs))
        if rank == 0:
            torch.save({"flat": bucket.flat.clone(), "bounds": shard_bounds(costs, world)}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    out_path = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    got = torch.load(out_path)

    from oracle import dgcnn_oracle as orc
    from dgcnn_b200.synth import CONFIGS, collate, make_graphs
    cfg = CONFIGS["mutag"]
    graphs = make_graphs(cfg, 12, seed=5)
    torch.manual_seed(324)
    model = orc.OracleModel(cfg.num_features, cfg.num_classes, cfg.k).eval()
    full = collate(graphs)
    logp = model(full)
    torch.nn.functional.nll_loss(logp, full.y).backward()        # mean over the GLOBAL batch
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    flat = got["flat"]
    assert got["bounds"][0][0] == 0 and got["bounds"][-1][1] == 12
    assert all(hi > lo for lo, hi in got["bounds"])
    torch.testing.assert_close(flat[:-2], want, rtol=1e-4, atol=1e-6)
    loss_sum = torch.nn.functional.nll_loss(logp, full.y, reduction="sum")
    torch.testing.assert_close(flat[-2], loss_sum.detach(), rtol=1e-5, atol=1e-6)
    assert int(flat[-1]) == int((logp.argmax(1) == full.y).sum())


def test_grad_bucket_aliases_parameter_grads():
    from dgcnn_b200 import GradBucket
    lin = torch.nn.Linear(3, 2)
    bucket = GradBucket(lin.parameters(), extra=2)
    lin(torch.ones(4, 3)).sum().backward()
    assert bucket.flat.numel() == 3 * 2 + 2 + 2
    assert torch.equal(bucket.flat[:6].view(2, 3), lin.weight.grad)
    assert bucket.flat[:6].abs().sum() > 0
    bucket.all_reduce(global_batch=4)                            # no process group: scale only
    torch.testing.assert_close(lin.bias.grad, torch.ones(2))
    bucket.zero_()
    assert lin.weight.grad.abs().sum() == 0


def _resident_worker(rank, world, port, out_path):
    """The resident-data-set flavour (SURVEY 8e + 8f N1): every rank holds the whole data set,
    a step's batch is a list of graph ids, each rank trains on its shard_ids slice."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from oracle import dgcnn_oracle as orc
        from dgcnn_b200 import GradBucket, shard_ids
        from dgcnn_b200.synth import CONFIGS, collate, make_graphs
        torch.set_num_threads(1)
        cfg = CONFIGS["mutag"]
        graphs = make_graphs(cfg, 30, seed=9)                    # the "data set", replicated
        nodes = np.array([g["x"].shape[0] for g in graphs])
        edges = np.array([g["edge_index"].shape[1] for g in graphs])
        ids = np.random.RandomState(1).permutation(30)[:14]      # one shuffled batch, same on all ranks
        mine = shard_ids(ids, nodes, edges, world, rank)
        torch.manual_seed(324)
        model = orc.OracleModel(cfg.num_features, cfg.num_classes, cfg.k).eval()
        bucket = GradBucket(model.parameters(), extra=2)
        bucket.zero_()
        local = collate([graphs[int(i)] for i in mine])
        logp = model(local)
        loss_sum = torch.nn.functional.nll_loss(logp, local.y, reduction="sum")
        loss_sum.backward()
        bucket.extra[0] = loss_sum.detach()
        bucket.extra[1] = (logp.argmax(1) == local.y).sum()
        bucket.all_reduce(global_batch=len(ids))
        gathered = [None] * world
        dist.all_gather_object(gathered, [int(i) for i in mine])
        if rank == 0:
            torch.save({"flat": bucket.flat.clone(), "ids": ids.tolist(), "parts": gathered}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_resident_id_sharding_matches_single_process(tmp_path):
    out_path = str(tmp_path / "rank0.pt")
    mp.spawn(_resident_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    got = torch.load(out_path)
    assert got["parts"][0] + got["parts"][1] == got["ids"] and all(len(p) > 0 for p in got["parts"])

    from oracle import dgcnn_oracle as orc
    from dgcnn_b200.synth import CONFIGS, collate, make_graphs
    cfg = CONFIGS["mutag"]
    graphs = make_graphs(cfg, 30, seed=9)
    torch.manual_seed(324)
    model = orc.OracleModel(cfg.num_features, cfg.num_classes, cfg.k).eval()
    full = collate([graphs[i] for i in got["ids"]])
    logp = model(full)
    torch.nn.functional.nll_loss(logp, full.y).backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    torch.testing.assert_close(got["flat"][:-2], want, rtol=1e-4, atol=1e-6)
    assert int(got["flat"][-1]) == int((logp.argmax(1) == full.y).sum())
