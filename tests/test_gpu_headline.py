"""GPU parity tests of the HEADLINE path at BASELINE.json's full sizes ("parity
first"): the fused kernels KS / KSB and FusedTrainer.step on COLLAB-synth bs512 against the
float64 oracle, configs[2] D&D and configs[4] power-law at full size, determinism with
poisoned buffers, and the drop-in `Model(data)` / `GCNConv(x, edge_index)` calls of the
reference (model.py:27-35, train.py:36-37) reaching the fused kernels without any hint.

Tolerances: x_cat / pooled within 1e-5 absolute of the float64 oracle; permutation equal to
the oracle's wherever its sorted float64 keys are further apart than 1e-5; gradients within
2e-3 of the largest gradient entry of the tensor (sums of ~1e5 fp32 terms); CSR bit-exact.
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import dgcnn_b200 as dg
from dgcnn_b200 import ops
from dgcnn_b200.synth import CONFIGS, make_batch
from oracle import dgcnn_oracle as orc
from test_gpu_parity import assert_perm_matches

ATOL = 1e-5
DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def oracle_and_model(cfg, seed=324, train=False):
    """float64 oracle with non-zero GCN biases and our model loaded with the same values."""
    torch.manual_seed(seed)
    ref = orc.OracleModel(cfg.num_features, cfg.num_classes, cfg.k).double().eval()
    with torch.no_grad():
        for c in (ref.conv1, ref.conv2, ref.conv3, ref.conv4):
            c.bias.uniform_(-0.1, 0.1)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k)
    model.load_state_dict({k_: v.float() for k_, v in ref.state_dict().items()})
    model = model.to(DEV)
    return ref, (model.train() if train else model.eval())


def oracle_pooled_from_perm(rx, perm, count, k):
    """The oracle's pooled output continued from OUR (validated) permutation: near-tied keys
    may legitimately swap ranks, the gather itself is exact."""
    pcpu = perm.cpu().long()
    rpool = torch.where((pcpu >= 0).unsqueeze(-1), rx[pcpu.clamp(min=0)], rx.new_zeros(()))
    return rpool.reshape(count, k * 97)


def check_forward(cfg, batch, model, ref, expect_fused):
    data = batch.to(DEV)
    before = ops.LAUNCHES["stack_fwd"]
    g = model.build_graph(data)
    pooled, xcat, perm = model.hot_path(data.x, g)
    g.check()
    assert (ops.LAUNCHES["stack_fwd"] > before) == expect_fused, "unexpected kernel path"
    count = batch.num_graphs
    rx, _ = ref.hot_path(batch.x.double(), batch.edge_index, batch.batch, count)
    err = (xcat.detach().cpu().double() - rx.detach()).abs().max().item()
    assert err <= ATOL, f"x_cat differs from the float64 oracle by {err:.3e}"
    _, rperm = orc.sort_aggregation(rx.detach(), batch.batch, cfg.k, count, return_perm=True)
    assert_perm_matches(perm.cpu().numpy(), rperm.numpy(), rx.detach().numpy()[:, -1], batch.ptr.numpy(), cfg.k)
    rpool = oracle_pooled_from_perm(rx, perm, count, cfg.k)
    assert (pooled.detach().cpu().double() - rpool.detach()).abs().max().item() <= ATOL
    # size-independent properties: sorted keys, zero padding, rows are x_cat rows of the own graph
    keys = pooled.view(count, cfg.k, 97)[:, :, -1]
    valid = perm >= 0
    both = valid[:, 1:] & valid[:, :-1]
    assert (keys[:, :-1][both] >= keys[:, 1:][both]).all()
    assert (pooled.view(count, cfg.k, 97)[~valid] == 0).all()
    gp = g.gptr.long()
    lo, hi = gp[:-1].view(-1, 1), gp[1:].view(-1, 1)
    assert ((perm >= lo) & (perm < hi))[valid].all()
    assert torch.equal(pooled.view(-1, 97)[valid.view(-1)], xcat[perm[valid].long()])
    return data, rx, rpool, pooled, xcat, perm


def check_gradients(model, ref, data, batch, rpool, rtol=2e-3):
    out = model(data)
    rout = ref.tail(rpool)
    assert (out.detach().cpu().double() - rout.detach()).abs().max().item() <= 1e-4
    model.zero_grad()
    F.nll_loss(out, data.y).backward()
    F.nll_loss(rout, batch.y).backward()
    rp = dict(ref.named_parameters())
    for pname, p in model.named_parameters():
        want = rp[pname].grad
        scale = max(1e-3, float(want.abs().max()))
        err = (p.grad.cpu().double() - want).abs().max().item()
        assert err <= rtol * scale, f"{pname}: {err:.3e} vs scale {scale:.3e}"


# ---------------------------------------------------------------- BASELINE configs at full size
def test_collab_full_size_fused_forward_backward_vs_float64_oracle():
    """configs[3] COLLAB-synth bs512 (the workload bench.py times) on KS / KSB, no hint on the batch."""
    cfg = CONFIGS["collab"]
    batch = make_batch("collab")                       # the bench's own batch (degree-only features: exact ties)
    ref, model = oracle_and_model(cfg)
    before_b = ops.LAUNCHES["stack_bwd"]
    data, rx, rpool, *_ = check_forward(cfg, batch, model, ref, expect_fused=True)
    check_gradients(model, ref, data, batch, rpool)
    assert ops.LAUNCHES["stack_bwd"] > before_b, "KSB was not used"


def test_dd_full_size_vs_float64_oracle():
    """configs[2] D&D-synth bs64 F90 k291, with its 5748-node graph: forward and backward of the
    path that serves graphs beyond one SM's shared memory."""
    cfg = CONFIGS["dd"]
    batch = make_batch("dd", tie_free=True)
    assert int((batch.ptr[1:] - batch.ptr[:-1]).max()) == 5748
    ref, model = oracle_and_model(cfg)
    data, rx, rpool, *_ = check_forward(cfg, batch, model, ref,
                                        expect_fused=ops.stack_fwd_supported(cfg.num_features, 5748))
    check_gradients(model, ref, data, batch, rpool)


def test_powerlaw_full_size_vs_float64_oracle():
    """configs[4] power-law 1000 nodes x ~10k edges, F64, bs256, k512 (256k nodes, 5.07M edges)."""
    cfg = CONFIGS["powerlaw"]
    batch = make_batch("powerlaw", tie_free=True)
    assert batch.num_nodes == 256000 and batch.num_graphs == 256
    ref, model = oracle_and_model(cfg)
    data, rx, rpool, *_ = check_forward(cfg, batch, model, ref,
                                        expect_fused=ops.stack_fwd_supported(cfg.num_features, 1000))
    check_gradients(model, ref, data, batch, rpool)


# ---------------------------------------------------------------- FusedTrainer.step == what bench.py times
def test_fused_trainer_step_collab_bs512_vs_float64_oracle():
    """One optimisation step of FusedTrainer on COLLAB-synth bs512 (exactly bench.py's step):
    loss, #correct, all 16 parameter gradients and the parameters after Adam against the
    float64 oracle + torch.optim.Adam.  The dropout mask and the (validated) permutation are
    taken from our own forward: the oracle cannot reproduce a counter-hash RNG, and near-tied
    keys may swap ranks."""
    cfg = CONFIGS["collab"]
    batch = make_batch("collab", seed=1324)
    ref, model = oracle_and_model(cfg, train=True)
    ref.train()
    data = batch.to(DEV)
    b, k = batch.num_graphs, cfg.k

    # (a) the step's kernels one by one through the operator layer, on a copy of the model (the
    #     SURVEY 8f N2 sequence when the batch fits it: exactly what the native step launches)
    m_seq = copy.deepcopy(model)
    fused5 = ops.conv5_fusable(cfg.num_features, int((batch.ptr[1:] - batch.ptr[:-1]).max()))
    stats, got, perm, keep_u8, xcat = step_kernels(m_seq, data, k, fused5)
    keep = keep_u8.cpu().double()

    # (b) the oracle in float64 from our permutation and our dropout mask
    rx, _ = ref.hot_path(batch.x.double(), batch.edge_index, batch.batch, b)
    assert (xcat.cpu().double() - rx.detach()).abs().max().item() <= ATOL
    _, rperm = orc.sort_aggregation(rx.detach(), batch.batch, k, b, return_perm=True)
    assert_perm_matches(perm.cpu().numpy(), rperm.numpy(), rx.detach().numpy()[:, -1], batch.ptr.numpy(), k)
    rpool = oracle_pooled_from_perm(rx, perm, b, k)
    h = rpool.view(b, 1, -1)
    h = ref.pool(F.relu(ref.conv5(h)))
    h = F.relu(ref.conv6(h)).flatten(1)
    h = F.relu(ref.classifier_1(h)) * keep                             # Dropout(0.5): `keep` holds OUR multiplier (0 or 2)
    rlogp = F.log_softmax(ref.classifier_2(h), dim=-1)
    rloss = F.nll_loss(rlogp, batch.y, reduction="sum")
    rloss.backward()
    assert abs(float(stats[0]) - float(rloss)) <= 1e-4 * max(1.0, abs(float(rloss)))
    assert float(stats[1]) == float((rlogp.argmax(1) == batch.y).sum())
    names = ["conv1.lin.weight", "conv1.bias", "conv2.lin.weight", "conv2.bias", "conv3.lin.weight", "conv3.bias",
             "conv4.lin.weight", "conv4.bias", "conv5.weight", "conv5.bias", "conv6.weight", "conv6.bias",
             "classifier_1.weight", "classifier_1.bias", "classifier_2.weight", "classifier_2.bias"]
    rp = dict(ref.named_parameters())
    for name, gt in zip(names, got):
        want = rp[name].grad
        scale = max(1e-3, float(want.abs().max()))
        err = (gt.cpu().double().view_as(want) - want).abs().max().item()
        assert err <= 2e-3 * scale, f"{name}: {err:.3e} vs scale {scale:.3e}"

    # (c) FusedTrainer.step (one native call): same gradients bit for bit, and the parameters
    # after its Adam equal torch.optim.Adam on the oracle's mean-loss gradients
    trainer = dg.FusedTrainer(model, lr=1e-3)
    before = ops.LAUNCHES.get("train_step", 0)
    st = trainer.step(data).clone()
    assert ops.LAUNCHES.get("train_step", 0) > before, "the native one-call step was not used"
    assert torch.equal(st, stats)
    flat_seq = torch.cat([t.reshape(-1) for t in got])
    assert torch.equal(trainer.grad[:trainer.num_params], flat_seq), "native step != operator sequence"
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    for p in ref.parameters():
        p.grad.div_(b)                                                  # train.py:39: mean NLL
    opt.step()
    mp = dict(model.named_parameters())
    for name in names:
        err = (mp[name].detach().cpu().double() - rp[name].detach()).abs().max().item()
        assert err <= 2e-5, f"{name} after Adam: {err:.3e}"              # |update| = lr = 1e-3 on step 1


# ---------------------------------------------------------------- determinism with poisoned buffers
@pytest.mark.parametrize("path", ["per-layer", "fused"])
def test_poisoned_buffers_give_bit_identical_results(path):
    """Every output / workspace of the operator layer pre-filled with NaN or 0xFF bytes
    (ops.set_poison): K0 -> K1 x 4 -> K2 (+ K3 / K4) and K0 + K0b -> KS -> KT -> KSB must return
    the same bits as with fresh memory, five times over (a kernel that
    reads memory it did not write shows up here)."""
    cfg = CONFIGS["proteins"]
    batch = make_batch("proteins", num_graphs=40, tie_free=True)
    torch.manual_seed(3)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).eval()
    data = batch.to(DEV)
    dg.set_fused(path == "fused")
    try:
        ref = None
        for kind in (None, "nan", "ff", "nan", None):
            ops.set_poison(kind)
            junk = [ops._empty(s, dtype=torch.float32, device=DEV) for s in (1 << 22, 1 << 20, 1 << 16, 1 << 12)]
            del junk                                     # poisoned blocks go back to the allocator
            model.zero_grad()
            g = model.build_graph(data)
            pooled, xcat, perm = model.hot_path(data.x, g)
            out = model.tail(pooled)
            F.nll_loss(out, data.y).backward()
            torch.cuda.synchronize()
            got = [xcat.detach().clone(), perm.clone(), pooled.detach().clone(), out.detach().clone(),
                   g.rowptr.clone(), g.col[:batch.num_edges].clone(), g.dis.clone(), g.gptr.clone()]
            got += [p.grad.clone() for p in model.parameters()]
            assert not any(torch.isnan(t).any() for t in got if t.is_floating_point())
            if ref is None:
                ref = got
            else:
                for i, (a, b_) in enumerate(zip(ref, got)):
                    assert torch.equal(a, b_), f"poison={kind}: result {i} changed"
    finally:
        ops.set_poison(None)
        dg.set_fused(True)


def test_poisoned_native_train_step_is_bit_identical():
    """The one-call training step (arena-carved buffers) with a NaN-poisoned arena."""
    cfg = CONFIGS["collab"]
    batch = make_batch("collab", num_graphs=64)
    data = batch.to(DEV)
    torch.manual_seed(3)
    base = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).train()
    flats = []
    try:
        for kind in (None, "nan", "ff"):
            ops.set_poison(kind)
            tr = dg.FusedTrainer(copy.deepcopy(base))
            for _ in range(2):
                st = tr.step(data).clone()
            flats.append((tr.flat.clone(), st))
    finally:
        ops.set_poison(None)
    for flat, st in flats[1:]:
        assert torch.equal(flat, flats[0][0]) and torch.equal(st, flats[0][1])
        assert not torch.isnan(flat).any()


# ---------------------------------------------------------------- the reference's own calls
class PlainBatch:
    """What PyG's DataLoader yields, reduced to the attributes model.py:27 and train.py:36 touch;
    no num_graphs, no max_nodes, no ptr."""

    def __init__(self, x, edge_index, batch, y):
        self.x, self.edge_index, self.batch, self.y = x, edge_index, batch, y


def test_drop_in_model_call_reaches_the_fused_kernels_without_hints():
    cfg = CONFIGS["collab"]
    hb = make_batch("collab", num_graphs=128)
    data = PlainBatch(hb.x.to(DEV), hb.edge_index.to(DEV), hb.batch.to(DEV), hb.y.to(DEV))
    ref, model = oracle_and_model(cfg)
    f0, b0 = ops.LAUNCHES["stack_fwd"], ops.LAUNCHES["stack_bwd"]
    out = model(data)                                   # train.py:37
    F.nll_loss(out, data.y).backward()                  # train.py:39-40
    assert ops.LAUNCHES["stack_fwd"] == f0 + 1 and ops.LAUNCHES["stack_bwd"] > b0
    assert data.max_nodes == int((hb.ptr[1:] - hb.ptr[:-1]).max())     # cached: one read per batch
    rx, _ = ref.hot_path(hb.x.double(), hb.edge_index, hb.batch, hb.num_graphs)
    with torch.no_grad():
        _, xcat, _ = model.hot_path(data.x, model.build_graph(data))
    assert (xcat.cpu().double() - rx.detach()).abs().max().item() <= ATOL


def test_reference_forward_body_with_our_modules_builds_the_graph_once():
    """model.py:27-35 verbatim with dgcnn_b200's GCNConv / SortAggregation / remove_self_loops
    (INTEGRATION.md section 1): K0 runs once for the four layers, results match the oracle."""
    cfg = CONFIGS["proteins"]
    hb = make_batch("proteins", num_graphs=32, tie_free=True)
    ref, model = oracle_and_model(cfg)
    x, edge_index, batch = hb.x.to(DEV), hb.edge_index.to(DEV), hb.batch.to(DEV)
    k0, hits = ops.LAUNCHES["build_graph"], ops.GRAPH_CACHE_HITS
    edge_index, _ = dg.remove_self_loops(edge_index)
    x_1 = torch.tanh(model.conv1(x, edge_index))
    x_2 = torch.tanh(model.conv2(x_1, edge_index))
    x_3 = torch.tanh(model.conv3(x_2, edge_index))
    x_4 = torch.tanh(model.conv4(x_3, edge_index))
    xc = torch.cat([x_1, x_2, x_3, x_4], dim=-1)
    pooled = model.sort_pool(xc, batch)
    assert ops.LAUNCHES["build_graph"] - k0 == 3, "K0 must run once per forward, not once per layer"
    assert ops.GRAPH_CACHE_HITS - hits == 3
    rx, _ = ref.hot_path(hb.x.double(), hb.edge_index, hb.batch, hb.num_graphs)
    assert (xc.detach().cpu().double() - rx.detach()).abs().max().item() <= ATOL
    opool = orc.sort_aggregation(xc.detach().cpu(), hb.batch, cfg.k, hb.num_graphs)
    assert torch.equal(pooled.detach().cpu(), opool)
    pooled.sum().backward()
    assert model.conv1.lin.weight.grad is not None and torch.isfinite(model.conv1.lin.weight.grad).all()


# ---------------------------------------------------------------- cluster-pair split of the largest graphs
@pytest.mark.parametrize("name,count", [("collab", 512), ("proteins", 128), ("proteins", 20), ("collab", 3)])
def test_cluster_split_is_bit_identical_to_the_plain_launch(name, count):
    """KS launched as clusters of two CTAs (largest graphs split over a pair, planes exchanged
    through distributed shared memory) against the same kernel launched without clusters:
    x_cat, perm and pooled bit for bit; also with an aggressive split threshold."""
    from dgcnn_b200 import _lib
    lib = _lib.load_library()
    cfg = CONFIGS[name]
    batch = make_batch(name, num_graphs=count, seed=77)
    torch.manual_seed(1)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).eval()
    data = batch.to(DEV)
    outs = []
    try:
        for pairs, pct in ((0, 80), (1, 80), (1, 20)):
            lib.dgcnn_stack_fwd_configure(pairs, pct)
            with torch.no_grad():
                g = model.build_graph(data)
                pooled, xcat, perm = model.hot_path(data.x, g)
            g.check()
            torch.cuda.synchronize()
            outs.append((pooled.clone(), xcat.clone(), perm.clone()))
    finally:
        lib.dgcnn_stack_fwd_configure(-1, 80)
    for got in outs[1:]:
        for a, b_ in zip(outs[0], got):
            assert torch.equal(a, b_)


# ---------------------------------------------------------------- conv5 + ReLU + max-pool inside KS
@pytest.mark.parametrize("name,count,seed", [("collab", 512, 324), ("proteins", 128, 5), ("mutag", 50, 7),
                                             ("collab", 5, 9)])
def test_fused_conv5_head_matches_oracle_and_the_unfused_kernels(name, count, seed):
    """dgcnn_stack_fwd_conv5: h1 / arg from the fused kernel against (a) conv5 -> ReLU -> MaxPool1d
    of the float64 oracle on OUR pooled rows and (b) dgcnn_tail_fwd's own conv5 kernel; x_cat,
    perm and pooled bit-identical to plain KS."""
    cfg = CONFIGS[name]
    batch = make_batch(name, num_graphs=count, seed=seed)
    ref, model = oracle_and_model(cfg, seed=seed)
    data = batch.to(DEV)
    convs = (model.conv1, model.conv2, model.conv3, model.conv4)
    weights, biases = [c.lin.weight for c in convs], [c.bias for c in convs]
    with torch.no_grad():
        g = model.build_graph(data)
        pooled0, xcat0, perm0 = ops.stack_fwd(data.x, g, weights, biases, cfg.k, 0)
        h1, arg, xcat, perm, pooled = ops.stack_fwd_conv5(data.x, g, weights, biases, model.conv5.weight,
                                                          model.conv5.bias, cfg.k, 0, want_pooled=True)
        h1n, argn, xcatn, permn, none = ops.stack_fwd_conv5(data.x, g, weights, biases, model.conv5.weight,
                                                            model.conv5.bias, cfg.k, 0)
    g.check()
    assert none is None
    assert torch.equal(xcat, xcat0) and torch.equal(perm, perm0) and torch.equal(pooled, pooled0)
    assert torch.equal(h1, h1n) and torch.equal(arg, argn) and torch.equal(xcatn, xcat0) and torch.equal(permn, perm0)
    # (a) float64 conv5 + ReLU + max-pool on our pooled rows
    z = F.conv1d(pooled0.cpu().double().view(count, 1, -1), ref.conv5.weight, ref.conv5.bias, stride=97)
    want = F.max_pool1d(F.relu(z), 2, 2)
    assert h1.shape == want.shape
    assert (h1.cpu().double() - want).abs().max().item() <= 2e-5 * max(1.0, float(want.abs().max()))
    zr = F.relu(z)[:, :, : 2 * (cfg.k // 2)].reshape(count, 16, cfg.k // 2, 2)
    clear = (zr[..., 0] - zr[..., 1]).abs() > 1e-4                      # winner not a rounding matter
    warg = torch.where(zr.max(-1).values <= 0, torch.full_like(zr[..., 0], 2.0), (zr[..., 1] > zr[..., 0]).double())
    dead_clear = (z[:, :, : 2 * (cfg.k // 2)].reshape(count, 16, cfg.k // 2, 2).abs() > 1e-4).all(-1)
    sel = clear & dead_clear
    assert torch.equal(arg.cpu().double()[sel], warg[sel])
    # (b) the unfused conv5 kernel of the dense tail
    tail = [model.conv5.weight, model.conv5.bias, model.conv6.weight, model.conv6.bias,
            model.classifier_1.weight, model.classifier_1.bias, model.classifier_2.weight, model.classifier_2.bias]
    with torch.no_grad():
        _, saved = ops.tail_fwd(pooled0, cfg.k, tail, False, 0, None)
    assert (h1 - saved[1]).abs().max().item() <= 2e-5 * max(1.0, float(want.abs().max()))


def step_kernels(model, data, k, fused_conv5, training=True):
    """The training step's kernels one by one through the operator layer: forward, NLL, backward.
    fused_conv5: SURVEY 8f N2 path (stack_fwd_conv5 -> tail from h1 -> tail_bwd_h1 -> stack_bwd_conv5)
    instead of (stack_fwd -> tail -> tail_bwd -> stack_bwd).  Returns (stats, 16 gradients, perm, keep)."""
    convs = (model.conv1, model.conv2, model.conv3, model.conv4)
    weights, biases = [c.lin.weight for c in convs], [c.bias for c in convs]
    tail = [model.conv5.weight, model.conv5.bias, model.conv6.weight, model.conv6.bias,
            model.classifier_1.weight, model.classifier_1.bias, model.classifier_2.weight, model.classifier_2.bias]
    g = model.build_graph(data)
    off = model._tail_rng_offset.clone()
    with torch.no_grad():
        if fused_conv5:
            h1, arg, xcat, perm, _ = ops.stack_fwd_conv5(data.x, g, weights, biases, tail[0], tail[1], k, 0)
            logp, saved = ops.tail_fwd(None, k, tail, training, model._tail_seed, off, h1=h1, arg=arg)
            stats, dlogp = ops.nll_sum(logp, data.y, 1.0, True)
            dh1, tg = ops.tail_bwd_h1(dlogp, logp, saved, k, tail)
            sg = ops.stack_bwd_conv5(dh1, arg, perm, xcat, data.x, g, weights, tail[0], k, 0)
            grads = [t for pair in sg for t in pair] + list(tg)
        else:
            pooled, xcat, perm = ops.stack_fwd(data.x, g, weights, biases, k, 0)
            logp, saved = ops.tail_fwd(pooled, k, tail, training, model._tail_seed, off)
            stats, dlogp = ops.nll_sum(logp, data.y, 1.0, True)
            dpooled, tg = ops.tail_bwd(dlogp, logp, saved, k, tail)
            sg = ops.stack_bwd(dpooled, perm, xcat, data.x, g, weights, k, 0)
            grads = [t for pair in sg for t in pair] + list(tg)
    g.check()
    torch.cuda.synchronize()
    return stats.clone(), [t.clone() for t in grads], perm.clone(), saved[5].clone(), xcat.clone()


PARAM_NAMES = ["conv1.lin.weight", "conv1.bias", "conv2.lin.weight", "conv2.bias", "conv3.lin.weight", "conv3.bias",
               "conv4.lin.weight", "conv4.bias", "conv5.weight", "conv5.bias", "conv6.weight", "conv6.bias",
               "classifier_1.weight", "classifier_1.bias", "classifier_2.weight", "classifier_2.bias"]


@pytest.mark.parametrize("name,count,seed", [("collab", 512, 324), ("proteins", 128, 5), ("mutag", 50, 7),
                                             ("dd", 40, 3), ("collab", 2, 11)])
def test_fused_conv5_backward_matches_the_unfused_kernels_and_the_oracle(name, count, seed):
    """SURVEY 8f N2 end to end: loss, #correct and all 16 parameter gradients of the path that
    never materialises pooled / dpooled against (a) the unfused kernel sequence and (b) the
    float64 oracle continued from our permutation and dropout mask; bit-reproducible."""
    cfg = CONFIGS[name]
    batch = make_batch(name, num_graphs=count, seed=seed, tie_free=(name != "collab"))
    mx = int((batch.ptr[1:] - batch.ptr[:-1]).max())
    if not (ops.stack_fwd_conv5_supported(cfg.num_features, mx) and ops.stack_bwd_conv5_supported(cfg.num_features, mx)):
        keep_ids = [i for i in range(count) if int(batch.ptr[i + 1] - batch.ptr[i]) <= 400]
        from dgcnn_b200.synth import collate, make_graphs
        graphs = make_graphs(cfg, count, seed, tie_free=True)
        batch = collate([graphs[i] for i in keep_ids])
        count = batch.num_graphs
    ref, model = oracle_and_model(cfg, seed=seed, train=True)
    data = batch.to(DEV)
    k = cfg.k
    st_f, gr_f, perm_f, keep_f, xcat_f = step_kernels(model, data, k, True)
    st_u, gr_u, perm_u, keep_u, xcat_u = step_kernels(model, data, k, False)
    assert torch.equal(perm_f, perm_u) and torch.equal(xcat_f, xcat_u)
    assert abs(float(st_f[0]) - float(st_u[0])) <= 1e-4 * max(1.0, abs(float(st_u[0])))
    for pname, a, b_ in zip(PARAM_NAMES, gr_f, gr_u):
        scale = max(1e-3, float(b_.abs().max()))
        if torch.equal(keep_f, keep_u):                  # same dropout decisions: same function
            err = (a.reshape(-1) - b_.reshape(-1)).abs().max().item()
            assert err <= 2e-3 * scale, f"{pname}: fused vs unfused {err:.3e} (scale {scale:.3e})"
    st_f2, gr_f2, *_ = step_kernels(model, data, k, True)
    assert torch.equal(st_f, st_f2) and all(torch.equal(a, b_) for a, b_ in zip(gr_f, gr_f2)), "not reproducible"
    # (b) the float64 oracle from our permutation and our mask
    ref.train()
    rx, _ = ref.hot_path(batch.x.double(), batch.edge_index, batch.batch, count)
    rpool = oracle_pooled_from_perm(rx, perm_f, count, k)
    h = ref.pool(F.relu(ref.conv5(rpool.view(count, 1, -1))))
    h = F.relu(ref.conv6(h)).flatten(1)
    h = F.relu(ref.classifier_1(h)) * keep_f.cpu().double()
    rlogp = F.log_softmax(ref.classifier_2(h), dim=-1)
    rloss = F.nll_loss(rlogp, batch.y, reduction="sum")
    rloss.backward()
    assert abs(float(st_f[0]) - float(rloss.detach())) <= 1e-4 * max(1.0, abs(float(rloss.detach())))
    rp = dict(ref.named_parameters())
    for pname, gt in zip(PARAM_NAMES, gr_f):
        want = rp[pname].grad
        scale = max(1e-3, float(want.abs().max()))
        err = (gt.cpu().double().view_as(want) - want).abs().max().item()
        assert err <= 2e-3 * scale, f"{pname}: vs float64 oracle {err:.3e} (scale {scale:.3e})"


def oracle_gradients_from_our_perm(ref, batch, count, k, perm, keep):
    """loss and the 16 parameter gradients of the float64 oracle continued from OUR permutation
    and OUR dropout mask"""
    ref.train()
    ref.zero_grad()
    rx, _ = ref.hot_path(batch.x.double(), batch.edge_index, batch.batch, count)
    rpool = oracle_pooled_from_perm(rx, perm, count, k)
    h = ref.pool(F.relu(ref.conv5(rpool.view(count, 1, -1))))
    h = F.relu(ref.conv6(h)).flatten(1)
    h = F.relu(ref.classifier_1(h)) * keep.cpu().double()
    rlogp = F.log_softmax(ref.classifier_2(h), dim=-1)
    rloss = F.nll_loss(rlogp, batch.y, reduction="sum")
    rloss.backward()
    return rloss.detach(), dict(ref.named_parameters()), rx.detach()


def collab_batch_with_big_graphs(big_sizes, small, seed):
    """`small` ordinary COLLAB-synth graphs plus one dense ego-net-like graph per entry of big_sizes"""
    from dgcnn_b200.synth import collate, make_graphs, _gnm_pairs, _symmetrise_sorted
    cfg = CONFIGS["collab"]
    graphs = make_graphs(cfg, small, seed=seed)
    rng = np.random.RandomState(seed + 1)
    for j, m in enumerate(big_sizes):
        ei = _symmetrise_sorted(_gnm_pairs(rng, m, 33 * m), m)
        deg = np.bincount(ei[1], minlength=m).astype(np.float32)
        graphs.insert((7 * j) % (len(graphs) + 1), {"x": (deg / deg.max())[:, None].astype(np.float32),
                                                    "edge_index": ei, "y": int(rng.randint(0, 3))})
    return collate(graphs)


@pytest.mark.parametrize("big_sizes,small", [((492,), 511), ((505, 470, 455), 60), ((512,) * 3 + (500,) * 4 + (450,) * 5, 200),
                                             ((497,), 2), ((492, 300), 147)])
def test_graphs_beyond_one_cta_are_split_over_a_cta_pair(big_sizes, small):
    """COLLAB's largest graphs (up to 492 nodes) do not fit one CTA's shared memory in the conv5-fused
    kernels: the plan gives them a CTA PAIR (mandatory split; more such graphs than pairs go round
    by round).  Forward, loss and all 16 gradients of the fused path against the float64 oracle and
    the unfused kernel sequence; bit-reproducible; the training step stays on the fused path (21 launches)."""
    cfg = CONFIGS["collab"]
    batch = collab_batch_with_big_graphs(big_sizes, small, seed=sum(big_sizes))
    count = batch.num_graphs
    mx = int((batch.ptr[1:] - batch.ptr[:-1]).max())
    assert mx == max(big_sizes)
    assert ops.stack_fwd_conv5_supported(cfg.num_features, mx) and ops.stack_bwd_conv5_supported(cfg.num_features, mx)
    ref, model = oracle_and_model(cfg, seed=3, train=True)
    data = batch.to(DEV)
    k = cfg.k
    st_f, gr_f, perm_f, keep_f, xcat_f = step_kernels(model, data, k, True)
    st_u, gr_u, perm_u, keep_u, xcat_u = step_kernels(model, data, k, False)
    assert torch.equal(perm_f, perm_u) and torch.equal(xcat_f, xcat_u)
    st_f2, gr_f2, *_ = step_kernels(model, data, k, True)
    assert torch.equal(st_f, st_f2) and all(torch.equal(a, b_) for a, b_ in zip(gr_f, gr_f2)), "not reproducible"
    rloss, rp, rx = oracle_gradients_from_our_perm(ref, batch, count, k, perm_f, keep_f)
    assert (xcat_f.cpu().double() - rx).abs().max().item() <= ATOL
    _, rperm = orc.sort_aggregation(rx, batch.batch, k, count, return_perm=True)
    assert_perm_matches(perm_f.cpu().numpy(), rperm.numpy(), rx.numpy()[:, -1], batch.ptr.numpy(), k)
    assert abs(float(st_f[0]) - float(rloss)) <= 1e-4 * max(1.0, abs(float(rloss)))
    for pname, gt, gu in zip(PARAM_NAMES, gr_f, gr_u):
        want = rp[pname].grad
        scale = max(1e-3, float(want.abs().max()))
        err = (gt.cpu().double().view_as(want) - want).abs().max().item()
        assert err <= 2e-3 * scale, f"{pname}: vs float64 oracle {err:.3e} (scale {scale:.3e})"
        if torch.equal(keep_f, keep_u):
            erru = (gt.reshape(-1) - gu.reshape(-1)).abs().max().item()
            assert erru <= 2e-3 * scale, f"{pname}: fused vs unfused {erru:.3e}"
    tr = dg.FusedTrainer(model)
    assert tr.supported(data)
    before = ops.LAUNCHES.get("train_step", 0)
    tr.step(data)
    assert ops.LAUNCHES.get("train_step", 0) - before == 21


@pytest.mark.parametrize("seed", range(8))
def test_random_batches_through_the_split_plan(seed):
    """Fuzz of the per-CTA plan (graph_mma.cuh: split set, rounds, toy grids): random batch sizes on
    both sides of the SM count with a random number of graphs beyond one CTA's shared memory.  The
    conv5-fused kernels against the unfused sequence (same permutation and x_cat bit for bit, loss and
    gradients to fp32 rounding), and the cluster launch against the plain one where the plain one can
    hold the batch."""
    from dgcnn_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.RandomState(100 + seed)
    small = int(rng.choice([1, 5, 40, 140, 160, 330]))
    nbig = int(rng.choice([0, 1, 2, 9, 14]))
    hi = 512 if seed % 2 else 432
    big = tuple(int(v) for v in rng.randint(300, hi + 1, size=nbig))
    cfg = CONFIGS["collab"]
    batch = collab_batch_with_big_graphs(big, small, seed=200 + seed)
    mx = int((batch.ptr[1:] - batch.ptr[:-1]).max())
    torch.manual_seed(seed)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).train()
    data = batch.to(DEV)
    assert ops.conv5_fusable(cfg.num_features, mx)
    st_f, gr_f, perm_f, keep_f, xcat_f = step_kernels(model, data, cfg.k, True, training=False)
    st_u, gr_u, perm_u, keep_u, xcat_u = step_kernels(model, data, cfg.k, False, training=False)
    assert torch.equal(perm_f, perm_u) and torch.equal(xcat_f, xcat_u)
    assert abs(float(st_f[0]) - float(st_u[0])) <= 1e-4 * max(1.0, abs(float(st_u[0])))
    for pname, a, b_ in zip(PARAM_NAMES, gr_f, gr_u):
        scale = max(1e-3, float(b_.abs().max()))
        assert (a.reshape(-1) - b_.reshape(-1)).abs().max().item() <= 2e-3 * scale, pname
    if mx <= 432:
        try:
            lib.dgcnn_stack_fwd_configure(0, 80)
            st_p, gr_p, perm_p, _, xcat_p = step_kernels(model, data, cfg.k, True, training=False)
        finally:
            lib.dgcnn_stack_fwd_configure(-1, 80)
        assert torch.equal(perm_p, perm_f) and torch.equal(xcat_p, xcat_f) and torch.equal(st_p, st_f)
        for pname, a, b_ in zip(PARAM_NAMES, gr_p, gr_f):
            scale = max(1e-3, float(b_.abs().max()))
            assert (a - b_).abs().max().item() <= 2e-5 * scale, pname


@pytest.mark.parametrize("name,count", [("collab", 512), ("proteins", 128), ("collab", 3)])
def test_backward_cluster_split_matches_the_plain_launch(name, count):
    """KSB launched as clusters of two CTAs (largest graphs split over a pair: each CTA takes half of
    the row tiles, the gradient rows cross through L2) against the plain launch: the same 16
    gradients up to fp32 summation order (each CTA of a pair leaves its own partial sums)."""
    from dgcnn_b200 import _lib
    lib = _lib.load_library()
    cfg = CONFIGS[name]
    batch = make_batch(name, num_graphs=count, seed=78)
    torch.manual_seed(1)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).train()
    data = batch.to(DEV)
    outs = {}
    try:
        for pairs in (0, 1):
            lib.dgcnn_stack_fwd_configure(pairs, 20)     # (aggressive threshold: several split graphs)
            for fused in (True, False):
                outs[(pairs, fused)] = step_kernels(model, data, cfg.k, fused, training=False)
    finally:
        lib.dgcnn_stack_fwd_configure(-1, 80)
    for fused in (True, False):
        st0, gr0, perm0, _, xcat0 = outs[(0, fused)]
        st1, gr1, perm1, _, xcat1 = outs[(1, fused)]
        assert torch.equal(perm0, perm1) and torch.equal(xcat0, xcat1) and torch.equal(st0, st1)
        for pname, a, b_ in zip(PARAM_NAMES, gr0, gr1):
            scale = max(1e-3, float(a.abs().max()))
            assert (a - b_).abs().max().item() <= 2e-5 * scale, f"{pname} (fused={fused})"


@pytest.mark.parametrize("kind", ["collab", "big", "multigraph", "proteins", "asymmetric"])
def test_lazy_adjacency_maps_give_bit_identical_training_steps(kind):
    """dgcnn_train_step with the forward kernel building (and exporting) its own adjacency maps
    against the same step after the full K0b: losses and parameters bit for bit over three steps --
    plain batches, graphs split over a CTA pair, multigraphs (duplicate edges: CSR walk) and a
    batch K0 finds asymmetric (transposed bitmap from the gated K0b fill)."""
    if kind == "big":
        cfg = CONFIGS["collab"]
        batches = [collab_batch_with_big_graphs((492, 505 - 20 * i, 330), 150, seed=60 + i) for i in range(3)]
    elif kind in ("multigraph", "asymmetric"):
        from dgcnn_b200.synth import collate, make_graphs
        cfg = CONFIGS["collab"]
        batches = []
        for i in range(3):
            graphs = make_graphs(cfg, 40, seed=70 + i)
            rng = np.random.RandomState(i)
            for g_ in graphs[::3]:
                ei = g_["edge_index"]
                if kind == "multigraph":
                    extra = ei[:, rng.randint(0, ei.shape[1], size=max(1, ei.shape[1] // 10))]
                    both = np.concatenate([ei, extra, extra[::-1]], axis=1)
                    order = np.lexsort((both[0], both[1]))
                    g_["edge_index"] = both[:, order]
                else:
                    keep = np.ones(ei.shape[1], dtype=bool)
                    keep[rng.randint(0, ei.shape[1], size=max(1, ei.shape[1] // 8))] = False
                    g_["edge_index"] = ei[:, keep]
            batches.append(collate(graphs))
    else:
        cfg = CONFIGS[kind]
        batches = [make_batch(kind, seed=50 + i, num_graphs=200 if kind == "collab" else 64) for i in range(3)]
    batches = [b_.to(DEV) for b_ in batches]
    torch.manual_seed(4)
    base = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).train()
    results = {}
    try:
        for lazy in (True, False):
            ops.set_lazy_maps(lazy)
            tr = dg.FusedTrainer(copy.deepcopy(base))
            assert tr.supported(batches[0])
            stats = [tr.step(d).clone() for d in batches]
            tr.check_status()
            results[lazy] = (stats, tr.flat.clone())
    finally:
        ops.set_lazy_maps(True)
    for a, b_ in zip(results[True][0], results[False][0]):
        assert torch.equal(a, b_)
    assert torch.equal(results[True][1], results[False][1])


def test_training_with_and_without_the_conv5_fusion_agree():
    """Three optimisation steps of FusedTrainer (native one-call step and the Python sequence) with
    SURVEY 8f N2 on and off: same losses and parameters up to fp32 rounding; the fused step
    launches four kernels fewer."""
    cfg = CONFIGS["collab"]
    batches = [make_batch("collab", seed=40 + i, num_graphs=96).to(DEV) for i in range(3)]
    torch.manual_seed(9)
    base = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).train()
    results = {}
    try:
        for fuse in (True, False):
            ops.set_fuse_conv5(fuse)
            tr = dg.FusedTrainer(copy.deepcopy(base))
            before = ops.LAUNCHES.get("train_step", 0)
            losses = [float(tr.step(d)[0]) for d in batches]
            results[fuse] = (losses, tr.flat.clone(), ops.LAUNCHES.get("train_step", 0) - before)
    finally:
        ops.set_fuse_conv5(True)
    (la, pa, na), (lb, pb, nb) = results[True], results[False]
    assert na == 3 * 21 and nb == 3 * 25                 # (lazy adjacency maps: one K0b launch less)
    for x_, y_ in zip(la, lb):
        assert abs(x_ - y_) <= 1e-4 * max(1.0, abs(y_))
    assert (pa - pb).abs().max().item() <= 2e-5


def test_model_autograd_uses_the_conv5_fusion_and_matches_the_unfused_model():
    """Model(data) + loss.backward() (the reference's train.py:37-40) goes through the N2 autograd
    nodes; gradients equal the unfused autograd path up to fp32 rounding."""
    cfg = CONFIGS["proteins"]
    data = make_batch("proteins", num_graphs=64, seed=3, tie_free=True).to(DEV)
    torch.manual_seed(2)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(DEV).eval()
    grads = {}
    try:
        for fuse in (True, False):
            ops.set_fuse_conv5(fuse)
            model.zero_grad()
            before = ops.LAUNCHES["tail_fwd"]
            out = model(data)
            F.nll_loss(out, data.y).backward()
            grads[fuse] = (out.detach().clone(), [p.grad.clone() for p in model.parameters()])
    finally:
        ops.set_fuse_conv5(True)
    assert (grads[True][0] - grads[False][0]).abs().max().item() <= 1e-5
    for (n_, _), a, b_ in zip(model.named_parameters(), grads[True][1], grads[False][1]):
        scale = max(1e-3, float(b_.abs().max()))
        assert (a - b_).abs().max().item() <= 2e-3 * scale, n_


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("b,k,c", [(512, 130, 3), (37, 30, 2), (5, 61, 2)])
def test_fused_tail_head_equals_the_separate_kernels(b, k, c, training):
    """dgcnn_tail_fwd_loss / dgcnn_tail_bwd_after_loss (fc1 epilogue + fc2 + log_softmax + NLL +
    d(logits) + fc2's row backward in one kernel) against dgcnn_tail_fwd + dgcnn_nll_sum +
    dgcnn_tail_bwd: logp, the saved tensors, dpooled and all eight gradients bit for bit; the loss
    sum to fp32 summation order; the dropout stream advances by one."""
    torch.manual_seed(b + k)
    model = dg.Model(7, c, k).to(DEV)
    tail = [model.conv5.weight, model.conv5.bias, model.conv6.weight, model.conv6.bias,
            model.classifier_1.weight, model.classifier_1.bias, model.classifier_2.weight, model.classifier_2.bias]
    pooled = torch.randn(b, k * 97, device=DEV)
    y = torch.randint(0, c, (b,), device=DEV)
    off_a, off_b = torch.full((1,), 5, dtype=torch.int64, device=DEV), torch.full((1,), 5, dtype=torch.int64, device=DEV)
    with torch.no_grad():
        logp, saved = ops.tail_fwd(pooled, k, tail, training, 1234, off_a)
        stats, dlogp = ops.nll_sum(logp, y, 1.0, True)
        dpooled, grads = ops.tail_bwd(dlogp, logp, saved, k, tail)
        stats2 = torch.zeros(2, device=DEV)
        logp2, saved2, ctx = ops.tail_fwd_loss(pooled, k, tail, y, training, 1234, off_b)
        dpooled2, grads2 = ops.tail_bwd_after_loss(ctx, logp2, saved2, k, tail, stats2, False)
    torch.cuda.synchronize()
    assert torch.equal(logp, logp2) and torch.equal(dpooled, dpooled2)
    for a, b_ in zip(saved[1:], saved2[1:]):
        assert torch.equal(a, b_)
    for a, b_ in zip(grads, grads2):
        assert torch.equal(a, b_)
    assert abs(float(stats[0]) - float(stats2[0])) <= 1e-5 * max(1.0, abs(float(stats[0])))
    assert float(stats[1]) == float(stats2[1])
    assert int(off_a) == int(off_b) == (6 if training else 5)
