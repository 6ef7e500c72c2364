#!/bin/bash
# Round 2, the last call: KSB with the CTA's running gradient sum in its HBM slot (15 KB more shared memory for teams)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
timeout 200 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 100 python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
import dgcnn_b200 as dg
from bench import timed
from dgcnn_b200 import ops
from dgcnn_b200.synth import CONFIGS, make_batch
dev = torch.device("cuda:0"); cfg = CONFIGS["collab"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for seed in (324, 325, 326, 327):
    hb = make_batch("collab", seed=seed); data = hb.to(dev)
    data.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
    torch.manual_seed(324)
    model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).train()
    convs = (model.conv1, model.conv2, model.conv3, model.conv4)
    ws, bs = [c.lin.weight for c in convs], [c.bias for c in convs]
    with torch.enable_grad():
        g_t = model.build_graph(data)
    with torch.no_grad():
        h1, arg, xcat, perm, _ = ops.stack_fwd_conv5(data.x, g_t, ws, bs, model.conv5.weight, model.conv5.bias, cfg.k, 0)
        dh1 = torch.randn_like(h1)
        t = timed(lambda: ops.stack_bwd_conv5(dh1, arg, perm, xcat, data.x, g_t, ws, model.conv5.weight, cfg.k, 0), flush, reps=10)
    tr = dg.FusedTrainer(model, lr=1e-3)
    ts = timed(lambda: tr.step(data), flush, reps=10)
    print(f"seed {seed} largest {data.max_nodes}: KSB-conv5 {t*1e6:6.1f} us  step {ts*1e6:6.1f} us", flush=True)
PY
