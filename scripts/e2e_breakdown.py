"""Where does the end-to-end step (bench.py `e2e`) spend its time?  H2D alone, eager step +
loss.item() alone on resident batches, and both overlapped, for int64 and int32 indices."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgcnn_b200 as dg
from dgcnn_b200.synth import CONFIGS, make_batch

dev = torch.device("cuda:0")
cfg = CONFIGS["collab"]
hbs = [make_batch("collab", seed=324 + i).pin_memory() for i in range(4)]
for hb in hbs:
    hb.max_nodes = int((hb.ptr[1:] - hb.ptr[:-1]).max())
cbs = []
for hb in hbs:
    cb = hb.compact().pin_memory(); cb.max_nodes = hb.max_nodes; cbs.append(cb)
torch.manual_seed(324)
model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).train()
trainer = dg.FusedTrainer(model, lr=1e-3)
K = 30

def t_h2d(batches):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(K):
        d = batches[i % 4].to(dev, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e6

def t_step(batches, item=True):
    devb = []
    for hb in batches:
        d = hb.to(dev); d.max_nodes = hb.max_nodes; devb.append(d)
    for i in range(3): trainer.step(devb[i % 4])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(K):
        s = trainer.step(devb[i % 4])
        if item: float(s[0].item())
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e6

def t_cpu_only(batches):
    devb = []
    for hb in batches:
        d = hb.to(dev); d.max_nodes = hb.max_nodes; devb.append(d)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(K):
        trainer.step(devb[i % 4])
    t1 = time.perf_counter()          # launches issued, GPU not waited for
    torch.cuda.synchronize()
    return (t1 - t0) / K * 1e6

for tag, b in (("int64", hbs), ("int32", cbs)):
    print(tag, "bytes/step", b[0].nbytes(), "H2D only us/step %.0f" % t_h2d(b),
          "| step+item us %.0f" % t_step(b), "| step no item us %.0f" % t_step(b, False),
          "| CPU issue time us %.0f" % t_cpu_only(b), flush=True)
