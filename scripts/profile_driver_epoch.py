"""Where does the host time of driver.train_epoch go?  perf_counter around its phases (no extra syncs),
COLLAB-synth, 4608 graphs resident, 9 steps per epoch.
    python scripts/profile_driver_epoch.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import dgcnn_b200 as dg
from dgcnn_b200 import driver as drv
from dgcnn_b200.synth import CONFIGS, make_graphs

dev = torch.device("cuda:0")
cfg = CONFIGS["collab"]
graphs = []
for i in range(9):
    graphs += make_graphs(cfg, cfg.batch_size, seed=324 + i)
ds = dg.DeviceDataset(graphs, dev, num_classes=cfg.num_classes)
torch.manual_seed(324)
model = dg.Model(cfg.num_features, cfg.num_classes, cfg.k).to(dev).train()
trainer = dg.FusedTrainer(model, lr=1e-3)
ids = np.arange(len(ds), dtype=np.int64)
gen = torch.Generator().manual_seed(1)
bs = cfg.batch_size
for _ in range(3):
    drv.train_epoch(trainer, ds, ids, bs, gen)
torch.cuda.synchronize()

T = {k: 0.0 for k in ("randperm", "plan", "ids_h2d", "stats_alloc", "steps_host", "result")}
epochs = 20
t_all = time.perf_counter()
for _ in range(epochs):
    t0 = time.perf_counter()
    order = ids[torch.randperm(ids.size, generator=gen).numpy()]
    t1 = time.perf_counter()
    starts, n_b, e_b, mx_b = drv.epoch_plan(order, ds.nodes, ds.edges, bs)
    t2 = time.perf_counter()
    ids_dev = ds.ids_to_device_pinned(order)
    t3 = time.perf_counter()
    st = drv.EpochStats(dev, len(starts))
    t4 = time.perf_counter()
    for i, lo in enumerate(starts):
        hi = min(lo + bs, order.size)
        stats = trainer.step_resident(ds, order[lo:hi], ids_device=ids_dev[lo:hi],
                                      plan=(int(n_b[i]), int(e_b[i]), int(mx_b[i])))
        st.add(stats, hi - lo)
    t5 = time.perf_counter()
    st.result(trainer)
    t6 = time.perf_counter()
    for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
        T[k] += v
total = time.perf_counter() - t_all
print(f"{epochs} epochs of {len(starts)} steps: {total / epochs * 1e6:.0f} us per epoch "
      f"({len(ds) * epochs / total / 1e6:.3f} M graphs/s)")
for k, v in T.items():
    print(f"  {k:12s} {v / epochs * 1e6:8.1f} us per epoch")
print(f"  steps_host per step {T['steps_host'] / epochs / len(starts) * 1e6:.1f} us")
# the same loop through the product function, for reference
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(epochs):
    drv.train_epoch(trainer, ds, ids, bs, gen)
torch.cuda.synchronize()
t = time.perf_counter() - t0
print(f"driver.train_epoch: {t / epochs * 1e6:.0f} us per epoch ({len(ds) * epochs / t / 1e6:.3f} M graphs/s)")
import ctypes
from dgcnn_b200 import _lib
cnt = (ctypes.c_int64 * 4)()
_lib.load_library().dgcnn_train_step_graph_counts(ctypes.cast(cnt, ctypes.c_void_p))
print("graph-replayed steps: updated in place / instantiated / eager / capture failed =", list(cnt))
# host time of one graphed call (capture + update + launch), device kept busy: no sync inside
gs = trainer.graph_stream()
gs.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(gs):
    order = ids.copy()
    starts, n_b, e_b, mx_b = drv.epoch_plan(order, ds.nodes, ds.edges, bs)
    ids_dev = ds.ids_to_device_pinned(order)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(5):
        for i, lo in enumerate(starts):
            trainer.step_resident(ds, order[lo:lo + bs], ids_device=ids_dev[lo:lo + bs],
                                  plan=(int(n_b[i]), int(e_b[i]), int(mx_b[i])), graphed=True)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print(f"graphed step: host {1e6 * (t1 - t0) / (5 * len(starts)):.1f} us per call, "
      f"{1e6 * (t2 - t0) / (5 * len(starts)):.1f} us per step end to end")
