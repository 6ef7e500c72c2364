"""The reference's training driver on the B200-native path.

    python -m dgcnn_b200.driver --data_type MUTAG --batch_size 50 --num_epochs 100 --seed 324

Same command line, same 10-fold protocol, same output files as train.py:69-148:

    epochs/{data_type}_{fold}.pth                   model.state_dict() (PyG parameter names)
    statistics/{data_type}_results_{fold}.csv       index 'epoch'; train_loss, test_loss,
                                                    train_accuracy, test_accuracy
    statistics/{data_type}_results_overall.csv      index 'fold'; train_accuracy, test_accuracy

What is different is where the work happens: the data set is parsed once (TU raw text files,
``data.read_tu_dataset``), lives in HBM (``DeviceDataset``), every training step is ONE
library call on a list of graph ids (``FusedTrainer.step_resident``: gather, forward, NLL,
backward, Adam), evaluation gathers its batches on the device too, and loss / accuracy
accumulate ON THE DEVICE and are read once per epoch instead of twice per batch
(train.py:44-45).  visdom plots (train.py:72, 122-125) are replaced by one JSON line per epoch
on stdout.

The TU downloads need PyG + network, neither of which exists here: ``--synthetic`` writes a
TU-format stand-in with the data set's shape (``synth.CONFIGS``) under ``--data_root`` first, and
makes up the fold files when the reference's ``10fold_idx`` is not there.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .data import DeviceDataset, epoch_batches, load_fold, read_tu_dataset, write_tu_dataset
from .nn import Model
from .trainer import FusedTrainer

DATA_TYPES = ["DD", "PTC_MR", "NCI1", "PROTEINS", "IMDB-BINARY", "IMDB-MULTI", "MUTAG", "COLLAB"]
SYNTH_SHAPES = {"MUTAG": "mutag", "PROTEINS": "proteins", "DD": "dd", "COLLAB": "collab"}


def get_args(argv: Optional[Sequence[str]] = None):
    """train.py:16-24, plus where to read / write and the knobs the reference hard-codes."""
    p = argparse.ArgumentParser(description="Train Model")
    p.add_argument("--data_type", default="DD", type=str, choices=DATA_TYPES, help="dataset type")
    p.add_argument("--batch_size", default=50, type=int, help="train batch size")
    p.add_argument("--num_epochs", default=100, type=int, help="train epochs number")
    p.add_argument("--seed", default=324, type=int, help="random seed")
    p.add_argument("--data_root", default="data", type=str, help="holds {data_type}/ (train.py:82)")
    p.add_argument("--out_root", default=".", type=str, help="holds epochs/ and statistics/")
    p.add_argument("--k", default=30, type=int, help="SortPooling k (model.py:17 hard-codes 30)")
    p.add_argument("--folds", default=10, type=int, help="folds to run (train.py:94: 10)")
    p.add_argument("--synthetic", action="store_true",
                   help="write a TU-format synthetic stand-in when the raw files are absent")
    p.add_argument("--synthetic_graphs", default=0, type=int, help="graphs in the stand-in (0: shape default)")
    return p.parse_args(argv)


def set_determ(seed: int) -> None:
    """set_determ.py:1-31: seed python, numpy and torch (all devices)."""
    random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    np.random.seed(seed)


def fold_split(data_root: str, data_type: str, fold: int, num_graphs: int, folds: int = 10,
               seed: int = 324) -> Tuple[np.ndarray, np.ndarray]:
    """train.py:102-105 when the fold files exist; otherwise a seeded permutation cut into
    ``folds`` nearly equal test sets (the stand-in data sets have no published folds)."""
    try:
        return load_fold(data_root, data_type, fold)
    except (OSError, FileNotFoundError):
        perm = np.random.RandomState(seed).permutation(num_graphs)
        bounds = np.linspace(0, num_graphs, folds + 1).astype(np.int64)
        test = np.sort(perm[bounds[fold - 1]:bounds[fold]])
        train = np.sort(np.setdiff1d(perm, test))
        return train.astype(np.int64), test.astype(np.int64)


def ensure_dataset(args) -> Tuple[List[dict], int, int]:
    root = os.path.join(args.data_root, args.data_type)
    try:
        return read_tu_dataset(root, args.data_type)
    except FileNotFoundError:
        if not args.synthetic:
            raise
    from .synth import CONFIGS, make_graphs
    cfg = CONFIGS[SYNTH_SHAPES.get(args.data_type, "mutag")]
    count = args.synthetic_graphs or cfg.num_graphs
    graphs = make_graphs(cfg, count, seed=args.seed)
    raw = os.path.join(root, "raw")
    write_tu_dataset(raw, args.data_type, graphs)
    return read_tu_dataset(root, args.data_type)


class EpochStats:
    """running_loss / correct of train.py:33, 44-45 kept on the device: one copy per step into a
    per-epoch history, one host read per epoch."""

    def __init__(self, device, steps: int = 0):
        self.device = device
        self.hist = torch.zeros(max(steps, 1), 2, dtype=torch.float32, device=device)
        self.graphs = []                                      # graphs per recorded step (host)

    def add(self, stats: torch.Tensor, batch_graphs: int) -> None:
        # stats = [sum of NLL over the batch, #correct]; the reference adds the batch MEAN
        i = len(self.graphs)
        if i >= self.hist.size(0):
            self.hist = torch.cat([self.hist, torch.zeros_like(self.hist)])
        self.hist[i].copy_(stats[:2])                         # (one small launch per step)
        self.graphs.append(int(batch_graphs))

    def result(self, trainer: Optional[FusedTrainer] = None) -> Tuple[float, float]:
        """(mean batch loss, accuracy in %).  With ``trainer``: its status words are read in the
        same device-to-host round trip and checked (FusedTrainer.check_status)."""
        steps = len(self.graphs)
        if steps == 0:
            if trainer is not None:
                trainer.check_status()
            return 0.0, 0.0
        if trainer is not None:
            words = torch.cat([self.hist[:steps].reshape(-1), trainer.status_words()])
            host = words.cpu().double().numpy()              # the epoch's only host sync
            h = host[:2 * steps].reshape(steps, 2)
            trainer.check_status(values=host[2 * steps:])
        else:
            h = self.hist[:steps].cpu().double().numpy()
        g = np.asarray(self.graphs, dtype=np.float64)
        loss = float((h[:, 0] / g).sum()) / steps
        return loss, float(h[:, 1].sum()) / float(g.sum()) * 100.0


def epoch_plan(order: np.ndarray, nodes: np.ndarray, edges: np.ndarray, batch_size: int):
    """(first position, nodes, edges, largest graph) of every batch of an epoch that walks the graph
    ids ``order`` in steps of ``batch_size`` (last batch short): ``DeviceDataset.plan`` for all
    batches in one vectorised pass over the per-graph tables."""
    order = np.asarray(order, dtype=np.int64)
    starts = np.arange(0, order.size, int(batch_size))
    nn, ne = np.asarray(nodes)[order], np.asarray(edges)[order]
    return (starts, np.add.reduceat(nn, starts).astype(np.int64), np.add.reduceat(ne, starts).astype(np.int64),
            np.maximum.reduceat(nn, starts).astype(np.int64))


def train_epoch(trainer: FusedTrainer, ds: DeviceDataset, ids: np.ndarray, batch_size: int,
                generator: torch.Generator) -> Tuple[float, float]:
    """train.py:27-47.  The host side of a step is kept below the device time of the fused step
    (~0.2 ms): the epoch's shuffled ids cross PCIe ONCE (pinned), every batch's sizes come from one
    vectorised pass over the per-graph tables, and a step is one library call plus one small copy."""
    trainer.model.train()
    ids = np.asarray(ids, dtype=np.int64)
    if ids.size == 0:
        return 0.0, 0.0
    if ids.min() < 0 or ids.max() >= ds.num_graphs:
        raise IndexError("train_epoch: graph id outside the data set")
    order = ids[torch.randperm(ids.size, generator=generator).numpy()]   # (= epoch_batches(shuffle=True))
    starts, n_b, e_b, mx_b = epoch_plan(order, ds.nodes, ds.edges, batch_size)
    # DGCNN_GRAPHED_STEP=1: replay the steps as ONE CUDA graph that is updated in place from batch to
    # batch (dgcnn_train_step_resident_graphed; graphs need a stream of their own).  Measured: 44 us
    # of host time per call instead of 108 us and 219 us per step on the device instead of 233 us
    # (eager: 16 launches plus side-stream events), but no gain for a nine-step epoch as a whole
    # (2315 against 2277 us: the epoch's one host read drains the pipeline either way) -- off by default.
    graphed = os.environ.get("DGCNN_GRAPHED_STEP", "0") == "1"
    caller = torch.cuda.current_stream(ds.device)
    stream = trainer.graph_stream() if graphed else caller
    stream.wait_stream(caller)
    with torch.cuda.stream(stream):
        ids_dev = ds.ids_to_device_pinned(order)              # one H2D per epoch (reused pinned buffer)
        st = EpochStats(ds.device, len(starts))
        fits = {}                                             # largest graph -> fused step possible
        for i, lo in enumerate(starts):
            hi = min(lo + batch_size, order.size)
            mx = int(mx_b[i])
            ok = fits.get(mx)
            if ok is None:
                ok = fits[mx] = trainer.resident_supported(ds, None, max_nodes=mx)
            if ok:                                            # one library call: gather .. Adam
                stats = trainer.step_resident(ds, order[lo:hi], ids_device=ids_dev[lo:hi],
                                              plan=(int(n_b[i]), int(e_b[i]), mx), graphed=graphed)
            else:
                # a graph of this batch exceeds the fused kernels (D&D's 5748 nodes, PROTEINS' 620):
                # same step through Model(data) + autograd on the per-layer kernels, same flat Adam
                stats = trainer.step_autograd(ds.batch(order[lo:hi]))
            st.add(stats, hi - lo)
        out = st.result(trainer)                              # + comm timeout / bad input flags, same read
    caller.wait_stream(stream)
    return out


def test_epoch(model: Model, ds: DeviceDataset, ids: np.ndarray, batch_size: int) -> Tuple[float, float]:
    """train.py:49-66."""
    model.eval()
    st = EpochStats(ds.device)
    with torch.no_grad():
        for b in epoch_batches(ids, batch_size, shuffle=False):
            data = ds.batch(b)
            stats, _ = ops.nll_sum(model(data), data.y, 1.0, want_grad=False)
            st.add(stats, len(b))
    return st.result()


test_epoch.__test__ = False          # not a pytest test


def write_fold_csv(path: str, results: Dict[str, List[float]], index_label: str) -> None:
    """pd.DataFrame(data=results, index=range(1, n+1)).to_csv(path, index_label=...)
    (train.py:130-131, 143-144) without the pandas dependency: same header, same rows."""
    cols = list(results)
    n = len(results[cols[0]]) if cols else 0
    with open(path, "w") as fh:
        fh.write(",".join([index_label] + cols) + "\n")
        for i in range(n):
            fh.write(",".join([str(i + 1)] + [repr(float(results[c][i])) for c in cols]) + "\n")


def main(argv: Optional[Sequence[str]] = None) -> Dict[str, List[float]]:
    opt = get_args(argv)
    set_determ(opt.seed)
    if not torch.cuda.is_available():
        raise SystemExit("dgcnn_b200.driver: no CUDA device; the hot path has no CPU fallback")
    device = torch.device("cuda", torch.cuda.current_device())
    graphs, num_features, num_classes = ensure_dataset(opt)
    print(f"data_set.num_features={num_features}, data_set.num_classes={num_classes}", flush=True)
    ds = DeviceDataset(graphs, device, num_classes=num_classes)
    os.makedirs(os.path.join(opt.out_root, "epochs"), exist_ok=True)
    os.makedirs(os.path.join(opt.out_root, "statistics"), exist_ok=True)
    generator = torch.Generator().manual_seed(opt.seed)

    over_results = {"train_accuracy": [], "test_accuracy": []}
    for fold_number in range(1, opt.folds + 1):
        model = Model(num_features, num_classes, k=opt.k).to(device)           # train.py:98
        trainer = FusedTrainer(model)                                          # NLL + Adam defaults
        train_idx, test_idx = fold_split(opt.data_root, opt.data_type, fold_number, len(ds),
                                         max(opt.folds, 2), opt.seed)
        fold_results = {"train_loss": [], "test_loss": [], "train_accuracy": [], "test_accuracy": []}
        for epoch in range(1, opt.num_epochs + 1):
            t0 = time.perf_counter()
            train_loss, train_acc = train_epoch(trainer, ds, train_idx, opt.batch_size, generator)
            test_loss, test_acc = test_epoch(model, ds, test_idx, opt.batch_size)
            fold_results["train_loss"].append(train_loss)
            fold_results["train_accuracy"].append(train_acc)
            fold_results["test_loss"].append(test_loss)
            fold_results["test_accuracy"].append(test_acc)
            print(json.dumps({"fold": fold_number, "epoch": epoch, "train_loss": train_loss,
                              "train_accuracy": train_acc, "test_loss": test_loss,
                              "test_accuracy": test_acc, "epoch_s": time.perf_counter() - t0}), flush=True)
        torch.save(model.state_dict(), os.path.join(opt.out_root, "epochs", f"{opt.data_type}_{fold_number}.pth"))
        write_fold_csv(os.path.join(opt.out_root, "statistics", f"{opt.data_type}_results_{fold_number}.csv"),
                       fold_results, "epoch")
        over_results["train_accuracy"].append(fold_results["train_accuracy"][-1])
        over_results["test_accuracy"].append(fold_results["test_accuracy"][-1])
        print(f"[{fold_number}] Train Acc: {fold_results['train_accuracy'][-1]:.2f}% "
              f"Test Acc: {fold_results['test_accuracy'][-1]:.2f}%", flush=True)
    write_fold_csv(os.path.join(opt.out_root, "statistics", f"{opt.data_type}_results_overall.csv"),
                   over_results, "fold")
    tr, te = np.array(over_results["train_accuracy"]), np.array(over_results["test_accuracy"])
    print("Overall Training Accuracy: %.2f%% (std: %.2f) Testing Accuracy: %.2f%% (std: %.2f)"
          % (tr.mean(), tr.std(), te.mean(), te.std()), flush=True)
    return over_results


if __name__ == "__main__":
    main(sys.argv[1:])
