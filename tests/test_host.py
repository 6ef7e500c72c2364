"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the
header declares; argument validation returns the documented codes before any CUDA
call; host logic (synthetic batches, collate, sharding, module surface)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import dgcnn_b200 as dg
from dgcnn_b200 import _lib
from dgcnn_b200.synth import CONFIGS, collate, make_batch, make_graphs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dgcnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dgcnn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    lib = _lib.load_library()
    syms = header_symbols()
    assert len(syms) >= 11
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dgcnn_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.dgcnn_abi_version() == 2
    assert lib.dgcnn_status_string(0) == b"ok"
    assert b"workspace" in lib.dgcnn_status_string(-3)


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\w+", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_workspace_queries_are_pure_host_arithmetic():
    lib = _lib.load_library()
    assert lib.dgcnn_build_graph_workspace_bytes(0, 0) >= 0
    small = lib.dgcnn_build_graph_workspace_bytes(1000, 5000)
    big = lib.dgcnn_build_graph_workspace_bytes(256000, 5068800)
    assert 8 * 1000 + 8 * 5000 <= small < big
    assert lib.dgcnn_graph_conv_bwd_workspace_bytes(1000, 32, 32) >= 2 * 1000 * 32 * 4
    assert lib.dgcnn_sort_pool_workspace_bytes(1000, 10) >= 16 * 1000


def test_argument_validation_without_touching_cuda():
    lib = _lib.load_library()
    INVALID, UNSUPPORTED, WORKSPACE = -1, -2, -3
    # K1: bad channel counts / enums are rejected before any launch
    assert lib.dgcnn_graph_conv_fwd(None, 0, 0, None, None, None, None, None, None, 0, 0, 5, 0, 0, None) == INVALID
    assert lib.dgcnn_graph_conv_fwd(None, 200, 200, None, None, None, None, None, None, 32, 32, 5, 0, 0, None) == UNSUPPORTED
    assert lib.dgcnn_graph_conv_fwd(None, 8, 8, None, None, None, None, None, None, 32, 32, 5, 7, 0, None) == INVALID
    assert lib.dgcnn_graph_conv_fwd(None, 8, 8, None, None, None, None, None, None, 32, 32, 5, 0, 0, None) == INVALID
    assert lib.dgcnn_graph_conv_fwd(None, 8, 8, None, None, None, None, None, None, 32, 32, 0, 0, 0, None) == 0
    # K0: missing workspace
    assert lib.dgcnn_build_graph(1, 4, 1, 3, 1, 1, 1, None, None, 1, 1, None, None, 0, None, 0, None) == WORKSPACE
    assert lib.dgcnn_build_graph(None, -1, None, 3, 1, None, None, None, None, None, None, None, None, 0, None, 0, None) == INVALID
    # K2: k < 1
    assert lib.dgcnn_sort_pool_fwd(None, 97, 97, None, 10, 2, 0, 0, None, None, None, 0, None) == INVALID
    assert lib.dgcnn_sort_pool_fwd(None, 97, 97, None, 0, 0, 30, 0, None, None, None, 0, None) == 0


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="CUDA-only"):
        dg.ops.graph_conv_fwd(torch.zeros(2, 2), torch.zeros(3, dtype=torch.int32),
                              torch.zeros(1, dtype=torch.int32), torch.zeros(2), torch.zeros(2, 2),
                              None, 0, 0, torch.zeros(2, 2))
    with pytest.raises(NotImplementedError):
        torch.ops.dgcnn_b200.sort_pool_bwd(torch.zeros(1, 4), torch.zeros(1, 2, dtype=torch.int32), 3)
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import dgcnn_b200; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % ROOT)
    assert subprocess.run([sys.executable, "-c", code]).returncode == 0, \
        "importing the product must not import the oracle"


def test_product_sources_never_import_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|import_module\(.oracle|oracle/|oracle\.dgcnn", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dgcnn_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f"{f} reaches into oracle/"


@pytest.mark.parametrize("f,c,count", [(8, 2, 52035), (19, 2, 52387), (38, 2, 52995), (5, 2, 51939),
                                       (90, 2, 54659), (1, 3, 51940), (1, 2, 51811)])
def test_model_surface_matches_reference(f, c, count):
    m = dg.Model(f, c)
    assert sum(p.numel() for p in m.parameters()) == count           # README.md:96-104
    names = [n for n, _ in m.named_children()]
    assert names == ["conv1", "conv2", "conv3", "conv4", "sort_pool", "conv5", "conv6", "pool",
                     "classifier_1", "drop_out", "classifier_2", "relu"]          # model.py:13-24
    sd = m.state_dict()
    assert sd["conv1.lin.weight"].shape == (32, f) and sd["conv4.bias"].shape == (1,)
    assert torch.count_nonzero(sd["conv2.bias"]) == 0                 # PyG zeros() bias init
    a = (6.0 / (32 + 32)) ** 0.5
    assert sd["conv2.lin.weight"].abs().max() <= a                    # PyG glorot() bound
    assert dg.GraphConvolution is dg.GCNConv and dg.SortPool is dg.SortAggregation


def test_model_k_parameter():
    for k, width in [(30, 352), (60, 832), (291, 4512), (130, 1952), (512, 8064)]:
        assert dg.Model(5, 2, k=k).classifier_1.in_features == width
    with pytest.raises(ValueError):
        dg.Model(5, 2, k=8)


def test_remove_self_loops():
    ei = torch.tensor([[0, 1, 1, 2, 2], [0, 2, 1, 1, 0]])
    out, attr = dg.remove_self_loops(ei)
    assert attr is None and out.tolist() == [[1, 2, 2], [2, 1, 0]]


@pytest.mark.parametrize("name", ["mutag", "proteins", "dd", "collab"])
def test_synthetic_batches_have_the_reference_layout(name):
    cfg = CONFIGS[name]
    b = make_batch(name)
    assert b.num_graphs == cfg.batch_size and b.x.shape[1] == cfg.num_features
    assert b.x.dtype == torch.float32 and b.edge_index.dtype == torch.int64
    ei = b.edge_index
    assert (ei[0] != ei[1]).all()
    key = ei[0] * b.num_nodes + ei[1]
    assert (key[1:] > key[:-1]).all()                      # sorted by (src,dst), no duplicates
    assert torch.equal(torch.sort(ei[1] * b.num_nodes + ei[0]).values, key)   # symmetric
    assert (b.batch[1:] >= b.batch[:-1]).all() and int(b.batch[-1]) == b.num_graphs - 1
    assert torch.equal(b.ptr, torch.cat([torch.zeros(1, dtype=torch.long),
                                         torch.bincount(b.batch, minlength=b.num_graphs).cumsum(0)]))
    assert (b.batch[ei[0]] == b.batch[ei[1]]).all()        # block diagonal
    # last column = in-degree / max in-degree per graph (utils.py:18-33)
    deg = torch.bincount(ei[1], minlength=b.num_nodes).float()
    gmax = torch.zeros(b.num_graphs).scatter_reduce_(0, b.batch, deg, "amax")
    np.testing.assert_allclose(b.x[:, -1].numpy(), (deg / gmax[b.batch]).numpy(), rtol=1e-6)
    b2 = make_batch(name)
    assert torch.equal(b.x, b2.x) and torch.equal(b.edge_index, b2.edge_index)   # deterministic
    if name == "dd":
        assert int((b.ptr[1:] - b.ptr[:-1]).max()) == 5748


def test_collate_offsets():
    gs = make_graphs(CONFIGS["mutag"], 4, seed=0)
    b = collate(gs)
    off = 0
    for i, g in enumerate(gs):
        n = g["x"].shape[0]
        sel = (b.batch[b.edge_index[0]] == i)
        np.testing.assert_array_equal(b.edge_index[:, sel].numpy(), g["edge_index"] + off)
        off += n
    assert collate([]).num_graphs == 0


def test_shard_bounds_balance_and_cover():
    costs = [1, 1, 1, 100, 1, 1, 50, 50, 1, 1]
    for w in (1, 2, 3, 4, 8):
        bounds = dg.shard_bounds(costs, w)
        assert len(bounds) == w and bounds[0][0] == 0 and bounds[-1][1] == len(costs)
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(w - 1))
    two = dg.shard_bounds(costs, 2)
    loads = [sum(costs[lo:hi]) for lo, hi in two]
    assert max(loads) <= 0.6 * sum(costs)
    eq = dg.shard_bounds([1.0] * 512, 8)
    assert [hi - lo for lo, hi in eq] == [64] * 8


def header_prototypes():
    """name -> number of parameters, parsed from the header's declarations."""
    text = open(os.path.join(ROOT, "include", "dgcnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for name, args in re.findall(r"\b(dgcnn_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        args = args.strip()
        protos[name] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_ctypes_table_has_the_headers_arity():
    protos = header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == protos[name], f"{name}: header has {protos[name]} parameters, ctypes {len(argtypes)}"


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """include/dgcnn_b200.h compiles as C (gcc, no CUDA, no C++), and the two structs that cross
    the boundary have the layout the ctypes mirrors assume."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc unavailable")
    fields = {"dgcnn_dataset": [f for f, _ in _lib.DgcnnDataset._fields_],
              "dgcnn_batch_graph": [f for f, _ in _lib.DgcnnBatchGraph._fields_]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dgcnn_b200.h"', "int main(void) {"]
    for struct, names in fields.items():
        lines.append(f'  printf("{struct} %zu\\n", sizeof({struct}));')
        for f in names:
            lines.append(f'  printf("{struct}.{f} %zu\\n", offsetof({struct}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True)
               .stdout.strip().splitlines())
    for struct, cls in (("dgcnn_dataset", _lib.DgcnnDataset), ("dgcnn_batch_graph", _lib.DgcnnBatchGraph)):
        assert int(out[struct]) == ctypes.sizeof(cls)
        for f, _ in cls._fields_:
            assert int(out[f"{struct}.{f}"]) == getattr(cls, f).offset, f"{struct}.{f}"


def test_resident_entry_points_validate_before_touching_cuda():
    lib = _lib.load_library()
    INVALID, WORKSPACE = -1, -3
    ds, out = _lib.DgcnnDataset(), _lib.DgcnnBatchGraph()
    assert lib.dgcnn_collate(None, None, 1, 1, 0, None, None, None, 0, None) == INVALID
    assert lib.dgcnn_collate(ctypes.byref(ds), 8, 1, 1, 0, ctypes.byref(out), None, None, 0, None) == INVALID
    assert lib.dgcnn_dataset_prepare(None, None, None, None) == INVALID
    assert lib.dgcnn_dataset_prepare(ctypes.byref(ds), 16, None, None) == INVALID      # empty data set
    assert lib.dgcnn_dataset_prepare(ctypes.byref(ds), 12, None, None) == INVALID      # gext not 16-byte aligned
    # a structurally complete data set / batch with a missing workspace: the size check comes last
    for f in ("x", "y", "gptr", "rowptr", "col", "dis", "gext"):
        setattr(ds, f, 256)
    ds.num_graphs, ds.num_nodes, ds.num_edges, ds.num_features, ds.symmetric, ds.ldx = 4, 40, 100, 3, 1, 3
    for f in ("rowptr", "col", "dis", "gptr"):
        setattr(out, f, 256)
    assert lib.dgcnn_collate(ctypes.byref(ds), 256, 2, 20, 50, ctypes.byref(out), None, None, 0, None) == WORKSPACE
    out.bitmap = 256                                                                   # maps come as a set
    assert lib.dgcnn_collate(ctypes.byref(ds), 256, 2, 20, 50, ctypes.byref(out), None, 256, 1 << 20, None) == INVALID
    # workspace queries: pure host arithmetic, monotone, resident >= the batch buffers it adds
    assert lib.dgcnn_collate_workspace_bytes(512) >= 6 * 512 * 4
    assert lib.dgcnn_collate_workspace_bytes(4096) > lib.dgcnn_collate_workspace_bytes(512)
    a = lib.dgcnn_train_step_workspace_bytes(38898, 2359942, 512, 1, 130, 3, 492)
    b = lib.dgcnn_train_step_resident_workspace_bytes(38898, 2359942, 512, 1, 130, 3, 492)
    assert a > 0 and b > 0 and lib.dgcnn_train_step_resident_workspace_bytes(0, 0, 0, 1, 130, 3, 1) == 0
    assert lib.dgcnn_train_step_resident(None, None, 1, 0, 1, 30, 2, 1, 0, None, None, None, None, None,
                                         1e-3, 0.9, 0.999, 1e-8, 1, 0, 0, None, None, 1, 0, None, None, None,
                                         None, 0, None) == INVALID


def test_bench_reads_the_kernel_traffic_from_the_committed_profile(tmp_path, monkeypatch):
    """bench.py's roofline.traffic comes from profiles/r02_stack_fwd_mma.md (an ncu --set full
    summary), not from a constant: the parser on the committed file and on a hand-made one."""
    import bench
    got = bench.ks_traffic_from_profile()
    assert got is not None and 100_000 < got < 200_000_000
    fake = tmp_path / "ks.md"
    fake.write_text("| metric | value | unit |\n|---|---:|---|\n| gpu__time_duration.sum | 32.4 | us |\n"
                    "| dram__bytes_read.sum | 1.5 | Mbyte |\n| dram__bytes_write.sum | 250.0 | Kbyte |\n"
                    "| dram__bytes_read.sum | 9.0 | Mbyte |\n")
    monkeypatch.setattr(bench, "KS_PROFILE", str(fake))
    assert bench.ks_traffic_from_profile() == 1_750_000
    monkeypatch.setattr(bench, "KS_PROFILE", str(tmp_path / "missing.md"))
    assert bench.ks_traffic_from_profile() is None
