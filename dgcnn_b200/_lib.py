"""Build and load ``libdgcnn_b200.so`` -- the C-ABI shared library declared in
``include/dgcnn_b200.h`` -- and bind it with ctypes.

There is NO CPU fallback: if the library cannot be loaded every operator raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC_DIR = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "lib" / "libdgcnn_b200.so"
INCLUDE_DIR = PKG_DIR.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no PTX for other archs
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]

class DgcnnDataset(ctypes.Structure):
    """``dgcnn_dataset`` of include/dgcnn_b200.h: a HOST struct of device pointers."""
    _fields_ = [("num_graphs", c_int64), ("num_nodes", c_int64), ("num_edges", c_int64),
                ("num_features", c_int32), ("symmetric", c_int32),
                ("x", c_void_p), ("ldx", c_int64), ("y", c_void_p),
                ("gptr", c_void_p), ("rowptr", c_void_p), ("col", c_void_p),
                ("rowptr_t", c_void_p), ("col_t", c_void_p), ("dis", c_void_p),
                ("bitmap", c_void_p), ("bitmap_t", c_void_p), ("bmoff", c_void_p), ("gflags", c_void_p),
                ("gflags_t", c_void_p), ("fragmap", c_void_p), ("fgoff", c_void_p), ("gext", c_void_p)]


class DgcnnBatchGraph(ctypes.Structure):
    """``dgcnn_batch_graph`` of include/dgcnn_b200.h: where dgcnn_collate writes a batch."""
    _fields_ = [("x", c_void_p), ("ldx", c_int64), ("batch32", c_void_p), ("y", c_void_p),
                ("rowptr", c_void_p), ("col", c_void_p), ("rowptr_t", c_void_p), ("col_t", c_void_p),
                ("dis", c_void_p), ("gptr", c_void_p), ("gorder", c_void_p),
                ("bitmap", c_void_p), ("bitmap_t", c_void_p), ("bmoff", c_void_p), ("gflags", c_void_p),
                ("gflags_t", c_void_p), ("fragmap", c_void_p), ("fgoff", c_void_p), ("gdesc", c_void_p)]


_DATASET_P = ctypes.POINTER(DgcnnDataset)
_BATCH_P = ctypes.POINTER(DgcnnBatchGraph)
_STEP_TAIL = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_int64,
              c_int32, c_uint64, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p,
              c_void_p, c_void_p, c_size_t, c_void_p]

# name -> (restype, argtypes); mirrors include/dgcnn_b200.h one to one
SIGNATURES = {
    "dgcnn_collate_workspace_bytes": (c_size_t, [c_int64]),
    "dgcnn_dataset_prepare": (c_int32, [_DATASET_P, c_void_p, c_void_p, c_void_p]),
    "dgcnn_collate": (c_int32, [_DATASET_P, c_void_p, c_int64, c_int64, c_int64, _BATCH_P, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    "dgcnn_train_step_resident_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int32, c_int32,
                                                             c_int32, c_int64]),
    "dgcnn_train_step_resident": (c_int32, [_DATASET_P, c_void_p, c_int64, c_int64, c_int64, c_int32,
                                            c_int32, c_int64, c_int32] + _STEP_TAIL),
    "dgcnn_train_step_resident_graphed": (c_int32, [_DATASET_P, c_void_p, c_int64, c_int64, c_int64, c_int32,
                                            c_int32, c_int64, c_int32] + _STEP_TAIL),
    "dgcnn_abi_version": (c_int32, []),
    "dgcnn_status_string": (c_char_p, [c_int32]),
    "dgcnn_build_graph_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "dgcnn_build_graph": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                    c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                    c_void_p, c_size_t, c_void_p]),
    "dgcnn_build_graph_i32": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                        c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                        c_void_p, c_size_t, c_void_p]),
    "dgcnn_graph_bitmap_words": (c_int64, [c_int64, c_int64, c_int64]),
    "dgcnn_graph_fragmap_words": (c_int64, [c_int64, c_int64, c_int64]),
    "dgcnn_build_bitmaps": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int64, c_int64, c_int64,
                                      c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int32, c_void_p]),
    "dgcnn_graph_ptr": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "dgcnn_graph_conv_fwd": (c_int32, [c_void_p, c_int64, c_int32,
                                       c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p,
                                       c_void_p, c_int64, c_int32,
                                       c_int64, c_int32, c_int32, c_void_p]),
    "dgcnn_graph_conv_fwd_graphs": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_int64, c_int64,
                                              c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                              c_int64, c_int32, c_int32, c_void_p]),
    "dgcnn_project_rows": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    "dgcnn_graph_conv_bwd_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "dgcnn_graph_conv_bwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64,
                                       c_void_p, c_int64, c_int32,
                                       c_void_p, c_void_p, c_void_p,
                                       c_void_p,
                                       c_void_p, c_int64, c_int32,
                                       c_void_p, c_void_p, c_int32,
                                       c_int64, c_int32, c_int32,
                                       c_void_p, c_size_t, c_void_p]),
    "dgcnn_sort_pool_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "dgcnn_sort_pool_fwd": (c_int32, [c_void_p, c_int64, c_int32,
                                      c_void_p, c_int64, c_int64, c_int32,
                                      c_int64, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p]),
    "dgcnn_stack_fwd_supported": (c_int32, [c_int32, c_int64]),
    "dgcnn_stack_fwd_workspace_bytes": (c_size_t, []),
    "dgcnn_stack_fwd_set_trace": (None, [c_void_p]),
    "dgcnn_stack_fwd_configure": (None, [c_int32, c_int32]),
    "dgcnn_stack_fwd": (c_int32, [c_void_p, c_int64, c_int32,
                                  c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p,
                                  c_int64, c_int64, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_void_p, c_void_p, c_int32,
                                  c_int32, c_int32, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "dgcnn_stack_fwd_conv5_supported": (c_int32, [c_int32, c_int64]),
    "dgcnn_stack_fwd_conv5": (c_int32, [c_void_p, c_int64, c_int32,
                                        c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p,
                                        c_int64, c_int64, c_int64,
                                        c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p,
                                        c_void_p, c_int64, c_void_p, c_void_p, c_int32,
                                        c_void_p, c_void_p, c_int32, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "dgcnn_stack_bwd_conv5_supported": (c_int32, [c_int32, c_int64]),
    "dgcnn_stack_conv5_num_params": (c_int64, [c_int32]),
    "dgcnn_stack_bwd_conv5": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32,
                                        c_void_p, c_int64, c_void_p, c_int64,
                                        c_int32, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p,
                                        c_int64, c_int64, c_int64,
                                        c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int32, c_void_p, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "dgcnn_tail_bwd_h1": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int32,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_void_p] * 7 +
                          [c_int32, c_void_p, c_size_t, c_void_p]),
    "dgcnn_stack_bwd_supported": (c_int32, [c_int32, c_int64]),
    "dgcnn_stack_bwd_set_trace": (None, [c_void_p]),
    "dgcnn_stack_num_params": (c_int64, [c_int32]),
    "dgcnn_stack_bwd_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int64]),
    "dgcnn_stack_bwd": (c_int32, [c_void_p, c_void_p, c_int32,
                                  c_void_p, c_int64, c_void_p, c_int64,
                                  c_int32, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int64, c_int64,
                                  c_int64, c_void_p, c_void_p, c_void_p,
                                  c_int32, c_int32, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "dgcnn_tail_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "dgcnn_tail_fwd": (c_int32, [c_void_p, c_int64, c_int32] + [c_void_p] * 8 +
                       [c_int32, c_int32, c_uint64, c_void_p] + [c_void_p] * 6 +
                       [c_void_p, c_size_t, c_void_p]),
    "dgcnn_tail_bwd": (c_int32, [c_void_p, c_void_p, c_int64, c_int32] + [c_void_p] * 4 + [c_int32] +
                       [c_void_p] * 6 + [c_void_p] * 9 + [c_int32, c_void_p, c_size_t, c_void_p]),
    "dgcnn_tail_fwd_loss": (c_int32, [c_void_p, c_int64, c_int32] + [c_void_p] * 8 +
                            [c_int32, c_void_p, c_int32, c_uint64, c_void_p] + [c_void_p] * 6 +
                            [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "dgcnn_tail_bwd_after_loss": (c_int32, [c_void_p, c_int64, c_int32] + [c_void_p] * 4 + [c_int32] +
                                  [c_void_p] * 6 + [c_void_p] * 10 + [c_void_p, c_void_p, c_int32,
                                                                      c_void_p, c_size_t, c_void_p]),
    "dgcnn_tail_bwd_join": (c_int32, [c_void_p]),
    "dgcnn_adam_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_float, c_float, c_float, c_float, c_float, c_void_p]),
    "dgcnn_allreduce_adam_exchange_bytes": (c_size_t, [c_int64, c_int32]),
    "dgcnn_exchange_create": (c_int32, [c_int64, c_int32, c_void_p, c_void_p]),
    "dgcnn_exchange_open": (c_int32, [c_void_p, c_void_p]),
    "dgcnn_exchange_close": (c_int32, [c_void_p]),
    "dgcnn_exchange_destroy": (c_int32, [c_void_p]),
    "dgcnn_allreduce_set_trace": (None, [c_void_p]),
    "dgcnn_allreduce_adam": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                       c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_float,
                                       c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "dgcnn_train_step_configure": (None, [c_int32]),
    "dgcnn_train_step_configure_maps": (None, [c_int32]),
    "dgcnn_train_step_graph_counts": (None, [c_void_p]),
    "dgcnn_train_step_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int32, c_int32, c_int32,
                                                    c_int64]),
    "dgcnn_train_step_num_params": (c_int64, [c_int32, c_int32, c_int32]),
    "dgcnn_train_step": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int64,
                                   c_int64, c_int32, c_int32, c_int32, c_int64, c_int32, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_int64,
                                   c_int32, c_uint64, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "dgcnn_nll_sum": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p, c_void_p,
                                c_void_p]),
    "dgcnn_sort_pool_bwd": (c_int32, [c_void_p, c_void_p, c_int64,
                                      c_int32, c_int32, c_void_p, c_int64, c_int64,
                                      c_void_p]),
}

_lock = threading.Lock()
_lib = None


def sources():
    return sorted(CSRC_DIR.glob("*.cu"))


def _source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive a copy of
    the tree to another machine; contents do)."""
    import hashlib
    h = hashlib.sha1()
    deps = sorted(CSRC_DIR.glob("*.cu")) + sorted(CSRC_DIR.glob("*.cuh")) + sorted(INCLUDE_DIR.glob("*.h"))
    for p in deps:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + os.environ.get("DGCNN_NVCC_EXTRA", "").split()).encode())
    return h.hexdigest()


def _stale() -> bool:
    stamp = LIB_PATH.with_suffix(".srchash")
    if not LIB_PATH.exists() or not stamp.exists():
        return True
    return stamp.read_text().strip() != _source_hash()


def _compile_one(nvcc: str, src: Path, obj: Path, flags) -> None:
    cmd = [nvcc, *flags, "-I", str(INCLUDE_DIR), "-c", str(src), "-o", str(obj)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        obj.unlink(missing_ok=True)
        raise RuntimeError(f"dgcnn_b200: nvcc failed\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    if proc.stderr.strip() and "-Xptxas=-v" in flags:
        print(proc.stderr)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """nvcc-compile every kernel for sm_100a into one in-tree shared library: one object per
    source (compiled in parallel, reused while neither the source, a header nor the flags
    changed), then one link."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("dgcnn_b200: nvcc not found; cannot build libdgcnn_b200.so")
    LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
    obj_dir = PKG_DIR / "build"
    obj_dir.mkdir(parents=True, exist_ok=True)
    extra = os.environ.get("DGCNN_NVCC_EXTRA", "").split()      # e.g. -DDGCNN_FWD_THREADS=768 (tuning runs)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas=-v"] if verbose else [])
    stamp = obj_dir / "flags.txt"
    flags_changed = (not stamp.exists()) or stamp.read_text() != " ".join(cflags)
    headers = list(CSRC_DIR.glob("*.cuh")) + list(INCLUDE_DIR.glob("*.h"))
    hdr_time = max(p.stat().st_mtime for p in headers)
    jobs = []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        if (force and verbose) or flags_changed or not obj.exists() or \
                obj.stat().st_mtime < max(src.stat().st_mtime, hdr_time):
            jobs.append((src, obj))
    if jobs:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            list(pool.map(lambda j: _compile_one(nvcc, j[0], j[1], cflags), jobs))
    stamp.write_text(" ".join(cflags))
    tmp = LIB_PATH.with_suffix(f".tmp{os.getpid()}.so")
    objs = [str(obj_dir / (src.stem + ".o")) for src in sources()]
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(tmp), *objs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"dgcnn_b200: link failed\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    os.replace(tmp, LIB_PATH)
    LIB_PATH.with_suffix(".srchash").write_text(_source_hash())
    return LIB_PATH


def load_library() -> ctypes.CDLL:
    """Load (building first if the sources are newer) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _stale():
            # one builder at a time across PROCESSES too (torchrun starts N ranks at once)
            import fcntl
            LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
            with open(LIB_PATH.with_suffix(".lock"), "w") as lock_fh:
                fcntl.flock(lock_fh, fcntl.LOCK_EX)
                try:
                    build_library()               # (a no-op when another rank just built it)
                except RuntimeError:
                    if not LIB_PATH.exists():
                        raise
                finally:
                    fcntl.flock(lock_fh, fcntl.LOCK_UN)
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError = header/library mismatch: loud
            fn.restype, fn.argtypes = restype, argtypes
        if lib.dgcnn_abi_version() != 2:
            raise RuntimeError("dgcnn_b200: ABI version mismatch between _lib.py and the library")
        _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load_library().dgcnn_status_string(status).decode()
        raise RuntimeError(f"dgcnn_b200: {what} failed: {msg} ({status})")
