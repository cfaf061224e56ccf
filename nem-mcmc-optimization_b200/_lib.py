"""ctypes binding of libnemb200.so (the C ABI declared in include/nemb200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
# NEMB200_LIB_VARIANT=<tag> loads csrc/libnemb200_<tag>.so (diagnostic builds, e.g. "trace" compiled with
# NEMB200_EXTRA_NVCC_FLAGS=-DNEMB200_TRACE); the product library is libnemb200.so
_VARIANT = os.environ.get("NEMB200_LIB_VARIANT", "")
_SO = os.path.join(_CSRC, f"libnemb200_{_VARIANT}.so" if _VARIANT else "libnemb200.so")
_SOURCES = ["dataset.cu", "graph.cu", "score_tc.cu", "mcmc.cu", "mcmc_fused.cu", "cell.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

_lib = None


class NemB200Error(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into csrc/libnemb200.so (in-tree)."""
    srcs = [os.path.join(_CSRC, s) for s in _SOURCES]
    deps = srcs + [os.path.join(_CSRC, "common.cuh"),
                   os.path.join(_CSRC, "..", "..", "include", "nemb200.h")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return _SO
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("NEMB200_EXTRA_NVCC_FLAGS", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", _SO] + srcs
    subprocess.check_call(cmd)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise NemB200Error(
            f"{_SO} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback.")
    L = C.CDLL(_SO)
    vp, i32, i64, u64, u32, dbl, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_double, C.c_size_t
    L.nemb200_version.restype = i32
    L.nemb200_last_error.restype = C.c_char_p
    L.nemb200_dataset_bytes.restype = sz
    L.nemb200_dataset_bytes.argtypes = [i32, i32]
    L.nemb200_dataset_create.argtypes = [vp, i32, i32, dbl, dbl, vp, sz, vp, C.POINTER(vp)]
    L.nemb200_dataset_destroy.argtypes = [vp]
    L.nemb200_dataset_info.argtypes = [vp, C.POINTER(i64), i32]
    L.nemb200_closure_u32.argtypes = [vp, i32, i32, vp, vp]
    L.nemb200_transitive_reduction_u32.argtypes = [vp, i32, i32, vp, vp]
    L.nemb200_score_workspace_bytes.restype = sz
    L.nemb200_score_workspace_bytes.argtypes = [vp, i32]
    L.nemb200_score_workspace_bytes_impl.restype = sz
    L.nemb200_score_workspace_bytes_impl.argtypes = [vp, i32, i32]
    L.nemb200_score_batch.argtypes = [vp, vp, i32, vp, vp, sz, i32, vp]
    L.nemb200_closure_score_batch.argtypes = [vp, vp, i32, vp, vp, sz, i32, vp]
    L.nemb200_pair_fallback_count.argtypes = [vp, vp, i32]
    L.nemb200_order_to_graph.argtypes = [vp, i32, i32, i32, vp, i32, vp, vp]
    L.nemb200_mcmc_workspace_bytes.restype = sz
    L.nemb200_mcmc_workspace_bytes.argtypes = [vp, i32]
    L.nemb200_mcmc_workspace_bytes_rule.restype = sz
    L.nemb200_mcmc_workspace_bytes_rule.argtypes = [vp, i32, i32]
    L.nemb200_mcmc_run.argtypes = [vp, i32, i64, i64, i64, vp, dbl, u64, u32, i32, vp, i32, i32, i32,
                                   vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.nemb200_replica_exchange.argtypes = [i32, i32, i32, i64, u64, i32, vp, vp, vp, vp, vp, vp]
    L.nemb200_replica_exchange_best.argtypes = [i32, i32, i32, i64, u64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.nemb200_edge_counts_reduce.argtypes = [vp, i32, i32, vp, vp]
    L.nemb200_cell_ratios.argtypes = [vp, vp, i32, vp, vp]
    L.nemb200_cell_ll_workspace_bytes.restype = sz
    L.nemb200_cell_ll_workspace_bytes.argtypes = [i32]
    L.nemb200_cell_ll.argtypes = [vp, i32, i32, vp, vp, vp, sz, vp]
    L.nemb200_synth_knockdown.argtypes = [vp, vp, i32, i32, dbl, dbl, u64, vp, vp, vp]
    L.nemb200_ancestor_x.argtypes = [vp, vp, i32, vp, vp, vp]
    L.nemb200_fit_edges.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.nemb200_selftest_partition.argtypes = [i32, i32, i32]
    _lib = L
    return L


EXPORTS = ["nemb200_version", "nemb200_last_error", "nemb200_dataset_bytes", "nemb200_dataset_create",
           "nemb200_dataset_destroy", "nemb200_dataset_info", "nemb200_closure_u32", "nemb200_transitive_reduction_u32",
           "nemb200_score_workspace_bytes", "nemb200_score_workspace_bytes_impl", "nemb200_score_batch", "nemb200_closure_score_batch", "nemb200_pair_fallback_count", "nemb200_order_to_graph",
           "nemb200_mcmc_workspace_bytes", "nemb200_mcmc_workspace_bytes_rule", "nemb200_mcmc_run", "nemb200_replica_exchange", "nemb200_replica_exchange_best",
           "nemb200_edge_counts_reduce", "nemb200_cell_ratios", "nemb200_cell_ll_workspace_bytes",
           "nemb200_cell_ll", "nemb200_synth_knockdown", "nemb200_ancestor_x", "nemb200_fit_edges", "nemb200_selftest_partition"]


def check(rc: int) -> None:
    if rc != 0:
        raise NemB200Error(f"nemb200 call failed ({rc}): {lib().nemb200_last_error().decode()}")
