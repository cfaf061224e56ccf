"""ctypes binding of oracle/nem_oracle.c (TEST INFRASTRUCTURE ONLY, see nem_oracle.py header).

Build with ``make -C oracle`` (``__graft_entry__.build()`` does it)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnem_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "nem_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_score.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def closure(adj):
    adj = np.ascontiguousarray(adj, dtype=np.uint8)
    out = np.empty_like(adj)
    lib().orc_closure(_p(adj, C.c_uint8), C.c_int(adj.shape[0]), _p(out, C.c_uint8))
    return out


def score(adj, D, A, B):
    adj = np.ascontiguousarray(adj, dtype=np.uint8)
    D = np.ascontiguousarray(D, dtype=np.uint8)
    S, E = D.shape
    return lib().orc_score(_p(adj, C.c_uint8), _p(D, C.c_uint8), C.c_int(S), C.c_int(E), C.c_double(A), C.c_double(B))


def score_batch(adj, D, A, B, nthreads=1):
    adj = np.ascontiguousarray(adj, dtype=np.uint8)
    D = np.ascontiguousarray(D, dtype=np.uint8)
    S, E = D.shape
    n = adj.shape[0]
    out = np.empty(n, dtype=np.float64)
    lib().orc_score_batch(_p(adj, C.c_uint8), C.c_int(n), _p(D, C.c_uint8), C.c_int(S), C.c_int(E),
                          C.c_double(A), C.c_double(B), _p(out, C.c_double), C.c_int(nthreads))
    return out


def mcmc_chain(D, A, B, order0, n_iter, gamma, swap_prob=0.95, seed=1234, chain=0, rule=0, cand=None,
               use_nem=True, burn_in=0, keep_scores=False):
    D = np.ascontiguousarray(D, dtype=np.uint8)
    S, E = D.shape
    order0 = np.ascontiguousarray(order0, dtype=np.int32)
    cand_a = np.ascontiguousarray(cand if cand is not None else np.ones((S, S)), dtype=np.uint8)
    best_score, curr_score = C.c_double(), C.c_double()
    n_acc = C.c_int64()
    best_order = np.empty(S, dtype=np.int32)
    curr_order = np.empty(S, dtype=np.int32)
    best_graph = np.empty((S, S), dtype=np.uint8)
    edge = np.empty((S, S), dtype=np.int64)
    alls = np.empty(n_iter + 1, dtype=np.float64) if keep_scores else None
    lib().orc_mcmc_chain(_p(D, C.c_uint8), C.c_int(S), C.c_int(E), C.c_double(A), C.c_double(B),
                         _p(order0, C.c_int32), C.c_int64(n_iter), C.c_double(gamma), C.c_double(swap_prob),
                         C.c_uint64(seed), C.c_uint32(chain), C.c_int(rule), _p(cand_a, C.c_uint8),
                         C.c_int(1 if use_nem else 0), C.c_int64(burn_in), C.byref(best_score),
                         _p(best_order, C.c_int32), _p(best_graph, C.c_uint8), C.byref(curr_score),
                         _p(curr_order, C.c_int32), C.byref(n_acc), _p(edge, C.c_int64),
                         _p(alls, C.c_double) if keep_scores else None)
    return dict(best_score=best_score.value, best_order=best_order, best_graph=best_graph,
                curr_score=curr_score.value, curr_order=curr_order, n_accept=n_acc.value,
                edge_counts=edge, all_score_list=alls)
