"""Drives the UNMODIFIED reference modules staged in oracle/_ref/ (see ref_stage.py) along the score path
exactly as nem.py:132-142 does (TEST INFRASTRUCTURE ONLY: bench.py's CPU legs and the tests import this).

`RefScorer(D, A, B)` builds the reference's resident tables without running NEM.__init__ (whose two
compute_real_score EM loops take hours at S=50, E=10k - SURVEY Q10): the instance is created with
object.__new__ and gets only the attributes the table builders read (num_s, num_e, A, B,
observed_knockdown_mat, nem.py:10-19); `score_table_list` / `U` then come from the reference's own
get_score_tables / get_node_lr_table (nem.py:49-64).  `score(adj)` = utils.ancestor on float input
(utils.py:37-54; float because the int64 path counts wrap for dense S >= 66) + the parents-list
construction of nem.py:132-140 + get_reduced_score_tables + compute_ll_ratios + calculate_ll (nem.py:66-86).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

from . import ref_stage


def _load(name):
    path = os.path.join(ref_stage.REF_DIR, name + ".py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod          # nem.py does `import utils`
    spec.loader.exec_module(mod)
    return mod


_mods = None


def modules():
    """(utils, nem) of the staged reference; raises if oracle/_ref is not staged."""
    global _mods
    if _mods is None:
        if not ref_stage.staged():
            raise RuntimeError("oracle/_ref is not staged (run python -m oracle.ref_stage where /root/reference is mounted)")
        saved = {k: sys.modules.get(k) for k in ("utils", "nem")}
        try:
            u = _load("utils")
            n = _load("nem")
        finally:
            for k, v in saved.items():   # do not leave the reference's flat module names behind
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
        _mods = (u, n)
    return _mods


class RefScorer:
    def __init__(self, D, A, B):
        self.utils, nem_mod = modules()
        D = np.asarray(D, dtype=np.float64)
        obj = object.__new__(nem_mod.NEM)
        obj.num_s, obj.num_e = D.shape
        obj.A, obj.B = A, B
        obj.observed_knockdown_mat = D
        self.nem = obj
        self.score_table_list = obj.get_score_tables(D)          # S tables of S x E float64
        self.U = obj.get_node_lr_table(self.score_table_list)

    def score(self, adj, close=True):
        adj = np.asarray(adj)
        s = self.nem.num_s
        phi = self.utils.ancestor(adj.astype(np.float64)) if close else adj
        parents_list = [[] for _ in range(s)]                    # nem.py:132-140
        weights = [[] for _ in range(s)]
        n_parents = np.zeros(s, dtype=int)
        for i in range(s):
            for j in range(s):
                if phi[i][j] == 1 and i != j:                    # (nem.py:90-91 zeroes the diagonal first)
                    parents_list[j].append(i)
                    n_parents[j] += 1
                    weights[j].append(1.0)
        reduced = self.nem.get_reduced_score_tables(self.score_table_list, parents_list)
        _, ll = self.nem.calculate_ll(self.nem.compute_ll_ratios(weights, reduced, n_parents, self.U))
        return float(ll)
