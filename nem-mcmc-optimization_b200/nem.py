"""Drop-in for the reference's ``nem.py``: class NEM with the same constructor, attributes and
method names.  The score of binary graphs runs on the B200 (Dataset.score_graphs); the
constructor's EM diagnostics (compute_real_score loop, nem.py:97-130) are out of scope
 - ``real_order_ll`` / ``obs_order_ll`` are None."""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _lib, engine, utils


class _ReducedTables(list):
    """list of per-node row stacks that remembers which parents the rows belong to, so that
    compute_ll_ratios can run as one dense-weight CUDA launch."""
    parents_list = None
    source = None  # id() of the knock-down matrix the tables were built from


class NEM:
    def __init__(self, adj_matrix, end_nodes, errors, num_s, num_e, seed=42, device=None):
        self.num_s = num_s
        self.num_e = num_e
        self.adj_matrix = adj_matrix
        alpha, beta = errors[0], errors[1]
        self.real_knockdown_mat = utils.create_real_knockdown_mat(adj_matrix, end_nodes)
        random.seed(seed)  # nem.py:16 (the observed matrix nevertheless always uses seed 42)
        self.A = np.log(alpha / (1.0 - beta))
        self.B = np.log(beta / (1.0 - alpha))
        self.observed_knockdown_mat = utils.create_observed_knockdown_mat(self.real_knockdown_mat, alpha, beta)
        self.device = engine._require_cuda(device)
        self.dataset = engine.Dataset(self.observed_knockdown_mat, self.A, self.B, self.device)
        self._score_dataset = None
        self._real_dataset = None
        self.U = self._lr_table(self.observed_knockdown_mat)
        self.real_order_ll, self.real_ll, self.real_parent_order = self.compute_real_score(self.real_knockdown_mat, adj_matrix)
        self.obs_order_ll, self.obs_ll, self.obs_parent_order = self.compute_real_score(self.observed_knockdown_mat, adj_matrix)

    @property
    def score_dataset(self):
        """The same data with the E-genes stored sorted by knock-down pattern (the score is a sum over E-genes, so
        their order is free): conflict-free table look-ups, about 4 % faster batched scoring / MCMC.  `dataset`
        keeps the caller's order for the per-E-gene outputs (cell ratios, posterior attachment weights)."""
        if self._score_dataset is None:
            self._score_dataset = engine.Dataset(self.observed_knockdown_mat, self.A, self.B, self.device, sort_egenes=True)
        return self._score_dataset

    # ---- tables (host-side, one-off; the GPU path only needs D, A, B) -------------------------------
    def compute_scores(self, node, knockdown_mat):
        """nem.py:25-34: base row of table `node` = T[node] + N."""
        kd = np.asarray(knockdown_mat)
        ones = (kd == 1)
        return np.where(ones[node], 0.0, self.B) + self.A * (ones.sum(axis=0) - ones[node])

    def build_score_table(self, node, knockdown_mat):
        """nem.py:36-47."""
        kd = np.asarray(knockdown_mat)
        table = np.where(kd == 0, self.B, -self.A).astype(np.float64)
        table[node, :] = self.compute_scores(node, kd)
        return table

    def get_score_tables(self, knockdown_mat):
        """nem.py:49-54 (S tables of S x E float64 - only for API compatibility)."""
        tabs = _ReducedTables(self.build_score_table(node, knockdown_mat) for node in range(self.num_s))
        tabs.source = np.asarray(knockdown_mat)
        return tabs

    def _lr_table(self, knockdown_mat):
        """get_node_lr_table(get_score_tables(kd)) without materialising the S x S x E tables."""
        kd = np.asarray(knockdown_mat)
        base = np.where(kd == 1, 0.0, self.B) + self.A * ((kd == 1).sum(axis=0)[None, :] - (kd == 1))
        return np.vstack([base, np.where(self.observed_knockdown_mat == 0, 0, self.A).sum(axis=0)])

    def get_node_lr_table(self, all_score_tables):
        """nem.py:56-64: rows = base rows, last row = N of the *observed* matrix."""
        rows = [all_score_tables[i][i] for i in range(self.num_s)]
        rows.append(np.where(self.observed_knockdown_mat == 0, 0, self.A).sum(axis=0))
        return np.vstack(rows)

    def get_reduced_score_tables(self, score_table_list, parents_list):
        """nem.py:66-70."""
        out = _ReducedTables(np.array([score_table_list[i][j] for j in parents_list[i]]) for i in range(self.num_s))
        out.parents_list = [list(p) for p in parents_list]
        out.source = getattr(score_table_list, "source", None)
        return out

    # ---- score ------------------------------------------------------------------------------------------
    def _dataset_for(self, kd):
        if kd is None or np.array_equal(kd, self.observed_knockdown_mat):
            return self.dataset
        if self._real_dataset is None or not np.array_equal(self._real_dataset[0], kd):
            self._real_dataset = (np.array(kd), engine.Dataset(kd, self.A, self.B, self.device))
        return self._real_dataset[1]

    def compute_ll_ratios(self, parent_weights, reduced_score_tables, real_n_parents, U):
        """nem.py:72-80.  With tables from get_reduced_score_tables this is one launch of the dense-weight
        CUDA kernel; rows 0..S-1 follow the tables' own data, row S is taken from U (nem.py:62 quirk)."""
        import ctypes as C
        parents = getattr(reduced_score_tables, "parents_list", None)
        if parents is None:
            return utils._rows_ll_ratios(U, real_n_parents, parent_weights, reduced_score_tables, self.device).cpu().numpy()
        S = self.num_s
        wmat = np.zeros((S, S), dtype=np.float64)
        for i in range(S):
            for k in range(int(real_n_parents[i])):
                wmat[i, parents[i][k]] = parent_weights[i][k]
        ds = self._dataset_for(getattr(reduced_score_tables, "source", None))
        w = torch.from_numpy(wmat).to(self.device)
        cell = torch.empty((S + 1, self.num_e), dtype=torch.float64, device=self.device)
        st = torch.cuda.current_stream(self.device)
        _lib.check(_lib.lib().nemb200_cell_ratios(ds._h, engine._ptr(w), 0, engine._ptr(cell), engine._stream_ptr(st)))
        out = cell.cpu().numpy()
        out[S] = np.asarray(U)[S]
        return out

    def calculate_ll(self, cell_ratios):
        """nem.py:82-86 -> (order_weights (S+1,E), ll)."""
        return utils.cell_ll(cell_ratios, want_weights=True, device=self.device)

    def score_graphs(self, graphs, close=False, impl="auto"):
        """Batched score of binary graphs on the observed data (new; the B200 hot path)."""
        return self.dataset.score_graphs(graphs, impl=impl, close=close)

    def compute_real_score(self, real_knockdown_mat, adj_mat):
        """nem.py:88-144 without the EM loop (:97-130, out of scope): returns
        (None, ll of `adj_mat` on the given data, parent_order).  Like the reference it zeroes the
        diagonal of the caller's adj_mat in place and takes the null row of U from the
        observed matrix (nem.py:62)."""
        for i in range(self.num_s):
            adj_mat[i][i] = 0
        adj = np.asarray(adj_mat)
        parent_order = np.argsort(np.sum(adj, axis=1))[::-1]
        real_U = self._lr_table(real_knockdown_mat)
        parents_list = [[i for i in range(self.num_s) if adj[i][j] == 1] for j in range(self.num_s)]
        weights = [[1.0] * len(p) for p in parents_list]
        n_parents = np.array([len(p) for p in parents_list], dtype=int)
        reduced = _ReducedTables()
        reduced.parents_list, reduced.source = parents_list, np.asarray(real_knockdown_mat)
        _, ll = self.calculate_ll(self.compute_ll_ratios(weights, reduced, n_parents, real_U))
        return None, ll, parent_order
