"""Drop-in for the reference's ``utils.py`` on the hot path: same names, argument meaning and
return types.  Graph/score arithmetic runs on the GPU through libnemb200; data synthesis, CSV
reading and small order utilities are host-side input preparation exactly as in the reference."""
from __future__ import annotations

import os
import random

import numpy as np
import torch

from . import _lib, engine


# ---- data synthesis (host side input preparation, utils.py:5-35) --------------------------------
def create_connection_mat(s_mat):
    """utils.py:5-13: conn[j][i] = 1 iff i == j or s_mat[i][j] == 1."""
    s = np.asarray(s_mat)
    n = s.shape[1]
    conn = np.eye(n)
    conn[np.nonzero(s.T == 1)] = 1
    return conn


def create_real_knockdown_mat(s_mat, e_arr):
    """utils.py:15-23: column k is the reachability column of the S-gene E-gene k hangs on."""
    conn = create_connection_mat(s_mat)
    return np.ascontiguousarray(conn[np.asarray(e_arr, dtype=np.int64)].T)


def create_observed_knockdown_mat(knockdown_mat, alpha, beta, seed=42):
    """utils.py:25-35: one Mersenne-Twister uniform per cell in row-major order; false positives with
    probability alpha, false negatives with probability beta (the module-global ``random`` is
    re-seeded, as the reference does)."""
    random.seed(seed)
    kd = np.asarray(knockdown_mat)
    u = np.fromiter((random.random() for _ in range(kd.size)), dtype=np.float64, count=kd.size).reshape(kd.shape)
    flipped = kd.copy()
    flipped[(kd == 0) & (u < alpha)] = 1
    flipped[(kd == 1) & (u < beta)] = 0
    return flipped


# ---- closure (GPU) ---------------------------------------------------------------------------------
def ancestor(incidence):
    """utils.py:37-54 - transitive closure of an S x S incidence matrix, computed by the bit-packed
    CUDA closure kernel.  Boolean semantics (the reference's int64 path count wraps for dense
    S >= 66; this does not)."""
    inc = np.asarray(incidence)
    out = engine.closure((inc > 0).astype(np.uint8)[None], S=inc.shape[0])[0]
    return out.astype(int)


def initial_order_guess(observed_knockdown_mat):
    """utils.py:56-64."""
    return np.argsort(-np.sum(np.asarray(observed_knockdown_mat), axis=1))


# ---- score helpers (GPU) -----------------------------------------------------------------------------
def _rows_ll_ratios(U, n_parents, parent_weights, reduced_score_tables, device):
    """cell[i] = U[i] + sum_j log(1 - w_ij + w_ij exp(row_ij)) for explicit table rows (torch CUDA ops;
    compatibility form for callers that pass raw rows - the batched hot path is Dataset.score_graphs)."""
    cell = torch.as_tensor(np.asarray(U, dtype=np.float64)).to(device).clone()
    for i in range(len(parent_weights)):
        k = int(n_parents[i])
        if k == 0:
            continue
        rows = torch.as_tensor(np.asarray(reduced_score_tables[i], dtype=np.float64)[:k]).to(device)
        w = torch.as_tensor(np.asarray(parent_weights[i], dtype=np.float64)[:k]).to(device)[:, None]
        cell[i] += torch.log(1.0 - w + w * torch.exp(rows)).sum(dim=0)
    return cell


def compute_ll_ratios(n_parents, U, parent_weights, reduced_score_tables):
    """utils.py:66-82 (Eq. 13): (S+1, E) cell log-ratios for real-valued parent weights."""
    dev = engine._require_cuda()
    return _rows_ll_ratios(U, n_parents, parent_weights, reduced_score_tables, dev).cpu().numpy()


def cell_ll(cell_ratios, want_weights=False, device=None):
    """Column logaddexp + sum on the GPU (nemb200_cell_ll)."""
    import ctypes as C
    dev = engine._require_cuda(device)
    cell = cell_ratios if isinstance(cell_ratios, torch.Tensor) else torch.as_tensor(
        np.ascontiguousarray(np.asarray(cell_ratios, dtype=np.float64)))
    cell = cell.to(dev, torch.float64).contiguous()
    rows, E = cell.shape
    L = _lib.lib()
    ws = torch.empty(max(L.nemb200_cell_ll_workspace_bytes(E), 8), dtype=torch.uint8, device=dev)
    ll = torch.empty(1, dtype=torch.float64, device=dev)
    ow = torch.empty_like(cell) if want_weights else None
    st = torch.cuda.current_stream(dev)
    _lib.check(L.nemb200_cell_ll(engine._ptr(cell), rows, E, engine._ptr(ow), engine._ptr(ll), engine._ptr(ws),
                                 ws.numel(), engine._stream_ptr(st)))
    return (ow.cpu().numpy() if want_weights else None), float(ll.item())


def compute_ll(cell_ratios):
    """utils.py:84-94 (Eq. 14): sum over E-genes of logaddexp over the S+1 rows."""
    return cell_ll(cell_ratios)[1]


# ---- input formats / small host utilities ---------------------------------------------------------------
def read_csv_to_adj(pathname):
    """utils.py:96-118.  Line 1 ``S,E``; edge lines ``i,j``; one line of E attachment indices;
    last line ``alpha,beta``.  (The reference detects the end-node line as the first line without
    exactly two fields, so E == 2 is ambiguous - SURVEY Q14; here the last two lines are always the
    end nodes and the errors.)"""
    with open(os.path.join(os.getcwd(), pathname), "r") as f:
        lines = [ln.strip() for ln in f if ln.strip()]
    num_s, num_e = (int(x) for x in lines[0].split(","))
    adj = np.zeros((num_s, num_s), dtype=int)
    for ln in lines[1:-2]:
        a, b = (int(x) for x in ln.split(","))
        adj[a, b] = 1
    end_nodes = np.array([int(x) for x in lines[-2].split(",")])
    errors = np.array([float(x) for x in lines[-1].split(",")])
    return adj, end_nodes, errors, num_s, num_e


def write_network_csv(adj, end_nodes, errors, output_file, output_file_red=None):
    """Writes a network in the reference's CSV format (DAGs/rnd_dag_gen.py:86-103, read back by
    utils.read_csv_to_adj :96-118): line 1 ``S,E``; one ``i,j`` line per edge i -> j; one line of E attachment
    indices; last line ``alpha,beta``; csv.writer's CRLF line ends like the shipped files.  `output_file` receives
    the transitive closure of `adj` (networkN.csv), `output_file_red` - if given - its transitive reduction
    (networkN_red.csv); both are computed on the GPU (closure / transitive_reduction of this module).  Edges are
    written in row-major order (the reference writes them in generation order, which its reader ignores)."""
    import csv
    adj = np.asarray(adj)
    S = adj.shape[0]
    closed = ancestor(adj)
    first = [S, len(end_nodes)]
    ends = [int(x) for x in end_nodes]
    probs = [float(x) for x in errors]

    def dump(path, mat):
        with open(path, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(first)
            w.writerows((int(i), int(j)) for i, j in zip(*np.nonzero(mat)))
            w.writerow(ends)
            w.writerow(probs)

    dump(output_file, closed)
    if output_file_red is not None:
        dump(output_file_red, transitive_reduction(closed))
    return closed


def transitive_reduction(adj_matrix):
    """utils.py:120-129 (result post-processing), on the GPU (`nemb200_transitive_reduction_u32`): the
    reference's in-place triple loop, reproduced exactly - the Hasse diagram for a closed DAG; nonzero entries count
    as edges and the result is 0/1 in the input's dtype."""
    a = np.asarray(adj_matrix)
    out = engine.transitive_reduction((a != 0).astype(np.uint8)[None], S=a.shape[0])[0]
    return out.astype(a.dtype if a.dtype != bool else np.uint8)


def hamming_distance(mat_a, mat_b):
    """utils.py:148-149 (sum |a - b|).  The reference always has int matrices here; this package hands out uint8
    graphs (closure, transitive_reduction, best_graphs), whose difference would wrap - subtract in int64."""
    return np.sum(np.abs(np.asarray(mat_a).astype(np.int64) - np.asarray(mat_b).astype(np.int64)))



# ---- result post-processing ----------------------------------------------------------------
def dot_string(adj_matrix):
    """The DOT text DAGs/dot.py:28-36 builds for an adjacency matrix: one ``i -> j;`` line per nonzero entry."""
    a = np.asarray(adj_matrix)
    lines = ["digraph {"]
    for i in range(len(a)):
        for j in range(len(a)):
            if a[i][j] != 0:
                lines.append("    {} -> {};".format(i, j))
    return "\n".join(lines) + "\n}"


def generate_dot_from_matrix(adj_matrix, output_file):
    """DAGs/dot.py:28-42 without its ``os.chdir('../')`` side effect: writes the DOT file where it is told to."""
    with open(output_file, "w") as f:
        f.write(dot_string(adj_matrix))
    return output_file


def output_handling(best_dag, output_dir, true_adj=None):
    """main.output_handling (main.py:44-55) without the graphviz rendering: closes the best graph (utils.ancestor,
    :46), reduces it (utils.transitive_reduction, :50), writes ``infer_closed.dot`` and ``infer_red.dot`` into
    `output_dir`, and - when the true network is given - reports the Hamming distance main.py:106 prints, between the closures and between their Hasse diagrams.
    -> dict(closed, reduced, paths, hamming_closed, hamming_reduced)."""
    os.makedirs(output_dir, exist_ok=True)
    best = (np.asarray(best_dag) != 0).astype(int)
    closed = ancestor(best)
    reduced = transitive_reduction(best)
    paths = [generate_dot_from_matrix(closed, os.path.join(output_dir, "infer_closed.dot")),
             generate_dot_from_matrix(reduced, os.path.join(output_dir, "infer_red.dot"))]
    res = {"closed": closed, "reduced": reduced, "paths": paths, "hamming_closed": None, "hamming_reduced": None}
    if true_adj is not None:
        truth = (np.asarray(true_adj) != 0).astype(int)
        truth_closed = ancestor(truth)
        res["hamming_closed"] = int(hamming_distance(closed, truth_closed))
        # canonical forms on both sides: the Hasse diagrams of the two closures
        res["hamming_reduced"] = int(hamming_distance(transitive_reduction(closed), transitive_reduction(truth_closed)))
    return res
