"""Synthetic inputs with the reference generator's distribution (DAGs/rnd_dag_gen.py:47-103), at any
configured (S, E) - SURVEY 8d: rank-layered DAG, nodes numbered topologically, each (older node, new node)
edge kept iff randint(0,100) < 40; attachment end_nodes[e] ~ U{0..S-1}; data flipped with the per-cell rule
of utils.py:25-35 (vectorised PCG64 for large E; the MT19937 path of the goldens is utils.create_*)."""
from __future__ import annotations

import numpy as np

_LAYOUT = {10: (2, 5), 20: (4, 5), 30: (5, 6), 50: (5, 10), 100: (10, 10)}


def rank_layout(S: int):
    """(nodes per rank p, ranks r) with p*r == S (the min/max arguments of gen_rnd_dag that pin S)."""
    if S in _LAYOUT:
        return _LAYOUT[S]
    p = max(d for d in range(1, int(S ** 0.5) + 1) if S % d == 0)
    return p, S // p


def random_dag(S: int, rng: np.random.Generator, percent: int = 40) -> np.ndarray:
    """Unclosed adjacency (adj[a][b]=1 iff edge a->b) drawn with rnd_dag_gen.py:63-79's edge law."""
    p, r = rank_layout(S)
    adj = np.zeros((S, S), dtype=np.uint8)
    for k in range(1, r):
        old = k * p
        adj[:old, old:old + p] = rng.integers(0, 101, size=(old, p)) < percent
    return adj


def closure_host(adj: np.ndarray) -> np.ndarray:
    """Host-side boolean closure for preparing the *true* network of a synthetic data set (input
    preparation only; candidate graphs are closed on the GPU)."""
    r = adj.astype(bool).copy()
    for k in range(r.shape[0]):
        r |= np.outer(r[:, k], r[k, :])
    return r.astype(np.uint8)


def dataset(S: int, E: int, seed: int = 0, alpha: float = 0.05, beta: float = 0.1):
    """-> (closed true adjacency, end_nodes, observed 0/1 matrix D of shape (S,E))."""
    rng = np.random.default_rng(seed)
    adj = closure_host(random_dag(S, rng))
    end_nodes = rng.integers(0, S, size=E)
    reach = adj.T.copy()
    reach[np.arange(S), np.arange(S)] = 1          # utils.py:5-13
    real = np.ascontiguousarray(reach[end_nodes].T)  # utils.py:15-23
    u = rng.random(real.shape)
    D = real.copy()
    D[(real == 0) & (u < alpha)] = 1
    D[(real == 1) & (u < beta)] = 0
    return adj, end_nodes, D.astype(np.uint8)


def nem_constants(alpha: float, beta: float):
    """nem.py:17-18."""
    return float(np.log(alpha / (1.0 - beta))), float(np.log(beta / (1.0 - alpha)))


def random_orders(n: int, S: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.argsort(rng.random((n, S)), axis=1).astype(np.int32)


def random_dag_bits(n: int, S: int, seed: int = 0, percent: int = 40) -> np.ndarray:
    """n independent unclosed DAGs with the generator's edge law on a fixed topological labelling, as
    bit rows (n,S,W) uint32 (BASELINE config 5)."""
    from .engine import pack_adj
    rng = np.random.default_rng(seed)
    p, r = rank_layout(S)
    rank = np.arange(S) // p
    allowed = rank[:, None] < rank[None, :]
    out = np.empty((n, S, (S + 31) // 32), dtype=np.uint32)
    step = 4096
    for lo in range(0, n, step):
        m = min(step, n - lo)
        a = (rng.integers(0, 101, size=(m, S, S)) < percent) & allowed[None]
        out[lo:lo + m] = pack_adj(a.astype(np.uint8))
    return out


def device_knockdown(adj, end_nodes, alpha: float = 0.05, beta: float = 0.1, seed: int = 0, device=None, want_real=False):
    """Real and observed knock-down matrices generated on the GPU (nemb200_synth_knockdown): the data model of
    utils.py:15-35 with Philox flips.  adj: (S,S) closed true graph; end_nodes: (E,) attachments.
    Returns the observed matrix as a uint8 CUDA tensor (S,E) (and the noise-free one if want_real)."""
    import ctypes as C
    import torch
    from . import _lib, engine
    dev = engine._require_cuda(device)
    adj = np.asarray(adj)
    S, E = adj.shape[0], len(end_nodes)
    bits = torch.from_numpy(engine.pack_adj(adj).view(np.int32)).to(dev)
    ends = torch.from_numpy(np.ascontiguousarray(np.asarray(end_nodes, dtype=np.int32))).to(dev)
    obs = torch.empty((S, E), dtype=torch.uint8, device=dev)
    real = torch.empty((S, E), dtype=torch.uint8, device=dev) if want_real else None
    st = torch.cuda.current_stream(dev)
    _lib.check(_lib.lib().nemb200_synth_knockdown(engine._ptr(bits), engine._ptr(ends), S, E, float(alpha), float(beta),
                                                 C.c_uint64(int(seed)), engine._ptr(real), engine._ptr(obs),
                                                 engine._stream_ptr(st)))
    return (obs, real) if want_real else obs
