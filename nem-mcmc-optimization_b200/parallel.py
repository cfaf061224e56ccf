"""Multi-GPU sharding of independent chains / graphs.  One process per GPU
(torchrun); chains are split in contiguous blocks, every chain keeps its GLOBAL id as the Philox
counter word so results do not depend on the number of GPUs.  The only collectives are the end-of-run
all-reduce of the S x S posterior edge-count matrix and the arg-max of the best score (NCCL over
NVLink on GPUs, gloo in the CPU tests); nothing is exchanged inside the MCMC loop."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int | None = None, world_size: int | None = None):
    """Contiguous block [lo, hi) of `n_items` owned by `rank` (sizes differ by at most one)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_dataset(D, A=None, B=None, src: int = 0, device=None, sort_egenes: bool = False):
    """the data matrix is broadcast ONCE at start, every rank then builds its own resident Dataset.
    On rank `src`, D is the observed knock-down matrix (S,E) 0/1 (numpy array or torch tensor, host or device) and
    A, B the NEM constants (nem.py:17-18); the other ranks pass D=None.  One broadcast of a 4-number header and one
    of the bit-packed matrix (S*ceil(E/8) bytes: 1.25 MB at S=100, E=100 000) - NCCL over NVLink on GPUs, gloo in
    the CPU tests.  Returns the engine.Dataset (or, with device="cpu", the (S,E) uint8 array itself, for host-only use)."""
    rank, ws = world()
    on_cpu = device is not None and str(device) == "cpu"
    # the collective runs where the backend lives: device memory for NCCL, host memory for gloo / a single process
    nccl = ws > 1 and dist.get_backend() == "nccl"
    backend_dev = torch.device(device if device is not None else "cuda") if (nccl and not on_cpu) else torch.device("cpu")
    if rank == src:
        d = D.detach().cpu().numpy() if isinstance(D, torch.Tensor) else np.asarray(D)
        S, E = d.shape
        head = torch.tensor([float(S), float(E), float(A), float(B)], dtype=torch.float64, device=backend_dev)
    else:
        head = torch.zeros(4, dtype=torch.float64, device=backend_dev)
    if ws > 1:
        dist.broadcast(head, src=src)
    S, E, A, B = int(head[0].item()), int(head[1].item()), float(head[2].item()), float(head[3].item())
    nb = (E + 7) // 8
    if rank == src:
        packed = torch.from_numpy(np.packbits(d.astype(np.uint8) != 0, axis=1)).to(backend_dev)
    else:
        packed = torch.empty((S, nb), dtype=torch.uint8, device=backend_dev)
    if ws > 1:
        dist.broadcast(packed, src=src)
    d = np.unpackbits(packed.cpu().numpy(), axis=1, count=E)
    if on_cpu:
        return d, A, B
    from . import engine
    return engine.Dataset(d, A, B, device=device, sort_egenes=sort_egenes)


def allreduce_edge_counts(counts: torch.Tensor) -> torch.Tensor:
    """SUM all-reduce of the int64 (S,S) edge-count matrix (in place)."""
    if world()[1] > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def allreduce_sum_scalar(x: float, device) -> float:
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def global_best(best_scores: torch.Tensor, best_orders: torch.Tensor, best_bits: torch.Tensor, chain_lo: int = 0):
    """Best chain over all ranks: all-gather of (local best score), broadcast of the winner's order and
    bit rows.  Returns (score, global chain id, order (S,), bits (S,W)) identical on every rank."""
    rank, ws = world()
    k = int(torch.argmax(best_scores).item()) if best_scores.numel() else 0
    local = torch.tensor([float(best_scores[k]) if best_scores.numel() else float("-inf"), float(chain_lo + k)],
                         dtype=torch.float64, device=best_scores.device)
    if ws == 1:
        return float(local[0]), int(local[1]), best_orders[k].clone(), best_bits[k].clone()
    gathered = [torch.empty_like(local) for _ in range(ws)]
    dist.all_gather(gathered, local)
    scores = torch.stack([g[0] for g in gathered])
    src = int(torch.argmax(scores).item())  # ties -> lowest rank, deterministic
    order = best_orders[k].clone() if rank == src else torch.empty_like(best_orders[0])
    bits = best_bits[k].clone() if rank == src else torch.empty_like(best_bits[0])
    dist.broadcast(order, src=src)
    dist.broadcast(bits, src=src)
    return float(gathered[src][0]), int(gathered[src][1]), order, bits


def run_chains_sharded(ds, orders0_all, n_iters, gamma_all, **kw):
    """Run this rank's block of a global set of chains and reduce the summaries over all ranks.
    orders0_all: (n_total,S) start orders, gamma_all scalar or (n_total,).  Returns a dict identical on
    every rank: best_score, best_chain, best_order, best_graph_bits, edge_counts (S,S) int64,
    n_accept (total), plus `state` (this rank's ChainState)."""
    n_total = len(orders0_all)
    lo, hi = shard_range(n_total)
    gam = np.broadcast_to(np.asarray(gamma_all, dtype=np.float64), (n_total,))[lo:hi]
    state = ds.run_chains(np.asarray(orders0_all)[lo:hi], n_iters, gam, chain_offset=lo, **kw)
    counts = allreduce_edge_counts(state.edge_count_matrix())
    score, chain, order, bits = global_best(state.best_scores, state.best_orders, state.best_bits, lo)
    n_acc = allreduce_sum_scalar(float(state.accept.sum().item()), ds.device)
    return dict(best_score=score, best_chain=chain, best_order=order, best_graph_bits=bits, edge_counts=counts,
                n_accept=int(n_acc), state=state, shard=(lo, hi))
