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
