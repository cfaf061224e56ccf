"""BASELINE config 4: S=100 x E=100 000, 4096 order-MCMC chains sharded over the GPUs of one box (torchrun),
edge-count all-reduce + best-graph broadcast at the end of the run.  One JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4_mcmc.py [iters]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import nem_b200 as nb
from nem_b200 import synth, parallel

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
S, E, CH = 100, 100000, 4096
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 64
adj, ends, _ = synth.dataset(S, 8, seed=0)                      # true graph only; the data are synthesised on the device
ends = np.random.default_rng(0).integers(0, S, size=E)
D = synth.device_knockdown(adj, ends, 0.05, 0.1, seed=0, device=dev)
ds = nb.Dataset(D, *synth.nem_constants(0.05, 0.1), device=dev, sort_egenes=True)
cand = (np.random.default_rng(7).random((S, S)) < 0.2).astype(np.uint8)
orders = synth.random_orders(CH, S, seed=5)
lo, hi = parallel.shard_range(CH, rank, world)
st = nb.ChainState(ds, orders[lo:hi], 2.0 * S / E, edge_counts=True, chain_offset=lo, seed=1234, rule=nb.RULE_MASKED,
                   cand=cand, use_nem=True)
st.run(32)
parallel.allreduce_edge_counts(st.edge_count_matrix())
parallel.global_best(st.best_scores, st.best_orders, st.best_bits, lo)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
st.run(iters)
counts = parallel.allreduce_edge_counts(st.edge_count_matrix())
best = parallel.global_best(st.best_scores, st.best_orders, st.best_bits, lo)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(ms.item())
    print(json.dumps({"config": "BASELINE config 4: S=100 x E=100000, 4096 chains", "n_gpus": world, "chains_per_gpu": hi - lo,
                      "iters": iters, "ms_per_iteration": ms / iters, "chain_iterations_per_s": CH * iters / (ms * 1e-3),
                      "graph_scores_tflops_algorithmic": 2.0 * S * S * E * CH * iters / (ms * 1e-3) / 1e12,
                      "best_score": best[0], "edge_count_total": int(counts.sum().item()),
                      "score_impl": f"multi-level exp table ({ds.lutm_levels} levels)" if ds.lutm_levels else "ex2 two-pass", "data": "synthetic, knock-downs synthesised on the device"}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
