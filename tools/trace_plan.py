"""Phase time line of the pair planner (k_pair_plan, CTA 500) - diagnostics (NEMB200_DEBUG=64)."""
import sys, os
os.environ["NEMB200_DEBUG"] = "64"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n = 50, 10000, 1024
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1))
bits = ds.order_graphs(nb.synth.random_orders(n, S, 1), rule=1, cand=(np.random.default_rng(0).random((S, S)) < 0.3).astype(np.uint8))
for _ in range(3): ds.score_bits(bits, impl="pair")
torch.cuda.synchronize()
ws = ds._score_ws
need = ds._score_ws.numel()  # the diagnostics time line sits at the end of the workspace that was used
tr = ws[need - 8 * 8192: need - 8 * 8192 + 9 * 8].cpu().numpy().view(np.int64)
names = ["load", "anc", "count", "rank", "costs", "greedy", "select", "emit", "tables"]
print(" ".join(f"{names[i]}={tr[i + 1] - tr[i]}" for i in range(8)), "total", tr[8] - tr[0])
