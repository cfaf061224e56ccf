"""Pair-merged score path at the config-3 shape: total time, planner phases and fallback count (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n = 50, 10000, 1024
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1), sort_egenes=True)
bits = ds.order_graphs(nb.synth.random_orders(n, S, 1), rule=1, cand=(np.random.default_rng(0).random((S, S)) < 0.3).astype(np.uint8))
w = ds.score_bits(bits, impl="lut").cpu().numpy()
for impl in ("pair", "lut"):
    for _ in range(5): got = ds.score_bits(bits, impl=impl)
    torch.cuda.synchronize()
    if impl == "pair":
        n_fb = int(ds._score_ws[:4].view(torch.int32).item())
        err = float(np.abs(got.cpu().numpy() / w - 1).max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ds.score_bits(bits, impl=impl)
    e1.record(); torch.cuda.synchronize()
    print(f"variant={os.environ.get('NEMB200_LIB_VARIANT','')} impl={impl}: {e0.elapsed_time(e1)/50*1e3:.1f} us" + (f" fallback graphs {n_fb}/{n} max rel err vs lut {err:.1e}" if impl == "pair" else ""))
