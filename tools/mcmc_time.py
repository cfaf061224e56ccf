"""MCMC loop timing at BASELINE config 3 (S=50, E=10 000, 1024 chains, masked rule + closure) per score implementation.
usage: python tools/mcmc_time.py [impl ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n, iters = 50, 10000, 1024, 96
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1), sort_egenes=True)
cand = (np.random.default_rng(0).random((S, S)) < 0.3).astype(np.uint8)
orders = nb.synth.random_orders(n, S, 1)
for impl in (sys.argv[1:] or ["auto", "pair"]):
    st = ds.run_chains(orders, 8, 0.01, seed=3, rule=1, cand=cand, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record(); st.run(iters, impl=impl); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    print(f"impl={impl}: {best*1e3:.1f} us per iteration of {n} chains ({n/best*1e3/1e6:.2f} M chain-iterations/s), accept {float(st.accept.float().mean())/ (8+3*iters):.3f}, best {float(st.best_scores.max()):.3f}")
