"""BASELINE config 2 (S=20, E=1000, 64 chains) on the persistent fused MCMC kernel: one timed run (diagnostics / ncu target)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n, iters = 20, 1000, 64, 2000
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1))
cand = (np.random.default_rng(0).random((S, S)) < 0.4).astype(np.uint8)
st = nb.ChainState(ds, nb.synth.random_orders(n, S, 1), 2.0 * S / E, seed=1234, rule=1, cand=cand, use_nem=True)
st.run(64); torch.cuda.synchronize()
t0 = time.perf_counter(); st.run(iters); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"S={S} E={E} chains={n}: {n*iters/dt:.0f} chain-iterations/s ({dt/iters*1e6:.2f} us/iteration)")
