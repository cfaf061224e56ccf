"""Order-MCMC iterations/s at the small BASELINE configs (1: S=10,E=100,1 chain; 2: S=20,E=1000,64 chains),
persistent fused kernel vs the CUDA-graph loop, next to the C oracle's CPU loop (one core)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
from oracle import c_oracle

for (S, E, n, iters) in [(10, 100, 1, 10000), (20, 1000, 64, 4000), (39, 453, 64, 2000)]:
    _, _, D = nb.synth.dataset(S, E, seed=0)
    A, B = nb.synth.nem_constants(0.05, 0.1)
    ds = nb.Dataset(D, A, B)
    cand = (np.random.default_rng(0).random((S, S)) < 0.4).astype(np.uint8)
    orders = nb.synth.random_orders(n, S, 1)
    gamma = 2.0 * S / E
    for impl in ("fused", "lut"):
        st = nb.ChainState(ds, orders, gamma, seed=1234, rule=1, cand=cand, use_nem=True)
        st.run(64, impl=impl); torch.cuda.synchronize()
        t0 = time.perf_counter(); st.run(iters, impl=impl); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"S={S} E={E} chains={n} {impl:6s}: {n*iters/dt:12.0f} chain-iterations/s ({dt/iters*1e6:8.2f} us/iteration, accept {st.accept.sum().item()/(n*st.iters_done):.3f})", flush=True)
    t0 = time.perf_counter()
    r = c_oracle.mcmc_chain(D, A, B, orders[0], min(iters, 2000), gamma, seed=1234, chain=0, rule=1, cand=cand, use_nem=True)
    dt = time.perf_counter() - t0
    print(f"S={S} E={E} C oracle, 1 core, 1 chain: {min(iters,2000)/dt:12.0f} iterations/s", flush=True)
