"""Fused persistent MCMC kernel vs the CUDA-graph loop around the tensor-core score, over problem sizes (diagnostics:
where mcmc_fused_preferred should draw the line)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
for (S, E, n, iters) in [(30, 5000, 64, 400), (50, 10000, 64, 200), (50, 10000, 296, 200), (20, 1000, 296, 2000), (20, 1000, 1024, 1000), (50, 2000, 296, 400), (40, 453, 592, 1000)]:
    _, _, D = nb.synth.dataset(S, E, seed=0)
    ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1), sort_egenes=True)
    cand = (np.random.default_rng(0).random((S, S)) < 0.4).astype(np.uint8)
    orders = nb.synth.random_orders(n, S, 1)
    for impl in ("fused", "lut"):
        st = nb.ChainState(ds, orders, 2.0 * S / E, seed=1234, rule=1, cand=cand, use_nem=True)
        st.run(32, impl=impl); torch.cuda.synchronize()
        t0 = time.perf_counter(); st.run(iters, impl=impl); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"S={S} E={E} chains={n} {impl:6s}: {n*iters/dt:12.0f} chain-iterations/s ({dt/iters*1e6:8.2f} us/iteration)", flush=True)
print("--- auto")
for (S, E, n, iters) in [(30, 5000, 64, 400), (50, 10000, 64, 200), (20, 1000, 1024, 1000), (50, 2000, 296, 400), (20, 1000, 8192, 200), (10, 100, 4096, 1000)]:
    _, _, D = nb.synth.dataset(S, E, seed=0)
    ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1), sort_egenes=True)
    cand = (np.random.default_rng(0).random((S, S)) < 0.4).astype(np.uint8)
    orders = nb.synth.random_orders(n, S, 1)
    for impl in ("auto", "fused", "lut"):
        st = nb.ChainState(ds, orders, 2.0 * S / E, seed=1234, rule=1, cand=cand, use_nem=True)
        st.run(32, impl=impl); torch.cuda.synchronize()
        t0 = time.perf_counter(); st.run(iters, impl=impl); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"S={S} E={E} chains={n} {impl:6s}: {n*iters/dt:12.0f} chain-iterations/s ({dt/iters*1e6:8.2f} us/iteration)", flush=True)
