"""Small end-to-end case for compute-sanitizer (memcheck): dataset, closure, all score epilogues, both MCMC paths."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
for (S, E, n) in [(50, 700, 9), (13, 126, 5), (100, 300, 3)]:
    _, _, D = nb.synth.dataset(S, E, seed=1)
    ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1))
    rng = np.random.default_rng(0)
    cand = (rng.random((S, S)) < 0.3).astype(np.uint8)
    orders = nb.synth.random_orders(n, S, 2)
    bits = ds.order_graphs(orders, rule=1, cand=cand, use_nem=True)
    for impl in ("witness", "ex2_2pass") + (("lut", "ex2") if ds.lut_valid else ()):
        s = ds.score_bits(bits, impl=impl).cpu().numpy()
        print(S, impl, s[:2])
    for impl in (("fused", "lut") if ds.lut_valid else ("auto",)):
        st = ds.run_chains(orders, 24, 2.0 * S / E, rule=1, cand=cand, use_nem=True, impl=impl, trace=True)
        print(S, impl, st.result().best_scores[:2])
print("sanitize case done")
