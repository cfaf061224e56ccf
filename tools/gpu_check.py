"""Quick staged sanity check on the GPU box (prints details instead of asserting)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
from oracle import c_oracle, nem_oracle as orc

def stage(name):
    print(f"--- {name}", flush=True)

stage("closure")
rng = np.random.default_rng(0)
for S in (10, 50, 100):
    a = (rng.random((64, S, S)) < 0.05).astype(np.uint8)
    got = nb.closure(a)
    bad = sum(not np.array_equal(got[g], c_oracle.closure(a[g])) for g in range(64))
    print("S", S, "closure mismatches", bad, flush=True)

for (S, E, n) in [(10, 100, 33), (50, 1000, 23), (20, 1000, 64), (100, 1500, 7)]:
    stage(f"score S={S} E={E} n={n}")
    _, _, D = orc.synth_dataset(S, E, seed=1)
    A, B = orc.nem_constants(0.05, 0.1)
    graphs = np.stack([orc.order_to_graph(rng.permutation(S), 1, (rng.random((S, S)) < 0.3), True) for _ in range(n)]).astype(np.uint8)
    ref = c_oracle.score_batch(graphs, D, A, B, nthreads=8)
    ds = nb.Dataset(D, A, B)
    print("lut_valid", ds.lut_valid, "scale", ds.scale_log2, "G", ds.graphs_per_tile, flush=True)
    for impl in ["witness", "ex2_2pass", "ex2", "lut"]:
        if impl == "lut" and not ds.lut_valid: continue
        if impl == "ex2" and S > 54: continue
        t0 = time.time()
        try:
            got = ds.score_graphs(graphs, impl=impl)
            err = np.abs(got / ref - 1).max()
            print(f"  {impl:10s} max rel err {err:.3e}  first got {got[:3]} ref {ref[:3]}  ({time.time()-t0:.3f}s)", flush=True)
        except Exception as ex:
            print(f"  {impl:10s} FAILED {ex}", flush=True)

stage("mcmc")
S, E = 13, 126
_, _, D = orc.synth_dataset(S, E, seed=7)
A, B = orc.nem_constants(0.05, 0.1)
ds = nb.Dataset(D, A, B)
orders0 = np.stack([rng.permutation(S) for _ in range(4)]).astype(np.int32)
st = ds.run_chains(orders0, 200, 2.0 * S / E, swap_prob=0.9, seed=99, burn_in=20, trace=True)
res = st.result()
for c in range(4):
    ref = c_oracle.mcmc_chain(D, A, B, orders0[c], 200, 2.0 * S / E, swap_prob=0.9, seed=99, chain=c, burn_in=20, keep_scores=True)
    print("chain", c, "acc", res.accept_counts[c], ref["n_accept"], "best", res.best_scores[c], ref["best_score"],
          "order eq", np.array_equal(res.curr_orders[c], ref["curr_order"]), flush=True)
print("done", flush=True)
