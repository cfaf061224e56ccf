"""MCMC loop timing at BASELINE config 3 (1024 chains, S=50, E=10000), three-kernel CUDA-graph loop."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n = 50, 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1))
cand = (np.random.default_rng(7).random((S, S)) < 0.3).astype(np.uint8)
st = nb.ChainState(ds, nb.synth.random_orders(n, S, 5), 2.0 * S / E, seed=1234, rule=1, cand=cand, use_nem=True)
IMPL = os.environ.get("IMPL", "auto")
st.run(64, impl=IMPL); torch.cuda.synchronize()
for iters in (800,):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.run(iters, impl=IMPL); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"chains={n} iters={iters}: {ms/iters*1e3:.1f} us/iteration, {n*iters/ms*1e3:.3e} chain-iterations/s, accept {st.accept.sum().item()/(n*st.iters_done):.3f}")
