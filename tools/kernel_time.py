"""Time the score kernel alone at a given shape (diagnostics)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (50, 10000, 1024)))
_, _, D = nb.synth.dataset(S, E, seed=0)
if os.environ.get('ZEROD'): D = D * 0
if os.environ.get('RANDD'): D = (np.random.default_rng(0).random(D.shape) < 0.5).astype(np.uint8)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1), sort_egenes=os.environ.get('SORT','1')=='1')
bits = ds.order_graphs(nb.synth.random_orders(n, S, 1), rule=1, cand=(np.random.default_rng(0).random((S, S)) < 0.3).astype(np.uint8))
if os.environ.get('GRAPHS') == 'empty': bits.zero_()
if os.environ.get('GRAPHS') == 'full': bits = ds.order_graphs(nb.synth.random_orders(n, S, 1))
impl = sys.argv[4] if len(sys.argv) > 4 else "auto"
for _ in range(5): ds.score_bits(bits, impl=impl)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 50
e0.record()
for _ in range(K): ds.score_bits(bits, impl=impl)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"S={S} E={E} n={n} impl={impl} debug={os.environ.get('NEMB200_DEBUG','0')} nomix={os.environ.get('NEMB200_NO_MIX','-')}: {ms*1e3:.1f} us  -> {2.0*S*S*E*n/ms/1e9:.1f} TFLOP/s alg")
