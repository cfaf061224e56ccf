import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n = 30, 5000, 8192
_, _, D = nb.synth.dataset(S, E, seed=5)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1))
raw = torch.from_numpy(nb.synth.random_dag_bits(n, S, seed=1).view(np.int32)).cuda()
closed = nb.closure(raw, S=S, return_bits=True)
a = ds.score_bits(closed).cpu().numpy()
b = ds.score_bits(closed).cpu().numpy()
print("same batch twice equal:", np.array_equal(a, b), np.abs(a - b).max())
for lo, m in ((0, 2048), (2048, 2048), (0, 1024), (100, 37), (0, 4096)):
    c = ds.score_bits(closed[lo:lo + m].contiguous()).cpu().numpy()
    d = np.abs(c - a[lo:lo + m])
    bad = np.nonzero(d)[0]
    print(f"slice [{lo},{lo+m}): equal {np.array_equal(c, a[lo:lo+m])}  n_diff {len(bad)}  max abs diff {d.max():.3e}  first bad {bad[:8]}  rel {d.max()/np.abs(a).max():.2e}")
w = ds.score_bits(closed[:2048].contiguous(), impl="witness").cpu().numpy()
print("vs witness rel err whole", np.abs(a[:2048] / w - 1).max())
