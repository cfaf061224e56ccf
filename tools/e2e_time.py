"""End-to-end sweep (Dataset.score_stream) at BASELINE config 3 for several group sizes.  usage: python tools/e2e_time.py [K]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n = 50, 10000, 1024
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1), sort_egenes=True)
rng = np.random.default_rng(1)
pos = np.argsort(rng.random((n, S)), axis=1)
cand = rng.random((S, S)) < 0.3
raw = nb.pack_adj(((pos[:, :, None] < pos[:, None, :]) & cand[None]).astype(np.uint8)).view(np.int32)
host = torch.empty((64, n, S, ds.W), dtype=torch.int32).pin_memory()
for b in range(64): host[b] = torch.from_numpy(raw[rng.permutation(n)])
out = torch.empty((K, n), dtype=torch.float64).pin_memory()
batches = [host[k % 64] for k in range(K)]
for group in (1, 2, 4, 8, 16):
    ds.score_stream(batches[:2 * group], out[:2 * group], group=group); torch.cuda.synchronize()
    res = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); ds.score_stream(batches, out, group=group); t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
        res.append((e0.elapsed_time(e1) / K * 1e3, (t1 - t0) / K * 1e6))
    print(f"group {group:2d}: " + "  ".join(f"{a:.1f} us/step (host issue {b:.1f})" for a, b in res))
