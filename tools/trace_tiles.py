"""Dump the clock64 time line of CTA 0 (NEMB200_DEBUG=32) - diagnostics."""
import sys, os
os.environ["NEMB200_DEBUG"] = str(32 | int(os.environ.get("NEMB200_DEBUG", "0")))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n = 50, 10000, 1024
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1))
bits = ds.order_graphs(nb.synth.random_orders(n, S, 1), rule=1, cand=(np.random.default_rng(0).random((S, S)) < 0.3).astype(np.uint8))
for _ in range(3): ds.score_bits(bits)
torch.cuda.synchronize()
ws = ds._score_ws
need = ds._score_ws.numel()  # the diagnostics time line sits at the end of the workspace that was used
tr_off = need - 8 * 8192
ws[tr_off:tr_off + 8 * 8192].zero_()
ds.score_bits(bits); torch.cuda.synchronize()
tr = ws[tr_off:tr_off + 8 * 8192].cpu().numpy().view(np.int64)
roles = tr[:4096].reshape(4, 256, 4)
ep = tr[4096:4096 + 4096].reshape(2, 256, 8)
t0 = roles[1, 0, 0]
f = lambda x: f"{x - t0:7d}" if x else "      -"
print("tile | mma: start afull-ok tempty-ok committed | warp4: top A-ready mathA-done B-ready released tfull-ok ldA-issued mathB-done | warp13: same")
for t in list(range(42, 58)):
    print(f"{t:3d} | " + " ".join(f(roles[1, t, k]) for k in range(4)) + " | " + " ".join(f(ep[0, t, k]) for k in range(8)) + " | " + " ".join(f(ep[1, t, k]) for k in range(8)))
d = np.diff(ep[0, 25:135], axis=0).mean(axis=0)
print("mean period per event (warp4):", d)
seg = (ep[0, 25:135, 1:] - ep[0, 25:135, :-1]).mean(axis=0)
print("mean segment lengths warp4 [A-wait, mathA, B-wait, release, tfull-wait+ldA, (gap), mathB]:", seg)
seg = (ep[1, 25:135, 1:] - ep[1, 25:135, :-1]).mean(axis=0)
print("mean segment lengths warp13:", seg)
