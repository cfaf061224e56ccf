"""Pair-merged score path at S=50, E=10 000: total, score kernels alone (NEMB200_DEBUG bit 128 re-uses the plan of the
previous call), and the planner's phase stamps (bit 64).  usage: python tools/pair_time.py [n] [sweep|masked|full]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, subprocess
if len(sys.argv) > 3:   # child: one mode
    os.environ["NEMB200_DEBUG"] = sys.argv[3]
import nem_b200 as nb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kind = sys.argv[2] if len(sys.argv) > 2 else "sweep"
S, E = 50, 10000
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1), sort_egenes=True)
if kind == "full":
    bits = ds.order_graphs(nb.synth.random_orders(n, S, 1))
elif kind == "sweep":
    rng = np.random.default_rng(1)
    pos = np.argsort(rng.random((n, S)), axis=1)
    cand = rng.random((S, S)) < 0.3
    raw = ((pos[:, :, None] < pos[:, None, :]) & cand[None]).astype(np.uint8)
    rawbits = torch.from_numpy(nb.pack_adj(raw).view(np.int32)).cuda()
    bits = nb.closure(rawbits, S=S, return_bits=True)
else:
    bits = ds.order_graphs(nb.synth.random_orders(n, S, 1), rule=1, cand=(np.random.default_rng(0).random((S, S)) < 0.3).astype(np.uint8))
ref = ds.score_bits(bits, impl="witness").cpu().numpy()
def t(impl, b=None, close=False):
    bits = globals()["bits"] if b is None else b
    got = ds.score_bits(bits, impl=impl, close=close).cpu().numpy()
    close_kw = dict(close=close, inputs_ready=os.environ.get("READY", "1") == "1")
    err = np.abs(got / ref - 1).max()
    for _ in range(5): ds.score_bits(bits, impl=impl, **close_kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(5):
        e0.record()
        for _ in range(20): ds.score_bits(bits, impl=impl, **close_kw)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    return best * 1e3, err
mode = os.environ.get("NEMB200_DEBUG", "0")
us, err = t("pair")
print(f"pair path n={n} {kind} debug={mode}: {us:.1f} us  max rel err {err:.2e}")
if kind == "sweep":
    us, err = t("pair", rawbits, True)
    print(f"   closure fused (planner closes): {us:.1f} us  max rel err {err:.2e}")
if mode == "64":
    ws = ds._score_ws
    tr = ws[ws.numel() - 8 * 8192:].cpu().numpy().view(np.int64)[:9]
    print("   planner phases of CTA 500 (clk since entry):", [int(x - tr[0]) for x in tr])
