"""Round-2 diagnostics: time the score kernel at one shape and check it against the fp64 witness kernel.
usage: python tools/r2_time.py S E n impl [graphs]   (graphs: masked (default) | full | sweep)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n = (int(x) for x in sys.argv[1:4])
impl = sys.argv[4] if len(sys.argv) > 4 else "auto"
kind = sys.argv[5] if len(sys.argv) > 5 else "masked"
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1), sort_egenes=os.environ.get('SORT', '1') == '1')
if kind == "full":
    bits = ds.order_graphs(nb.synth.random_orders(n, S, 1))
elif kind == "sweep":   # the bench's sweep graphs: random order, candidate mask 0.3, closed on the device
    rng = np.random.default_rng(1)
    pos = np.argsort(rng.random((n, S)), axis=1)
    cand = rng.random((S, S)) < 0.3
    raw = ((pos[:, :, None] < pos[:, None, :]) & cand[None]).astype(np.uint8)
    rawbits = torch.from_numpy(nb.pack_adj(raw).view(np.int32)).cuda()
    bits = nb.closure(rawbits, S=S, return_bits=True)
else:
    bits = ds.order_graphs(nb.synth.random_orders(n, S, 1), rule=1, cand=(np.random.default_rng(0).random((S, S)) < 0.3).astype(np.uint8))
ref = ds.score_bits(bits, impl="witness").cpu().numpy()
got = ds.score_bits(bits, impl=impl).cpu().numpy()
err = np.abs(got / ref - 1).max()
for _ in range(5): ds.score_bits(bits, impl=impl)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    K = 20
    e0.record()
    for _ in range(K): ds.score_bits(bits, impl=impl)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / K)
if kind == "sweep":   # the same batch unclosed, closure fused into the score kernel
    for _ in range(3): ds.score_bits(rawbits, impl=impl, close=True)
    torch.cuda.synchronize()
    bestf = 1e9
    for rep in range(5):
        e0.record()
        for _ in range(20): ds.score_bits(rawbits, impl=impl, close=True)
        e1.record(); torch.cuda.synchronize()
        bestf = min(bestf, e0.elapsed_time(e1) / 20)
    print(f"   fused closure+score of the unclosed batch: {bestf*1e3:.1f} us")
envs = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NEMB200_") or k == "SORT")
print(f"S={S} E={E} n={n} impl={impl} graphs={kind} {envs}: {best*1e3:.1f} us  {2.0*S*S*E*n/best/1e9:.0f} TFLOP/s alg  max rel err vs witness {err:.2e}", flush=True)
