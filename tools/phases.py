"""Where the fixed part of one score launch goes (development build with -DNEMB200_PHASES): %globaltimer stamps of CTA 0."""
import sys, os
os.environ["NEMB200_DEBUG"] = "32"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
S, E, n = 50, 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1), sort_egenes=True)
bits = ds.order_graphs(nb.synth.random_orders(n, S, 1), rule=1, cand=(np.random.default_rng(0).random((S, S)) < 0.3).astype(np.uint8))
for _ in range(3): ds.score_bits(bits, inputs_ready=True)
torch.cuda.synchronize()
ws = ds._score_ws
tr_off = ws.numel() - 8 * 8192
for rep in range(3):
    ws[tr_off:tr_off + 8 * 8192].zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ds.score_bits(bits, inputs_ready=True); ds.score_bits(bits, inputs_ready=True); e1.record(); torch.cuda.synchronize()
    tr = ws[tr_off:tr_off + 8 * 8192].cpu().numpy().view(np.int64)[8000:8010]
    t0 = tr[0]
    names = ["entry", "prologue done", "first B expanded", "first item ready (epi)", "epilogue warp done", "all warps done", "CTA0 exit", "finalize done (last CTA)"]
    full = ws[tr_off:tr_off + 8 * 8192].cpu().numpy().view(np.int64)
    ends = (full[7600:7600 + 148] - t0) / 1e3
    sms = full[7800:7800 + 148]
    o = np.argsort(ends)
    print("   CTA finish times (us after CTA 0's entry): min %.1f  p25 %.1f  median %.1f  p75 %.1f  max %.1f" % (ends.min(), np.percentile(ends, 25), np.median(ends), np.percentile(ends, 75), ends.max()))
    print("   slowest CTAs (cta, sm, t):", [(int(c), int(sms[c]), round(float(ends[c]), 1)) for c in o[-8:]], " fastest:", [(int(c), int(sms[c]), round(float(ends[c]), 1)) for c in o[:4]])
    if rep == 2:
        Gt = 8 if (ds.pair_ok and 44 <= S <= 50) else 4   # graphs per N tile of the kernel AUTO runs
        T, nm = ((n + Gt - 1) // Gt) * 79, 79
        rows = []
        for c in range(148):
            lo, hi = c * T // 148, (c + 1) * T // 148
            bounds = [b for b in range(((lo // nm) + 1) * nm, hi, nm)]
            rows.append((c, int(sms[c]), hi - lo, len(bounds) + 1, (bounds[0] - lo) if bounds else hi - lo, (hi - bounds[-1]) if bounds else hi - lo, round(float(ends[c]), 1)))
        for k in (2, 3):
            sel = [r[-1] for r in rows if r[3] == k]
            if sel: print(f"   CTAs with {k} items: {len(sel)}, mean finish {np.mean(sel):.1f} (min {min(sel):.1f}, max {max(sel):.1f})")
        sel137 = [r[-1] for r in rows if r[2] == 137]; sel136 = [r[-1] for r in rows if r[2] == 136]
        print(f"   137-tile CTAs: {len(sel137)} mean {np.mean(sel137):.1f}; 136-tile CTAs: {len(sel136)} mean {np.mean(sel136):.1f}")
        print("   by first-item length (tiles: CTAs, mean finish):", [(lo_, len(v), round(float(np.mean(v)), 1)) for lo_, v in sorted({b: [r[-1] for r in rows if r[4] // 10 * 10 == b] for b in set(r[4] // 10 * 10 for r in rows)}.items())])
        by_sm = sorted(rows, key=lambda r: r[1])
        print("   by SM id (sm, items, first, last, finish):", [(r[1], r[3], r[4], r[5], r[6]) for r in by_sm[::6]])
    print(f"launch {rep}: event time {e0.elapsed_time(e1)*1e3:.1f} us; " + "; ".join(f"{nm} +{(tr[k]-t0)/1e3:.1f}" for k, nm in enumerate(names) if tr[k]) + f" (last CTA {tr[8]})")
