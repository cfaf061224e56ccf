"""Does other stream activity in the process change the two-group MCMC loop's overlap? (diagnostics)"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nem_b200 as nb
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
S, E, n = 50, 10000, 1024
_, _, D = nb.synth.dataset(S, E, seed=0)
ds = nb.Dataset(D, *nb.synth.nem_constants(0.05, 0.1), sort_egenes=True)
cand = (np.random.default_rng(7).random((S, S)) < 0.3).astype(np.uint8)
if mode in ("streams", "all"):
    hb = [torch.from_numpy(nb.pack_adj((np.random.default_rng(k).random((n, S, S)) < 0.05).astype(np.uint8)).view(np.int32)).pin_memory() for k in range(8)]
    out = torch.zeros((8, n), dtype=torch.float64).pin_memory()
    for _ in range(20):
        ds.score_stream(hb, out)
    torch.cuda.synchronize()
stop = False
if mode in ("nvml", "all"):
    def poll():
        import pynvml as nv
        nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
        while not stop:
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM); nv.nvmlDeviceGetCurrentClocksEventReasons(h); time.sleep(0.02)
    threading.Thread(target=poll, daemon=True).start()
st = nb.ChainState(ds, nb.synth.random_orders(n, S, 5), 2.0 * S / E, seed=1234, rule=1, cand=cand, use_nem=True)
st.run(64); torch.cuda.synchronize()
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.run(800); e1.record(); torch.cuda.synchronize()
    print(mode, rep, f"{e0.elapsed_time(e1)/800*1e3:.1f} us/iteration")
stop = True
