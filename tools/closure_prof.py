"""Closure kernels at the BASELINE shapes for profiling: config 5 (S=30, 65 536 graphs -> k_closure_thread<30>), the same
kernel on 1 M graphs, and config 3's batch (S=50, 1024 graphs -> k_closure<2>).  Prints CUDA-event times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import nem_b200 as nb
L = nb._lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
reps = 1 if (len(sys.argv) > 1 and sys.argv[1] == 'prof') else 10   # 'prof': one warm + one timed launch per shape (short ncu capture)
for S, n in [(30, 65536), (30, 1 << 20), (50, 1024), (50, 1 << 19)]:
    W = (S + 31) // 32
    raw = torch.from_numpy(nb.synth.random_dag_bits(min(n, 4096), S, seed=1).view(np.int32)).cuda().repeat(max(n // 4096, 1), 1, 1)[:n].contiguous()
    out = torch.empty_like(raw)
    for _ in range(1 if reps == 1 else 3):
        nb._lib.check(L.nemb200_closure_u32(C.c_void_p(raw.data_ptr()), n, S, C.c_void_p(out.data_ptr()), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        nb._lib.check(L.nemb200_closure_u32(C.c_void_p(raw.data_ptr()), n, S, C.c_void_p(out.data_ptr()), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"S={S} n={n}: {ms*1e3:.2f} us  {2.0*S*W*4*n/ms/1e6:.0f} GB/s algorithmic  {n/ms/1e6:.3f} G closures/s", flush=True)
