"""Closure kernel alone on large batches (diagnostics): GB/s of algorithmic traffic (read + write of the bit rows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import nem_b200 as nb
L = nb._lib.lib()
for S, n in [(30, 65536), (30, 1 << 20), (31, 1 << 20), (20, 1 << 20), (32, 1 << 20), (10, 1 << 21), (50, 1 << 19), (100, 1 << 17)]:
    W = (S + 31) // 32
    raw = torch.from_numpy(nb.synth.random_dag_bits(4096, S, seed=1).view(np.int32)).cuda().repeat(n // 4096, 1, 1).contiguous()
    out = torch.empty_like(raw)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        nb._lib.check(L.nemb200_closure_u32(C.c_void_p(raw.data_ptr()), n, S, C.c_void_p(out.data_ptr()), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        nb._lib.check(L.nemb200_closure_u32(C.c_void_p(raw.data_ptr()), n, S, C.c_void_p(out.data_ptr()), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"S={S} n={n}: {ms*1e3:.1f} us  {2.0*S*W*4*n/ms/1e6:.0f} GB/s  {n/ms/1e6:.2f} G closures/s")
