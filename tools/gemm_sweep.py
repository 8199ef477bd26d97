"""Times ams_op_conv1x1 over a (M,K,N) sweep with CUDA events (first-principles look at what bounds the GEMM)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ams_b200 import _native as nat
L = nat.lib()
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
M = 8 * 257 * 513
cases = [(M, 16, 96), (M, 64, 96), (M, 16, 16), (M, 64, 16), (M, 16, 32), (M, 32, 16), (M // 4, 24, 144), (M // 4, 64, 144), (M // 4, 64, 256),
         (M // 4, 144, 24), (M // 16, 192, 32), (M // 16, 32, 192), (M // 64, 960, 160), (M // 64, 160, 960), (M // 64, 384, 64), (M // 64, 64, 384)]
for (m, k, n) in cases:
    a = torch.randn(m, k, device='cuda').to(torch.bfloat16)
    w = torch.randn(n, k, device='cuda').to(torch.bfloat16)
    out = torch.empty(m, n, device='cuda', dtype=torch.bfloat16)
    def run():
        rc = L.ams_op_conv1x1(P(a), P(w), m, n, k, None, None, None, 1, None, 0, P(out), 0, n, 0, None, sp)
        assert rc == 0, nat.last_error()
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    byt = 2.0 * m * (k + n)
    tiles = (m + 127) // 128
    print('M %8d K %4d N %4d : %8.1f us  %7.1f GB/s  per-tile/SM %.2f us' % (m, k, n, us, byt / us / 1e3, us / (tiles / 148.0)))
