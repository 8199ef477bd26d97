"""Runs the block-fused inverted-residual kernel alone (block 8 shape by default) for ncu / timing.
usage: python tools/micro/fused_block_run.py [cin cout dil n h w reps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ams_b200 import _native as nat
a = [int(x) for x in sys.argv[1:]]
cin, cout, dil, n, h, w, reps = (a + [64, 64, 1, 8, 33, 65, 20][len(a):])
cexp = 6 * cin
L = nat.lib()
dev = 'cuda'
g = torch.Generator(device='cpu').manual_seed(0)
x = torch.randn(n, h, w, cin, generator=g).to(dev, torch.float16)
we = (torch.randn(cexp, cin, generator=g) * (2.0 / cin) ** 0.5).to(dev, torch.float16)
wp = (torch.randn(cout, cexp, generator=g) * (2.0 / cexp) ** 0.5).to(dev, torch.float16)
wplo = (torch.randn(cout, cexp, generator=g) * 1e-4).to(dev, torch.float16) if cout <= 256 else None
wd = (torch.randn(3, 3, cexp, generator=g) * 0.4).to(dev)
v = lambda c, b=0.0: (torch.rand(c, generator=g) * 0.5 + 0.5 + b).to(dev)
s1, t1, s2, t2, s3, t3 = v(cexp), v(cexp), v(cexp), v(cexp), v(cout), v(cout)
out = torch.empty(n, h, w, cout, dtype=torch.float16, device=dev)      # dil >= 10 on the command line = stride 2 (output uses the top-left quarter)
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    rc = L.ams_op_fused_block(P(x), n, h, w, cin, cexp, cout, dil if dil < 10 else 1, 2 if dil >= 10 else 1, P(we), None, P(s1), P(t1), P(wd), P(s2), P(t2), P(wp), P(wplo), P(s3), P(t3),
                              1 if (cin == cout and dil < 10) else 0, P(out), sp)
    assert rc == 0, nat.last_error()
run()
tl = torch.zeros(3 * 64 * 4, dtype=torch.int64, device=dev)
L.ams_debug_fused_timeline(P(tl)); run(); L.ams_debug_fused_timeline(None)
t = tl.cpu().numpy().reshape(3, 64, 4)
t0 = t[t > 0].min()
nchunks = (cexp + 127) // 128
print('timeline of CTA 0 (us since its first event); MMA: expand first/last issue, project first/last issue | compute: chunk start, D1 ready, before the depthwise rows, done')
for g in range(min(64, max(6, 2 * nchunks))):
    f = lambda v: '%7.2f' % ((v - t0) / 1e3) if v > 0 else '    -  '
    print('chunk %2d  MMA expand %s %s project %s %s | compute %s %s %s %s | project unit 0: before wait %s after wait %s MMAs issued %s committed %s' % (
        g, f(t[0, g, 0]), f(t[0, g, 2]), f(t[0, g, 1]), f(t[0, g, 3]), f(t[1, g, 0]), f(t[1, g, 1]), f(t[1, g, 3]), f(t[1, g, 2]),
        f(t[2, g, 0]), f(t[2, g, 1]), f(t[2, g, 2]), f(t[2, g, 3])))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(reps):
    run()
e1.record(); torch.cuda.synchronize()
print('fused block cin %d cout %d dil %d, %dx%dx%d: %.1f us per call (incl. the parameter fill kernel and a stream sync)' % (cin, cout, dil, n, h, w, e0.elapsed_time(e1) * 1e3 / reps))
