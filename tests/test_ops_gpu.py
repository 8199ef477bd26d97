"""Kernel-level parity (GPU box): every CUDA kernel, called through the C ABI's op-level entry points on
torch-owned device buffers, against the oracle's restatement of the same TF op on the same inputs.
Storage types: forward activations / 1x1 weight operands fp16 (AC), activation gradients bf16 (BF).
Tolerances: 16-bit outputs are compared at one ulp of their type (fp16 2^-11, bf16 2^-8 relative) plus fp32
accumulation slack; integer outputs (argmax given logits, confusion matrix, selection mask, Adam in fp32 IEEE ops) are bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

import student_oracle as so
from _util import P, ac_round, bf16_round, call, err_stats, log, stream_ptr
from ams_b200 import _native as nat

pytestmark = pytest.mark.gpu
DEV = 'cuda'
BF = torch.bfloat16
AC = torch.float16
ULP = 2.0 ** -8           # bf16
AULP = 2.0 ** -11         # fp16


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize('M,K,N', [(300, 16, 96), (1000, 24, 144), (257, 144, 24), (4097, 32, 16), (2145, 960, 320),
                                   (2145, 160, 960), (5000, 384, 64), (2145, 320, 256), (129, 576, 160), (640, 64, 384),
                                   (8385, 728, 728), (2145, 1536, 2048), (1100, 2048, 256)])     # teacher: K >= 512 plans (whole SM, wave-aware N tiles)
def test_conv1x1_plain(M, K, N):
    L = nat.lib()
    a = ac_round(rnd(M, K, seed=1))
    w = ac_round(rnd(N, K, seed=2, scale=(2.0 / K) ** 0.5))
    ref = a @ w.t()
    out = torch.full((M, N), float('nan'), dtype=AC, device=DEV)
    call(L.ams_op_conv1x1, P(a.to(DEV, AC)), P(w.to(DEV, AC)), M, N, K, None, None, None, 1, None, 0, P(out), 0, N, 0, None, stream_ptr())
    torch.cuda.synchronize()
    ok, _ = err_stats('conv1x1 plain M%d K%d N%d' % (M, K, N), out, ref, AULP, 2e-4)
    assert ok


@pytest.mark.parametrize('M,K,N,fp32out', [(2145, 960, 160, False), (1000, 96, 24, False), (4097, 32, 16, False), (2145, 320, 256, False),
                                           (2145, 256, 19, True), (700, 16, 96, False)])
def test_conv1x1_split_weights(M, K, N, fp32out):
    """Forward convs with <= 256 output channels: W = hi + lo, both planes fp16, two MMAs per k-step into one fp32
    accumulator.  The result must match the product with the (almost) full-precision weight hi + lo -- and differ
    from the plain-fp16-weight product by far more than the tolerance, or the second plane is not being used."""
    L = nat.lib()
    a = ac_round(rnd(M, K, seed=51))
    w32 = rnd(N, K, seed=52, scale=(2.0 / K) ** 0.5)
    hi = ac_round(w32)
    lo = ac_round(w32 - hi)
    ref = a.double() @ (hi.double() + lo.double()).t()
    ldc = 32 if fp32out else N
    out = torch.full((M, ldc), float('nan'), dtype=torch.float32 if fp32out else AC, device=DEV)
    call(L.ams_op_conv1x1, P(a.to(DEV, AC)), P(hi.to(DEV, AC)), M, N, K, None, None, None, 1, None, 0, P(out), 1 if fp32out else 0, ldc, 0,
         P(lo.to(DEV, AC)), stream_ptr())
    torch.cuda.synchronize()
    got = out[:, :N]
    ok, _ = err_stats('conv1x1 split weights M%d K%d N%d fp32out%d' % (M, K, N, fp32out), got, ref.float(), 1e-6 if fp32out else AULP,
                      2e-5 if fp32out else 2e-4)
    assert ok
    if fp32out:
        plain = (a.double() @ hi.double().t()).float()
        assert float((got.float().cpu() - plain).abs().max()) > 1e-4          # the low plane contributes


@pytest.mark.parametrize('M,K,N,res', [(2145, 320, 960, False), (1000, 24, 144, True), (4097, 16, 32, False), (2145, 32, 256, False)])
def test_conv1x1_data_gradient_types(M, K, N, res):
    """The data-gradient GEMM: bf16 gradient x bf16 copy of the transposed weights -> bf16 gradient (+ bf16 skip gradient).
    (tcgen05.mma kind::f16 wants one element format for both operands: the forward GEMM is fp16 x fp16.)"""
    L = nat.lib()
    a = bf16_round(rnd(M, K, seed=41, scale=0.01))
    w = bf16_round(rnd(N, K, seed=42, scale=(2.0 / K) ** 0.5))
    r = bf16_round(rnd(M, N, seed=43, scale=0.01)) if res else None
    ref = a @ w.t() + (r if res else 0.0)
    out = torch.full((M, N), float('nan'), dtype=BF, device=DEV)
    call(L.ams_op_conv1x1, P(a.to(DEV, BF)), P(w.to(DEV, BF)), M, N, K, None, None, None, 1, P(r.to(DEV, BF)) if res else None, 0,
         P(out), 0, N, 1, None, stream_ptr())
    torch.cuda.synchronize()
    ok, _ = err_stats('conv1x1 dgrad types M%d K%d N%d res%d' % (M, K, N, res), out, ref, ULP, 1e-5)
    assert ok


@pytest.mark.parametrize('M,K,N,act,res,rows', [(1000, 96, 24, 0, True, 0), (2145, 64, 384, 2, False, 0),
                                                (2 * 2145, 256, 256, 1, False, 2145), (777, 192, 32, 0, True, 0)])
def test_conv1x1_epilogue(M, K, N, act, res, rows):
    L = nat.lib()
    a = ac_round(rnd(M, K, seed=3))
    w = ac_round(rnd(N, K, seed=4, scale=(2.0 / K) ** 0.5))
    scale = torch.rand(N) + 0.5
    shift = rnd(N, seed=5)
    r = ac_round(rnd(M, N, seed=6)) if res else None
    rb = rnd(M // rows, N, seed=7) if rows else None
    ref = a @ w.t()
    if rows:
        ref = ref + rb.repeat_interleave(rows, dim=0)
    ref = ref * scale + shift
    ref = {0: ref, 1: ref.clamp_min(0), 2: ref.clamp(0, 6)}[act]
    if res:
        ref = ref + r
    out = torch.full((M, N), float('nan'), dtype=AC, device=DEV)
    call(L.ams_op_conv1x1, P(a.to(DEV, AC)), P(w.to(DEV, AC)), M, N, K, P(scale.to(DEV)), P(shift.to(DEV)),
         P(rb.to(DEV)) if rows else None, max(rows, 1), P(r.to(DEV, AC)) if res else None, act, P(out), 0, N, 0, None, stream_ptr())
    torch.cuda.synchronize()
    ok, _ = err_stats('conv1x1 epilogue M%d K%d N%d act%d res%d rb%d' % (M, K, N, act, res, rows), out, ref, AULP, 4e-4)
    assert ok


def test_conv1x1_logits_fp32():
    L = nat.lib()
    M, K, N = 2145, 256, 19
    a = ac_round(rnd(M, K, seed=8))
    w = ac_round(rnd(N, K, seed=9, scale=0.1))
    bias = rnd(N, seed=10)
    ref = a @ w.t() + bias
    out = torch.full((M, 32), float('nan'), dtype=torch.float32, device=DEV)
    call(L.ams_op_conv1x1, P(a.to(DEV, AC)), P(w.to(DEV, AC)), M, N, K, None, P(bias.to(DEV)), None, 1, None, 0, P(out), 1, 32, 0, None, stream_ptr())
    torch.cuda.synchronize()
    ok, _ = err_stats('conv1x1 logits fp32', out[:, :N], ref, 1e-5, 1e-4)
    assert ok
    assert float(out[:, N:].abs().max()) == 0.0


@pytest.mark.parametrize('M,Cin,Cout', [(70001, 16, 96), (33153, 24, 144), (33153, 144, 24), (4290, 960, 320),
                                        (4290, 160, 960), (8385, 192, 32), (2145, 256, 24), (1000, 320, 256), (100, 32, 16)])
def test_wgrad(M, Cin, Cout):
    L = nat.lib()
    x = ac_round(rnd(M, Cin, seed=11))                        # activations: fp16 in HBM ...
    dz = bf16_round(rnd(M, Cout, seed=12, scale=0.05))        # gradients: bf16
    ref = (bf16_round(x).double().t() @ dz.double()).float()  # ... rewritten as bf16 tile by tile inside the kernel
    dw = torch.full((Cin, Cout), float('nan'), dtype=torch.float32, device=DEV)
    call(L.ams_op_wgrad, P(x.to(DEV, AC)), Cin, P(dz.to(DEV, BF)), Cout, M, P(dw), stream_ptr())
    torch.cuda.synchronize()
    ok, _ = err_stats('wgrad M%d Cin%d Cout%d' % (M, Cin, Cout), dw, ref, 1e-4, 1e-5 * float(ref.abs().max()) + 1e-6)
    assert ok
    # determinism: a second run is bit-identical (fixed-order split-K reduction)
    dw2 = torch.empty_like(dw)
    call(L.ams_op_wgrad, P(x.to(DEV, AC)), Cin, P(dz.to(DEV, BF)), Cout, M, P(dw2), stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(dw, dw2)


def _dw_ref(x, w, stride, dil):
    return so.conv2d_same(x, w.reshape(3, 3, -1, 1), stride, dil, True)


@pytest.mark.parametrize('n,h,w,c,stride,dil', [(2, 33, 65, 32, 1, 1), (1, 65, 129, 96, 2, 1), (2, 17, 33, 960, 1, 2),
                                                (1, 64, 30, 144, 2, 1), (1, 9, 17, 384, 1, 1), (3, 5, 7, 576, 1, 2),
                                                (1, 33, 65, 728, 1, 1), (1, 66, 40, 728, 2, 1), (1, 20, 33, 112, 1, 2)])   # 56-channel chunks (teacher)
def test_depthwise_fwd(n, h, w, c, stride, dil):
    L = nat.lib()
    x = ac_round(rnd(n, h, w, c, seed=13))
    wt = rnd(3, 3, c, seed=14, scale=0.4)
    scale = torch.rand(c) + 0.5
    shift = rnd(c, seed=15, scale=0.5)
    raw = _dw_ref(x, wt, stride, dil)
    ho, wo = raw.shape[1], raw.shape[2]
    out = torch.full((n, ho, wo, c), float('nan'), dtype=AC, device=DEV)
    call(L.ams_op_depthwise, P(x.to(DEV, AC)), P(wt.to(DEV)), n, h, w, c, stride, dil, None, None, 0, P(out), stream_ptr())
    torch.cuda.synchronize()
    ok1, _ = err_stats('depthwise raw %s s%d d%d' % ((n, h, w, c), stride, dil), out, raw, AULP, 2e-4)
    call(L.ams_op_depthwise, P(x.to(DEV, AC)), P(wt.to(DEV)), n, h, w, c, stride, dil, P(scale.to(DEV)), P(shift.to(DEV)), 2, P(out), stream_ptr())
    torch.cuda.synchronize()
    ok2, _ = err_stats('depthwise fold+relu6 %s s%d d%d' % ((n, h, w, c), stride, dil), out, (raw * scale + shift).clamp(0, 6), AULP, 4e-4)
    assert ok1 and ok2


@pytest.mark.parametrize('n,h,w,c,stride,dil', [(2, 33, 65, 32, 1, 1), (1, 65, 129, 96, 2, 1), (2, 17, 33, 960, 1, 2),
                                                (1, 64, 30, 144, 2, 1), (1, 9, 17, 384, 1, 1), (2, 129, 257, 192, 1, 1)])
def test_depthwise_fused_bn_on_load_and_stats(n, h, w, c, stride, dil):
    """Training-mode depthwise: the producer's BN + ReLU6 is applied while the tile is staged (the normalised tensor
    is never materialised; conv zero padding stays zero AFTER the activation) and the batch statistics of the stored
    bf16 output come out of the same kernel."""
    L = nat.lib()
    z = ac_round(rnd(n, h, w, c, seed=23, scale=2.0))
    wt = rnd(3, 3, c, seed=24, scale=0.4)
    sc = torch.rand(c) + 0.5
    sh = rnd(c, seed=25, scale=0.5)
    y = ac_round((z.double() * sc.double() + sh.double()).float().clamp(0, 6))   # fmaf(z, sc, sh): one rounding like the kernel
    raw = _dw_ref(y, wt, stride, dil)
    out = torch.full(raw.shape, float('nan'), dtype=AC, device=DEV)
    stats = torch.full((2, c), float('nan'), dtype=torch.float64, device=DEV)
    call(L.ams_op_depthwise_fused, P(z.to(DEV, AC)), P(wt.to(DEV)), n, h, w, c, stride, dil, P(sc.to(DEV)), P(sh.to(DEV)), 2,
         P(out), P(stats), stream_ptr())
    torch.cuda.synchronize()
    ok, _ = err_stats('depthwise fused (BN on load) %s s%d d%d' % ((n, h, w, c), stride, dil), out, raw, AULP, 4e-4)
    assert ok
    o = out.float().cpu().double().reshape(-1, c)
    assert torch.allclose(stats[0].cpu(), o.sum(0), rtol=1e-6, atol=1e-4)
    assert torch.allclose(stats[1].cpu(), (o * o).sum(0), rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize('n,h,w,c,stride,dil', [(2, 33, 65, 32, 1, 1), (1, 65, 129, 96, 2, 1), (2, 17, 33, 960, 1, 2), (1, 64, 30, 144, 2, 1)])
def test_depthwise_bwd(n, h, w, c, stride, dil):
    L = nat.lib()
    x = ac_round(rnd(n, h, w, c, seed=16)).requires_grad_(True)
    wt = rnd(3, 3, c, seed=17, scale=0.4).requires_grad_(True)
    y = _dw_ref(x, wt, stride, dil)
    dz = bf16_round(rnd(*y.shape, seed=18, scale=0.1))
    y.backward(dz)
    dx = torch.full((n, h, w, c), float('nan'), dtype=BF, device=DEV)
    dw = torch.full((3, 3, c), float('nan'), dtype=torch.float32, device=DEV)
    call(L.ams_op_depthwise_bwd, P(x.detach().to(DEV, AC)), P(dz.to(DEV, BF)), P(wt.detach().to(DEV)), n, h, w, c, stride, dil, P(dx), P(dw), stream_ptr())
    torch.cuda.synchronize()
    ok1, _ = err_stats('depthwise dX %s s%d d%d' % ((n, h, w, c), stride, dil), dx, x.grad, ULP, 1e-4)
    ok2, _ = err_stats('depthwise dW %s s%d d%d' % ((n, h, w, c), stride, dil), dw, wt.grad, 1e-4, 1e-5 * float(wt.grad.abs().max()))
    assert ok1 and ok2


def _fma32(a, b, c):
    """fp32 fmaf(a, b, c) (one rounding), emulated in float64."""
    return (a.double() * b.double() + c.double()).float()


@pytest.mark.parametrize('n,h,w,c,stride,dil', [(2, 33, 65, 32, 1, 1), (1, 65, 129, 96, 2, 1), (2, 17, 33, 960, 1, 2),
                                                (1, 64, 30, 144, 2, 1), (1, 9, 17, 384, 1, 1), (2, 129, 257, 192, 1, 1),
                                                (1, 33, 65, 192, 2, 1), (3, 5, 7, 576, 1, 2)])
def test_depthwise_bwd_fused(n, h, w, c, stride, dil):
    """Fused depthwise backward: BN-backward apply of the depthwise layer while staging (g, z), filter + data gradient,
    the producer's BN + ReLU6 recomputed from its raw output, the activation mask on the stored gradient and the column
    sums for the producer's BN backward -- against autograd through the same stored tensors (fp16 activations, bf16 gradients)."""
    L = nat.lib()
    zin = ac_round(rnd(n, h, w, c, seed=31, scale=2.0))
    isc = torch.rand(c, generator=torch.Generator().manual_seed(32)) + 0.5
    ish = rnd(c, seed=33, scale=0.5)
    pre = _fma32(zin, isc, ish)
    x = ac_round(pre.clamp(0, 6)).requires_grad_(True)
    wt = rnd(3, 3, c, seed=34, scale=0.4).requires_grad_(True)
    y = _dw_ref(x, wt, stride, dil)
    z = ac_round(y.detach())
    g = bf16_round(rnd(*y.shape, seed=35, scale=0.1))
    sc2 = torch.rand(c, generator=torch.Generator().manual_seed(36)) + 0.5
    sh2 = rnd(c, seed=37, scale=0.5) + 1.0
    coef = torch.stack([torch.rand(c, generator=torch.Generator().manual_seed(38)) + 0.5, rnd(c, seed=39, scale=0.01),
                        rnd(c, seed=40, scale=0.01)])
    yh = _fma32(z, sc2, sh2)
    gm = torch.where((yh > 0) & (yh < 6), g, torch.zeros_like(g))
    gz = bf16_round(_fma32(coef[0].expand_as(gm), gm, _fma32(coef[1].expand_as(z), z, coef[2].expand_as(z))))
    y.backward(gz)
    passm = (pre > 0) & (pre < 6)
    gout_ref = torch.where(passm, bf16_round(x.grad), torch.zeros_like(pre))
    gout = torch.full((n, h, w, c), float('nan'), dtype=BF, device=DEV)
    dw = torch.full((3, 3, c), float('nan'), dtype=torch.float32, device=DEV)
    sums = torch.full((2, c), float('nan'), dtype=torch.float64, device=DEV)
    call(L.ams_op_depthwise_bwd_fused, P(g.to(DEV, BF)), P(z.to(DEV, AC)), P(sc2.to(DEV)), P(sh2.to(DEV)), 2, P(coef.to(DEV)),
         P(zin.to(DEV, AC)), P(isc.to(DEV)), P(ish.to(DEV)), 2, P(wt.detach().to(DEV)), n, h, w, c, stride, dil, P(gout), P(dw),
         P(sums), stream_ptr())
    torch.cuda.synchronize()
    tag = '%s s%d d%d' % ((n, h, w, c), stride, dil)
    ok1, _ = err_stats('fused dw bwd: masked dX ' + tag, gout, gout_ref, 2 * ULP, 2e-4)   # 9-term fp32 sums in another order: 1-ulp flips
    ok2, _ = err_stats('fused dw bwd: dW ' + tag, dw, wt.grad, 1e-4, 1e-5 * float(wt.grad.abs().max()))
    go = gout.float().cpu().double()
    s1, s2 = go.reshape(-1, c).sum(0), (go * zin.double()).reshape(-1, c).sum(0)
    assert ok1 and ok2
    assert torch.allclose(sums[0].cpu(), s1, rtol=1e-5, atol=1e-3)
    assert torch.allclose(sums[1].cpu(), s2, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize('n,h,w,u8', [(2, 64, 128, True), (1, 33, 47, False), (1, 48, 80, True)])
def test_stem(n, h, w, u8):
    L = nat.lib()
    spec = so.load_spec('cityscapes')
    fr = so.synthetic_frames(n, h, w, 3)
    wt = rnd(3, 3, 3, 32, seed=19, scale=0.27).requires_grad_(True)
    x = so.preprocess(spec, fr.astype(np.float32))
    raw = so.conv2d_same(x, wt, 2, 1, False)
    ho, wo = raw.shape[1], raw.shape[2]
    dz = bf16_round(rnd(*raw.shape, seed=20, scale=0.1))
    raw.backward(dz)
    frames = torch.from_numpy(fr).to(DEV) if u8 else torch.from_numpy(fr.astype(np.float32)).to(DEV)
    out = torch.full((n, ho, wo, 32), float('nan'), dtype=AC, device=DEV)
    call(L.ams_op_stem, P(frames), 0 if u8 else 1, n, h, w, P(wt.detach().to(DEV)), None, None, P(out), stream_ptr())
    torch.cuda.synchronize()
    ok1, _ = err_stats('stem raw %s u8=%d' % ((n, h, w), u8), out, raw, AULP, 2e-4)
    scale = torch.rand(32) + 0.5
    shift = rnd(32, seed=21, scale=0.5)
    call(L.ams_op_stem, P(frames), 0 if u8 else 1, n, h, w, P(wt.detach().to(DEV)), P(scale.to(DEV)), P(shift.to(DEV)), P(out), stream_ptr())
    torch.cuda.synchronize()
    ok2, _ = err_stats('stem fold+relu6 %s' % ((n, h, w),), out, (raw.detach() * scale + shift).clamp(0, 6), AULP, 4e-4)
    dw = torch.full((3, 3, 3, 32), float('nan'), dtype=torch.float32, device=DEV)
    call(L.ams_op_stem_bwd, P(frames), 0 if u8 else 1, n, h, w, P(dz.to(DEV, BF)), P(dw), stream_ptr())
    torch.cuda.synchronize()
    ok3, _ = err_stats('stem dW %s' % ((n, h, w),), dw, wt.grad, 1e-4, 1e-5 * float(wt.grad.abs().max()))
    assert ok1 and ok2 and ok3


@pytest.mark.parametrize('M,C,act,res', [(2 * 33 * 65, 64, 2, False), (4 * 129 * 257, 16, 0, True), (2145, 960, 2, False), (1000, 144, 1, False)])
def test_bn_train_and_backward(M, C, act, res):
    L = nat.lib()
    z = ac_round(rnd(M, C, seed=22) * (torch.rand(C) + 0.5) + rnd(C, seed=23)).requires_grad_(True)
    gamma = (torch.rand(C) + 0.5).requires_grad_(True)
    beta = rnd(C, seed=24, scale=0.5).requires_grad_(True)
    r = ac_round(rnd(M, C, seed=25)) if res else None
    eps = 1e-3
    y, mean, _ = so.batch_norm(z.view(1, 1, M, C), gamma, beta, eps, 'batch')
    y = y.view(M, C)
    pre = y
    y = {0: y, 1: y.clamp_min(0), 2: y.clamp(0, 6)}[act]
    yo = y + r if res else y
    dy = bf16_round(rnd(M, C, seed=26, scale=0.01))
    yo.backward(dy)
    out = torch.full((M, C), float('nan'), dtype=AC, device=DEV)
    mean_d = torch.empty(C, device=DEV)
    rstd_d = torch.empty(C, device=DEV)
    zd = z.detach().to(DEV, AC)
    call(L.ams_op_bn_train, P(zd), M, C, P(gamma.detach().to(DEV)), P(beta.detach().to(DEV)), eps, act,
         P(r.to(DEV, AC)) if res else None, P(out), P(mean_d), P(rstd_d), stream_ptr())
    torch.cuda.synchronize()
    ok1, _ = err_stats('bn_train y M%d C%d act%d res%d' % (M, C, act, res), out, yo.detach(), AULP, 4e-4)
    ok2, _ = err_stats('bn_train mean', mean_d, mean.detach(), 1e-5, 1e-5)
    dz = torch.full((M, C), float('nan'), dtype=BF, device=DEV)
    dg = torch.empty(C, device=DEV)
    db = torch.empty(C, device=DEV)
    call(L.ams_op_bn_backward, P(dy.to(DEV, BF)), P(zd), M, C, P(gamma.detach().to(DEV)), P(beta.detach().to(DEV)), eps, act,
         P(dz), P(dg), P(db), stream_ptr())
    torch.cuda.synchronize()
    # pixels whose pre-activation sits within rounding distance of a ReLU knee may legitimately flip
    knee = (pre.detach().abs() < 1e-4) | ((pre.detach() - 6).abs() < 1e-4) if act else torch.zeros_like(pre, dtype=torch.bool)
    ref_dz = z.grad.clone()
    got_dz = dz.float().cpu()
    got_dz[knee] = ref_dz[knee]
    ok3, _ = err_stats('bn_backward dz', got_dz, ref_dz, 2 * ULP, 2e-3 * float(ref_dz.abs().max()))
    # a knee pixel that flips moves its channel's sums by one dy (times |xhat| <= ~5 for dgamma): excuse those channels
    flips = knee.sum(0) > 0
    dg_c, db_c = dg.float().cpu(), db.float().cpu()
    dg_c[flips & ((dg_c - gamma.grad).abs() < 5 * float(dy.abs().max()))] = gamma.grad[flips & ((dg_c - gamma.grad).abs() < 5 * float(dy.abs().max()))]
    db_c[flips & ((db_c - beta.grad).abs() < float(dy.abs().max()))] = beta.grad[flips & ((db_c - beta.grad).abs() < float(dy.abs().max()))]
    ok4, _ = err_stats('bn_backward dgamma', dg_c, gamma.grad, 2e-3, 2e-3 * float(gamma.grad.abs().max()))
    ok5, _ = err_stats('bn_backward dbeta', db_c, beta.grad, 2e-3, 2e-3 * float(beta.grad.abs().max()))
    assert ok1 and ok2 and ok3 and ok4 and ok5


@pytest.mark.parametrize('n,h,w,H,W,cls', [(2, 5, 9, 64, 128, list(range(19))), (1, 33, 65, 512, 1024, [0, 1, 2, 8, 10, 11, 13]),
                                          (1, 4, 6, 48, 80, [3, 7])])
def test_head_infer(n, h, w, H, W, cls):
    L = nat.lib()
    logits = rnd(n, h, w, 19, seed=27, scale=2.0)
    labels = so.synthetic_labels(n, H, W, seed=3, block=8)
    full = so.full_res_logits(logits, H, W)
    ref = so.head(full, labels, np.array(cls))
    lg = torch.zeros(n, h, w, 32)
    lg[..., :19] = logits
    pred = torch.full((n, H, W), -1, dtype=torch.int32, device=DEV)
    cc = len(cls)
    cm = np.zeros((cc, cc), dtype=np.int64)
    ls, nv = C.c_double(), C.c_longlong()
    cls_arr = (C.c_int * cc)(*cls)
    call(L.ams_op_head_infer, P(lg.to(DEV)), n, h, w, 32, H, W, cc, C.cast(cls_arr, C.c_void_p), 19,
         P(torch.from_numpy(labels).to(DEV)), P(pred), cm.ctypes.data_as(C.c_void_p), C.byref(ls), C.byref(nv), stream_ptr())
    torch.cuda.synchronize()
    agree = float((pred.cpu().numpy() == ref['predictions']).mean())
    log('head_infer %s argmax agreement %.6f n_valid %d/%d loss %.6f/%.6f' %
        ((n, h, w, H, W, cc), agree, nv.value, ref['n_valid'], ls.value / max(nv.value, 1), float(ref['loss'])))
    assert agree == 1.0                                   # bit-exact: same fp32 op order as the TF kernel restatement
    assert nv.value == ref['n_valid']
    cm_ref = so.confusion_matrix(ref['labels_reduced'], pred.cpu().numpy(), ref['weights'], cc)
    assert np.array_equal(cm.astype(np.float64), cm_ref)
    assert abs(ls.value / max(nv.value, 1) - float(ref['loss'])) < 1e-4


@pytest.mark.parametrize('n,h,w,H,W,cls', [(2, 5, 9, 64, 128, list(range(19))), (1, 9, 17, 128, 256, [0, 1, 2, 8, 10, 11, 13])])
def test_head_backward(n, h, w, H, W, cls):
    L = nat.lib()
    logits = rnd(n, h, w, 19, seed=28, scale=2.0).requires_grad_(True)
    labels = so.synthetic_labels(n, H, W, seed=4, block=8)
    ref = so.head(so.full_res_logits(logits, H, W), labels, np.array(cls))
    ref['loss'].backward()
    lg = torch.zeros(n, h, w, 32)
    lg[..., :19] = logits.detach()
    dl = torch.full((n, h, w, 32), float('nan'), device=DEV)
    loss = C.c_float()
    cc = len(cls)
    cls_arr = (C.c_int * cc)(*cls)
    call(L.ams_op_head_backward, P(lg.to(DEV)), n, h, w, H, W, cc, C.cast(cls_arr, C.c_void_p), 19,
         P(torch.from_numpy(labels).to(DEV)), P(dl), C.byref(loss), stream_ptr())
    torch.cuda.synchronize()
    ok, _ = err_stats('head_backward dlogits %s' % ((n, h, w, H, W, cc),), dl[..., :19], logits.grad, 1e-3, 1e-4 * float(logits.grad.abs().max()))
    assert ok
    assert float(dl[..., 19:].abs().max()) == 0.0
    assert abs(loss.value - float(ref['loss'])) < 1e-4 * max(1.0, abs(float(ref['loss'])))


@pytest.mark.parametrize('n,frac,ties', [(2113043, 0.05, False), (100003, 0.1, False), (2113043, 0.05, True), (5000, 0.2, True), (4097, 0.01, False)])
def test_select_bit_exact(n, frac, ties):
    L = nat.lib()
    rng = np.random.default_rng(5)
    before = rng.normal(size=n).astype(np.float32)
    if ties:
        # first Adam step from zero state: |delta| ~ lr for every coordinate
        step = (np.float32(1e-3) * np.sign(rng.normal(size=n))).astype(np.float32)
        step[rng.random(n) < 0.3] *= np.float32(0.999)
    else:
        step = (rng.normal(size=n) * 1e-3).astype(np.float32)
    after = (before + step).astype(np.float32)
    mask_ref, comb_ref, thr = so.select_coordinates({'v': before}, {'v': after}, ['v'], frac)
    a = torch.from_numpy(after.copy()).to(DEV)
    b = torch.from_numpy(before).to(DEV)
    m = torch.zeros(n, dtype=torch.uint8, device=DEV)
    kept, thr_d = C.c_longlong(), C.c_float()
    call(L.ams_op_select, P(a), P(b), n, frac, P(m), C.byref(kept), C.byref(thr_d), stream_ptr())
    torch.cuda.synchronize()
    log('select n=%d frac=%g ties=%d thr %.9g/%.9g kept %d/%d' % (n, frac, ties, thr_d.value, float(thr), kept.value, int(mask_ref['v'].sum())))
    assert np.float32(thr_d.value) == thr
    assert np.array_equal(m.cpu().numpy().astype(bool), mask_ref['v'])
    assert kept.value == int(mask_ref['v'].sum())
    assert np.array_equal(a.cpu().numpy(), comb_ref['v'])


def test_adam_bit_exact():
    L = nat.lib()
    spec = {'trainable_variables': [{'name': 'v', 'shape': [100000]}], 'convs': []}
    rng = np.random.default_rng(6)
    st = so.TrainState.__new__(so.TrainState)
    st.spec, st.trainable = spec, ['v']
    st.vars = {'v': rng.normal(size=100000).astype(np.float32)}
    st.m = {'v': np.zeros(100000, np.float32)}
    st.v = {'v': np.zeros(100000, np.float32)}
    st.beta1_power, st.beta2_power = so.BETA1, so.BETA2
    p = torch.from_numpy(st.vars['v'].copy()).to(DEV)
    m = torch.zeros(100000, device=DEV)
    v = torch.zeros(100000, device=DEV)
    mask = (rng.random(100000) < 0.3)
    b1p, b2p = np.float32(0.9), np.float32(0.999)
    for it in range(3):
        g = (rng.normal(size=100000) * 10.0 ** rng.integers(-8, 1, size=100000)).astype(np.float32)
        st.adam_apply({'v': g}, 1e-3, {'v': mask})
        call(L.ams_op_adam, P(p), P(torch.from_numpy(g).to(DEV)), P(m), P(v), P(torch.from_numpy(mask.astype(np.uint8)).to(DEV)),
             100000, 1e-3, float(b1p), float(b2p), stream_ptr())
        b1p, b2p = np.float32(b1p * np.float32(0.9)), np.float32(b2p * np.float32(0.999))
        torch.cuda.synchronize()
        assert np.array_equal(m.cpu().numpy(), st.m['v']), 'm differs at step %d' % it
        assert np.array_equal(v.cpu().numpy(), st.v['v']), 'v differs at step %d' % it
        assert np.array_equal(p.cpu().numpy(), st.vars['v']), 'params differ at step %d' % it


@pytest.mark.parametrize('n,h,w,cin,cout,dil,res,split,stride', [
    (1, 8, 16, 64, 64, 1, True, True, 1),          # exactly one tile
    (2, 33, 65, 64, 64, 1, True, True, 1),         # blocks 7-9: ragged tiles in both directions
    (1, 33, 65, 64, 96, 1, False, True, 1),        # block 10
    (2, 17, 33, 96, 96, 1, True, True, 1),         # blocks 11-12 (3 k-blocks, 4.5 chunks)
    (1, 33, 65, 160, 160, 2, True, True, 1),       # blocks 14-15: dilation 2 (20-wide halo rows, single TMEM stage)
    (1, 33, 65, 96, 160, 1, False, True, 1),       # block 13: single TMEM stage at dilation 1
    (1, 40, 70, 24, 24, 1, True, True, 1),         # block 2: Cin 24 (zero-filled k-block), Cexp 144 (padded chunk), Cout 24 (N granule)
    (1, 30, 50, 32, 32, 1, True, True, 1),         # blocks 4-5
    (2, 65, 129, 16, 24, 1, False, True, 2),       # block 1: stride 2, odd input size (pad 1 / 1), Cexp 96 (one padded chunk)
    (1, 64, 96, 24, 32, 1, False, True, 2),        # block 3: stride 2, even input size (pad 0 / 1), two chunks
    (1, 33, 65, 32, 64, 1, False, True, 2),        # block 6
])
def test_fused_inverted_residual_block(n, h, w, cin, cout, dil, res, split, stride):
    """The block-fused frozen-inference kernel against the same block computed layer by layer with the storage rounding
    of the unfused path (fp16 after the expand conv, after the depthwise conv and at the block output; split fp16 weights
    where the layer has at most 256 output channels)."""
    L = nat.lib()
    cexp = 6 * cin
    x = ac_round(rnd(n, h, w, cin, seed=61))
    we32 = rnd(cexp, cin, seed=62, scale=(2.0 / cin) ** 0.5)
    wp32 = rnd(cout, cexp, seed=63, scale=(2.0 / cexp) ** 0.5)
    wd = rnd(3, 3, cexp, seed=64, scale=0.4)
    g = torch.Generator().manual_seed(65)
    s1, s2, s3 = (torch.rand(c, generator=g) * 0.5 + 0.5 for c in (cexp, cexp, cout))
    t1, t2 = rnd(cexp, seed=66, scale=0.3) + 1.0, rnd(cexp, seed=67, scale=0.3) + 1.0
    t3 = rnd(cout, seed=68, scale=0.2)
    we_hi, wp_hi = ac_round(we32), ac_round(wp32)
    we_split = split and cexp <= 256
    we_lo = ac_round(we32 - we_hi) if we_split else None
    wp_lo = ac_round(wp32 - wp_hi) if split else None
    we_eff = we_hi + (we_lo if we_split else 0.0)
    wp_eff = wp_hi + (wp_lo if split else 0.0)
    y1 = ac_round(((x.reshape(-1, cin).double() @ we_eff.double().t()).float().reshape(n, h, w, cexp) * s1 + t1).clamp(0, 6))
    y2 = ac_round((_dw_ref(y1, wd, stride, dil) * s2 + t2).clamp(0, 6))
    ho, wo = y2.shape[1], y2.shape[2]
    y3 = (y2.reshape(-1, cexp).double() @ wp_eff.double().t()).float().reshape(n, ho, wo, cout) * s3 + t3
    ref = y3 + x if res else y3
    out = torch.full((n, ho, wo, cout), float('nan'), dtype=AC, device=DEV)
    dv = lambda t: P(t.to(DEV)) if t is not None else None
    hv = lambda t: P(t.to(DEV, AC)) if t is not None else None
    call(L.ams_op_fused_block, hv(x), n, h, w, cin, cexp, cout, dil, stride, hv(we_hi), hv(we_lo), dv(s1), dv(t1), dv(wd), dv(s2), dv(t2),
         hv(wp_hi), hv(wp_lo), dv(s3), dv(t3), 1 if res else 0, P(out), stream_ptr())
    torch.cuda.synchronize()
    # an fp16 flip of an intermediate (accumulation order of the tensor core vs the fp64 reference) moves the output by ~1e-3
    ok, _ = err_stats('fused block n%d %dx%d cin%d cout%d d%d s%d res%d' % (n, h, w, cin, cout, dil, stride, res), out, ref, 4 * AULP, 6e-3)
    assert ok
    rel = float((out.float().cpu() - ref).norm() / ref.norm())
    assert rel < 1e-3
