"""Shared helpers for the GPU parity tests (torch owns the device buffers; all compute goes through the C ABI)."""
import ctypes as C
import os

import numpy as np
import torch

from ams_b200 import _native as nat

LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'parity_log.txt')


def log(msg):
    print(msg)
    try:
        os.makedirs(os.path.dirname(LOG), exist_ok=True)
        with open(LOG, 'a') as f:
            f.write(msg + '\n')
    except OSError:
        pass


_KEEPALIVE = []


def P(t):
    """device pointer of a torch tensor (or None).  The tensor is kept alive until release() so that a
    temporary passed inline (`P(x.to('cuda'))`) is not handed back to the caching allocator before the
    asynchronous kernel that reads it has run."""
    if t is None:
        return None
    _KEEPALIVE.append(t)
    return C.c_void_p(t.data_ptr())


def release():
    torch.cuda.synchronize()
    del _KEEPALIVE[:]


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(fn, *args):
    nat.check(fn(*args), fn.__name__)


def bf16_round(t):
    """storage rounding of activation GRADIENTS"""
    return t.to(torch.bfloat16).to(torch.float32)


def ac_round(t):
    """storage rounding of forward ACTIVATIONS and 1x1 weight operands (IEEE fp16, saturating)"""
    return t.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32)


def err_stats(name, got, ref, rtol, atol):
    """max |got-ref| against atol + rtol*|ref|; logs a one-line summary and returns (ok, worst_ratio)."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    diff = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    ratio = float((diff / bound).max()) if diff.numel() else 0.0
    nbad = int((diff > bound).sum())
    rel = float(diff.norm() / (ref.norm() + 1e-30))
    finite = bool(torch.isfinite(got).all())
    log('%-46s max|d| %.3e  rel-L2 %.3e  worst/bound %.2f  bad %d/%d  finite %s' %
        (name, float(diff.max()) if diff.numel() else 0.0, rel, ratio, nbad, diff.numel(), finite))
    return (nbad == 0 and finite), ratio
