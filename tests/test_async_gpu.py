"""Steps without a host round trip (ams_train_step_async, ams_apply_optimizer_device): same results as the synchronous
entry points, losses delivered through page-locked slots."""
import numpy as np
import pytest
import torch

from _util import log
from ams_b200.parallel import DataParallelStudent
from ams_b200.student import Student
from ams_b200.synthetic import synthetic_checkpoint, synthetic_frames, synthetic_labels

pytestmark = pytest.mark.gpu
H, W, B, STEPS = 128, 256, 2, 4
CLASSES = [0, 1, 2, 8, 10, 11, 13]


def _student(ckpt):
    st = Student(19, H, W, CLASSES, device=0, queue_capacity=STEPS + 1)
    for k, v in ckpt.items():
        st.set_tensor(k, v)
    return st


def _feed(st):
    for i in range(STEPS):
        st.enqueue(synthetic_frames(B, H, W, seed=40 + i), synthetic_labels(B, H, W, seed=40 + i, block=16))


def test_async_step_is_the_synchronous_step():
    ckpt = synthetic_checkpoint('cityscapes', 1)
    a, b = _student(ckpt), _student(ckpt)
    _feed(a)
    _feed(b)
    la = [float(a.train_step(1e-3, False)) for _ in range(STEPS)]           # eager, capture, replay, replay
    slots = torch.zeros(STEPS, dtype=torch.float32).pin_memory().numpy()
    for i in range(STEPS):
        b.train_step_async(1e-3, False, slots[i:i + 1])
    b.synchronize()
    lb = [float(x) for x in slots]
    log('async vs sync losses: %s / %s' % (la, lb))
    assert la == lb
    assert np.array_equal(a.get_trainable_flat(), b.get_trainable_flat())
    a.close()
    b.close()


def test_device_terms_optimizer_matches_host_scale():
    """world = 1 exchange step: Adam reads 1 / n_valid from the device terms; against the synchronous hooks with the
    scale computed on the host (same arithmetic: fp64 reciprocal rounded to fp32) the result is bit-identical."""
    ckpt = synthetic_checkpoint('cityscapes', 1)
    a, b = _student(ckpt), _student(ckpt)
    _feed(a)
    _feed(b)
    la = []
    for _ in range(STEPS):
        nv, ls = a.train_forward_backward()
        a.apply_optimizer(1e-3, False, 1.0 / nv)
        la.append(np.float32(ls / nv))
    dp = DataParallelStudent(b)
    for _ in range(STEPS):
        dp.train_step_async(1e-3, False)
    lb = dp.losses()
    log('device-terms losses: %s / %s' % ([float(x) for x in la], lb))
    assert [float(x) for x in la] == lb
    assert np.array_equal(a.get_trainable_flat(), b.get_trainable_flat())
    dp.close()
    a.close()
    b.close()
