"""Teacher network on the GPU box (config C5): DeepLabv3+ / Xception-65 through the C ABI (ams_teacher_*) against the
oracle's torch restatement of the same public definition on the same seeded weights.  (Both are restatements -- the
teacher's graph and weights are not in the reference repository: parity unpinned AND unsourced)
  * logits at output stride 4 within the fp16-storage tolerance of the oracle's own fp16 mode, labels agree;
  * against the fp32 arithmetic: relative L2 and label agreement at the calibrated level of a RANDOM-INIT 65-layer network;
  * the extract_labels.py frame loop (pad, predict, crop, PNG dump) over the real teacher;
  * BASELINE size: 1025 x 2049 (1024 x 2048 + the reference's 1-px pad): deterministic, timed, TFLOP/s logged."""
import os

import numpy as np
import pytest
import torch

import teacher_oracle as to
from _util import log
from ams_b200 import extract_labels as el
from ams_b200.teacher import XceptionTeacher, synthetic_teacher_checkpoint

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,h,w', [(1, 65, 97), (2, 49, 81), (1, 64, 100)])
def test_teacher_logits_and_labels_against_oracle(n, h, w):
    torch.set_num_threads(os.cpu_count())
    V = to.synthetic_checkpoint(19, seed=5)
    fr = np.random.default_rng(3).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    t = XceptionTeacher(V, 19)
    pred, logits = t.predict_batch(fr, want_logits=True)
    pred2 = t.predict_batch(fr)
    t.close()
    assert np.array_equal(pred, pred2)                                           # deterministic
    with torch.no_grad():
        l16, p16 = to.forward(V, fr, precision='fp16')
        l32, p32 = to.forward(V, fr, precision='fp32')
    ld = torch.from_numpy(logits)
    assert ld.shape == l16.shape
    rel16 = float((ld - l16).norm() / l16.norm())
    rel32 = float((ld - l32).norm() / l32.norm())
    agree16 = float((torch.from_numpy(pred) == p16).float().mean())
    agree32 = float((torch.from_numpy(pred) == p32).float().mean())
    # argmax of the device's own logits, upsampled by the oracle: the integer part of the pipeline is exact
    own = to.so.resize_bilinear_align(ld, h, w).argmax(dim=3).numpy()
    log('teacher %dx%dx%d: logits rel-L2 vs fp16-storage oracle %.2e, vs fp32 oracle %.2e | labels agree %.4f / %.4f | |logit| max %.2f'
        % (n, h, w, rel16, rel32, agree16, agree32, float(l32.abs().max())))
    assert np.array_equal(pred, own)
    assert rel16 < 5e-3 and agree16 > 0.99
    assert rel32 < 2e-2 and agree32 > 0.97


def test_extract_labels_loop_over_the_xception_teacher(tmp_path):
    cv2 = pytest.importorskip('cv2')
    V = synthetic_teacher_checkpoint(19, seed=5)
    t = XceptionTeacher(V, 19)
    flags = el.default_flags()
    flags.dump_path = str(tmp_path) + os.sep
    flags.height = 64
    rng = np.random.default_rng(1)
    frames = [rng.integers(0, 256, size=(90, 160, 3), dtype=np.uint8) for _ in range(2)]       # BGR camera frames
    n = el.extract_labels(flags, t, frames_bgr=frames, log=lambda *a: None)
    assert n == 2
    gt = cv2.imread(os.path.join(str(tmp_path), 'gt_000001.png'), cv2.IMREAD_UNCHANGED)
    assert gt.shape == (64, 128) and gt.dtype == np.uint8 and gt.max() < 19
    # the same frame by hand: BGR->RGB, resize, 1-px symmetric pad, teacher, crop
    f = cv2.resize(cv2.cvtColor(frames[1], cv2.COLOR_BGR2RGB), (128, 64))
    want = t.predict(el.pad_top_left_symmetric(f))[1:, 1:]
    t.close()
    assert np.array_equal(gt, want.astype(np.uint8))


def test_teacher_at_baseline_size_is_deterministic_and_timed():
    V = synthetic_teacher_checkpoint(19, seed=1)
    t = XceptionTeacher(V, 19)
    fr = np.random.default_rng(0).integers(0, 256, size=(1, 1025, 2049, 3), dtype=np.uint8)
    a = t.predict_batch(fr)
    b = t.predict_batch(fr)
    ms = t.time_forward(5)
    t.close()
    assert a.shape == (1, 1025, 2049) and np.array_equal(a, b) and a.max() < 19 and len(np.unique(a)) > 1
    log('teacher DeepLabv3+/Xception-65 @ 1025x2049 (1024x2048 + 1-px pad), batch 1: %.2f ms per frame, %.1f frames/s on one B200'
        % (ms, 1000.0 / ms))
