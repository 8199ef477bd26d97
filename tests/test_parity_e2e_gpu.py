"""END-TO-END, UN-FORCED parity of the CUDA path against the reference's fp32 arithmetic (the fp32 oracle), at
BASELINE.json's sizes: batch 8 @ 512x1024, Cityscapes 19-class graph (config C2) and PASCAL VOC 21-class graph (C4).

The stated tolerance (north_star: "within a stated fp tolerance (e.g. max-abs 1e-2 ..., argmax agreement >= 99.9 %)"),
written here as the assertions:
  * frozen (client) inference: low-res logits relative L2 <= 6e-4 and max-abs <= 1e-3 x max|logit| (a trained
    student's logits reach |27|: measured max-abs 1.8e-2 .. 2.0e-2, i.e. 7e-4 of the range, asserted <= 2.5e-2 absolute;
    the example figure 1e-2 is met wherever |logit| <= 12 -- DESIGN.md 3 states this deviation and what bounds it),
    per-pixel argmax agreement >= 99.9 % (measured 99.98 %), confusion matrix / mIoU computed by the device equal the
    oracle's metric on the device's predictions bit for bit, mIoU delta vs the fp32 oracle's own predictions <= 1e-3
    (measured 1e-5 .. 3e-5);
  * one distillation step (forward with batch statistics, backward, BN moving-average update), nothing teacher-forced:
    loss within 1e-3 of the fp32 oracle's (measured 1e-6 at full size), every gradient tensor's cosine with the oracle's
    autograd gradient >= 0.99 (measured worst 0.9995 at full size, 0.997 at 2x128x256), moving statistics within 3e-4
    relative (measured 3e-5 at full size).
The checkpoint is the device-distilled student of tests/_trained.py (a trained network, like the missing shipped
weights -- not the chaotic random-init stack).  What makes these bounds reachable is the storage precision:
forward activations in IEEE fp16 (2^-11 relative rounding, 8x finer than bf16), 1x1 weights as split fp16 pairs
(hi + lo, ~21 bits) wherever the conv has <= 256 output channels, fp32 accumulation / statistics / logits, bf16 only
for the activation gradients.
The full-size cases need ~25 GB of host memory and ~1 minute of host time for the oracle's autograd; on a smaller host
they drop to batch 2 and say so in the log."""
import numpy as np
import psutil
import pytest
import torch

import student_oracle as so
from _trained import CONFIGS, scenes, trained_variables
from _util import log
from ams_b200 import _native as nat
from ams_b200.student import Student

pytestmark = pytest.mark.gpu

LOGIT_MAX_ABS = 2.5e-2          # absolute, at |logit| <= 27 (north_star's example figure 1e-2 assumes |logit| ~ 5-10)
LOGIT_MAX_REL = 1e-3            # max-abs / max|logit|
LOGIT_REL_L2 = 6e-4
ARGMAX_AGREE = 0.999            # north_star: "per-pixel argmax agreement >= 99.9 %"
LOSS_ABS = 1e-3
GRAD_COS = 0.99
MOVING_REL = 3e-4

CASES = [('cityscapes', 2, 128, 256), ('cityscapes', 8, 512, 1024), ('pascalvoc2012', 8, 512, 1024)]
IDS = ['small-cityscapes-b2-128x256', 'full-cityscapes-b8-512x1024', 'full-voc-b8-512x1024']


def _batch(n, h, w):
    if h * w >= 512 * 1024 and n > 2 and psutil.virtual_memory().available < n * 3.2e9:
        log('host has %.0f GB available: full-size parity case runs at batch 2 instead of %d'
            % (psutil.virtual_memory().available / 1e9, n))
        return 2
    return n


def _cos(a, b):
    a, b = a.reshape(-1).astype(np.float64), b.reshape(-1).astype(np.float64)
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))


@pytest.mark.parametrize('tag,n,h,w', CASES, ids=IDS)
def test_frozen_logits_and_argmax_against_fp32_oracle(tag, n, h, w):
    torch.set_num_threads(psutil.cpu_count(logical=True))
    V, _, digest = trained_variables(tag)
    cfg = CONFIGS[tag]
    spec = so.load_spec(tag)
    cls = np.array(cfg['classes'])
    fr, lab = scenes(tag, n, 7000, h, w)
    st = Student(cfg['num_classes'], h, w, cfg['classes'])
    for k, v in V.items():
        st.set_tensor(k, v)
    st.enqueue(fr, lab)
    pred, cm, loss = st.infer_metric(n, nat.BN_MOVING)
    logits = torch.from_numpy(st.get_logits(n))
    st.close()
    with torch.no_grad():
        sem, _ = so.forward(spec, {k: torch.tensor(v) for k, v in V.items()}, fr.astype(np.float32), bn_mode='moving', precision='fp32')
        ref = so.head(so.full_res_logits(sem, h, w), lab, cls)
    max_abs = float((logits - sem).abs().max())
    rel = float((logits - sem).norm() / sem.norm())
    agree = float((pred == ref['predictions']).mean())
    fl, wts = so.reduce_labels(lab, cls)
    assert np.array_equal(cm.astype(np.float64), so.confusion_matrix(fl, pred, wts, len(cls)))      # integer pipeline exact
    miou_dev = float(np.nanmean(so.calculate_miou(cm.astype(np.float64))))
    miou_ref = float(np.nanmean(so.calculate_miou(ref['conf_mat'])))
    log('[e2e frozen %s b%d %dx%d ckpt %s] logits max-abs %.3e (|logit| max %.2f) rel-L2 %.3e | argmax agreement %.6f | loss %.6f vs %.6f | '
        'mIoU %.6f vs %.6f (delta %+.2e)' % (tag, n, h, w, digest, max_abs, float(sem.abs().max()), rel, agree, float(loss),
                                             float(ref['loss']), miou_dev, miou_ref, miou_dev - miou_ref))
    assert miou_ref > 0.5, 'the parity checkpoint must be a student that learned its scenes'
    assert max_abs <= LOGIT_MAX_ABS and max_abs <= LOGIT_MAX_REL * float(sem.abs().max()) and rel <= LOGIT_REL_L2
    assert agree >= ARGMAX_AGREE
    assert abs(float(loss) - float(ref['loss'])) <= LOSS_ABS
    assert abs(miou_dev - miou_ref) <= 1e-3


@pytest.mark.parametrize('tag,n,h,w', CASES, ids=IDS)
def test_distillation_step_free_running_against_fp32_oracle(tag, n, h, w):
    torch.set_num_threads(psutil.cpu_count(logical=True))
    n = _batch(n, h, w)
    V, _, digest = trained_variables(tag)
    cfg = CONFIGS[tag]
    spec = so.load_spec(tag)
    cls = np.array(cfg['classes'])
    fr, lab = scenes(tag, n, 8000, h, w)
    st = Student(cfg['num_classes'], h, w, cfg['classes'])
    for k, v in V.items():
        st.set_tensor(k, v)
    st.enqueue(fr, lab)
    loss = st.train_step(1e-3, masked=False)
    g_dev = st.split_trainable(st.get_gradients())
    moving_dev = {nm: st.get_tensor(nm) for nm, _, tr, _ in st.variables if not tr}
    st.close()
    ts = so.TrainState(spec, V, precision='fp32')
    loss_ref, g_ref, stats, _ = ts.loss_and_grads(fr.astype(np.float32), lab, cls)
    ts.apply_moving_stats(stats)
    gmax = max(float(np.linalg.norm(v)) for v in g_ref.values())
    worst_cos, worst_rel, worst_name, n_checked = 1.0, 0.0, '', 0
    tot_a = np.concatenate([g_dev[k].reshape(-1) for k in ts.trainable])
    tot_b = np.concatenate([g_ref[k].reshape(-1) for k in ts.trainable])
    for name in ts.trainable:
        a, b = g_dev[name], g_ref[name]
        nb = float(np.linalg.norm(b))
        if nb < 1e-5 * gmax:
            # analytically-zero gradients (a conv bias / BN shift feeding a batch-statistics BN): rounding noise on both sides
            assert float(np.linalg.norm(a)) < 1e-3 * gmax, name
            continue
        cs, rel = _cos(a, b), float(np.linalg.norm(a - b) / nb)
        n_checked += 1
        if cs < worst_cos:
            worst_cos, worst_name = cs, name
        worst_rel = max(worst_rel, rel)
    worst_mv = 0.0
    for nm, got in moving_dev.items():
        ref = ts.vars[nm]
        worst_mv = max(worst_mv, float(np.abs(got - ref).max() / (np.abs(ref).max() + 1e-12)))
    log('[e2e train %s b%d %dx%d ckpt %s] loss %.6f vs fp32 oracle %.6f | %d gradient tensors: worst cosine %.5f (%s), worst rel-L2 %.4f, '
        'whole-arena cosine %.6f | moving statistics worst rel %.2e'
        % (tag, n, h, w, digest, loss, loss_ref, n_checked, worst_cos, worst_name, worst_rel, _cos(tot_a, tot_b), worst_mv))
    assert abs(loss - loss_ref) <= LOSS_ABS
    assert worst_cos >= GRAD_COS
    assert worst_mv <= MOVING_REL
