"""The drop-in facade on the GPU box: `ams_b200.SemanticNetwork` driven the way run.py drives the reference class
(build from <meta_dir>.npy, predict, predict_with_metric, train_with_deque with coord_desc_auto, delta, frozen
export -> frozen client), checked against the oracle where the result is an exact function of device state."""
import os
from collections import deque

import numpy as np
import pytest
import torch

import student_oracle as so
from _util import log
from ams_b200.SemanticNetwork import SemanticNetwork
from ams_b200.exp_configs import class_weights

pytestmark = pytest.mark.gpu
H = 64


def _checkpoint(tmp_path):
    spec = so.load_spec('cityscapes')
    frames = so.synthetic_frames(6, H, 2 * H, seed=3)
    V = so.calibrate_moving_stats(spec, so.synthetic_variables(spec, 2), frames[:2].astype(np.float32))
    prefix = os.path.join(str(tmp_path), 'model')
    np.save(prefix + '.npy', dict(V))
    return spec, V, frames, prefix


def test_facade_end_to_end(tmp_path):
    spec, V, frames, prefix = _checkpoint(tmp_path)
    labels = so.synthetic_labels(6, H, 2 * H, seed=3, block=16)
    cw = class_weights(12)
    net = SemanticNetwork(meta_dir=prefix, class_weights_exp=cw, height=H, gpu_id='0', scale=[1], mini_batch_size=2,
                          lr=1e-3, coord_frac=0.05, masked_gradients=True, cross_miou_compat=True)
    cls = net.class_indices_graph
    assert list(cls) == [0, 1, 2, 8, 10, 11, 13] and net.class_count == 7
    # --- predict_input / predict_with_metric (trainable graph => batch statistics)
    pred = net.predict_input(frames[:1])
    assert pred.shape == (1, H, 2 * H) and pred.dtype == np.int32 and pred.max() < 7
    p2, cm, iou, miou, loss = net.predict_with_metric(frames[:1], labels[:1])
    assert np.array_equal(pred, p2) and cm.shape == (7, 7) and cm.dtype == np.float64
    fl, w = so.reduce_labels(labels[:1], cls)
    assert np.array_equal(cm, so.confusion_matrix(fl, p2, w, 7))
    assert np.isclose(miou, np.nanmean(so.calculate_miou(cm))) and np.isfinite(loss)
    # --- calc_cross_miou (ASR phi-score)
    cm2, _, m2 = net.calc_cross_miou(np.stack([labels[0], labels[1]]).astype(np.int64))
    fa, wa = so.reduce_labels(labels[1], cls)
    fb, wb = so.reduce_labels(labels[0], cls)
    assert np.array_equal(cm2, so.confusion_matrix(fb, fa, wa * wb, 7))
    # --- restore surface
    with pytest.raises(KeyError):
        net.restore({'no/such/variable:0': np.zeros(3, np.float32)})
    net.restore({'MobilenetV2/Conv/weights/Adam:0': np.zeros(1)})          # optimizer slots are filtered out, not an error
    got = net.get_vars()
    assert np.array_equal(got['aspp0/weights:0'], V['aspp0/weights:0']) and 'aspp0/weights/Adam:0' in got
    # --- one training phase, coord_desc_auto
    with pytest.raises(NameError):
        net.get_train_mask('no_such_strategy')
    before = {k: got[k] for k in net.student.trainable_names}
    fr_deque = deque(frames[i] for i in range(6))
    lab_deque = deque(labels[i] for i in range(6))
    net.train_with_deque(fr_deque, lab_deque, 3, 'coord_desc_auto')
    assert len(net.curr_mask) == len(net.train_params) == 164
    kept = sum(int(m.sum()) for m in net.curr_mask)
    total = sum(m.size for m in net.curr_mask)
    log('facade: coord_desc_auto kept %d of %d coordinates (%.3f %%)' % (kept, total, 100.0 * kept / total))
    assert 0 < kept <= int(0.0502 * total)
    for name, m, p in zip(net.student.trainable_names, net.curr_mask, net.train_params):
        assert m.shape == p.shape == before[name].shape
        assert np.array_equal(p[~m], before[name][~m]), name            # unselected coordinates never moved
    blob = net.delta_bytes()
    assert blob == so.pack_delta(net.curr_mask, net.train_params)          # run.py:316-328 wire format
    # a second phase keeps Adam state, recomputes the mask (keep_mask=False)
    net.restore_initial()
    net.train_with_deque(fr_deque, lab_deque, 2, 'coord_desc_auto')
    # --- hard-coded strategy + random strategy
    np.random.seed(0)
    _, mk = net.get_train_mask('coord_desc_last')                          # coord_frac 0.05 table
    all_vars, n_train = net.train_vars_count(mk)
    assert abs(n_train / all_vars - 0.05) < 0.002
    net.coord_frac = 0.1
    _, mk = net.get_train_mask('coord_desc_rand')
    assert abs(net.train_vars_count(mk)[1] / all_vars - 0.1) < 0.002
    net.coord_frac = 0.05
    # --- full_model strategy path
    net2 = SemanticNetwork(meta_dir=prefix, class_weights_exp=cw, height=H, gpu_id='0', scale=[1], mini_batch_size=2,
                           lr=1e-3, masked_gradients=False)
    net2.train_with_deque(fr_deque, lab_deque, 2, 'full_model')
    assert all(m.all() for m in net2.curr_mask) and len(net2.train_params) >= 272
    net2.close_model()
    # --- frozen hand-off: server exports, client loads and runs on moving statistics
    out = os.path.join(str(tmp_path), 'run_10_final')
    net.save_to_frozen_graph(out)
    client = SemanticNetwork(meta_dir=out, class_weights_exp=cw, height=H, gpu_id='0', frozen=True)
    pc, cmc, _, mc, lc = client.predict_with_metric(frames[:1], labels[:1])
    logits = torch.from_numpy(client.student.get_logits(1))
    ref = so.head(so.full_res_logits(logits, H, 2 * H), labels[:1], cls)
    assert np.array_equal(pc, ref['predictions']) and np.array_equal(cmc, ref['conf_mat'])
    with pytest.raises(AssertionError):
        client.train_with_deque(fr_deque, lab_deque, 1)
    with pytest.raises(ValueError):
        open(out + '_tf.pb', 'wb').write(b'\x0a\x03abc')
        SemanticNetwork(meta_dir=out + '_tf', class_weights_exp=cw, height=H, frozen=True)
    client.close_model()
    net.close_model()
