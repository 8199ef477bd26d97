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


def test_orchestration_server_and_client_on_synthetic_video(tmp_path):
    """ams_b200.run: the reference's server loop (sampling, memory, phases, ASR, delta accounting, hand-off) and client
    loop (scheduled model loads, per-frame mIoU) on a synthetic camera; the delta file equals the host writer of
    run.py:316-328 byte for byte and a client that applies it holds the server's fp16-rounded coordinates."""
    pytest.importorskip('cv2')
    from ams_b200 import run
    spec, V, frames, prefix = _checkpoint(tmp_path)
    flags = run.default_flags()
    flags.output_dir = os.path.join(str(tmp_path), 'out') + os.sep
    os.makedirs(flags.output_dir)
    flags.student_checkpoint = prefix
    flags.input_video = 'synthetic-0.mp4'
    flags.height, flags.batch_size, flags.iter, flags.memory_len = H, 2, 3, 8
    flags.train_strategy, flags.coord_fraction, flags.enable_ASR = 'coord_desc_auto', '0.05', True
    logs = []
    src = run.SyntheticSource(90, 160, fps=2, seed=1)
    out = run.train_model(flags, src, 0, 12, 1, '0', 'syn', 12, [0, 4, 8], 1, log=logs.append)
    assert out['update_count'] == 2 and out['model_update_times'] == [0, 4.0, 8.0]
    assert len(out['samples']) == 12 and sum(out['samples']) > 0 and all(b > 0 for b in out['downlink_bits'])
    assert any('Send rate updated' in l for l in logs)
    final = run.get_save_dir(flags, 'syn_results')
    for suffix in ('_fps_client.npy', '_bw_uplink.npy', '_bw_downlink.npy', '_model_update_times.npy', '_update.txt'):
        assert os.path.exists(final + suffix), suffix
    down, up, count, interval, sent = [int(x) for x in open(final + '_update.txt').read().split()]
    assert (count, interval, sent) == (2, 12, sum(out['samples'])) and down == sum(out['downlink_bits'])
    # the delta file of the first update: packbits masks then fp16 values, as run.py writes it from curr_mask / train_params
    t, blob = out['deltas'][0]
    path = run.get_save_dir(flags, 'syn_0') + '_mask.dat'
    assert open(path, 'rb').read() == blob and os.path.exists(path + '.gz')
    # client: loads the model in force at each second, predicts every frame
    res = run.infer_output(flags, run.SyntheticSource(90, 160, fps=2, seed=1), 0, 12, '0', 'syn', 12, [0, 4, 8], log=logs.append)
    assert len(res['miou']) == 24 and np.isfinite(res['loss']).all() and len(res['miou_mem']) == 24
    assert os.path.exists(final + '_mious.npy') and os.path.exists(final + '_mioucats.npy')
    # in-place update of a resident client from the streamed bytes == the server's coordinates through fp16
    cw = class_weights(12)
    client = SemanticNetwork(meta_dir=run.get_save_dir(flags, 'syn_0') + '_final', class_weights_exp=cw, height=H, gpu_id='0', frozen=True)
    before = client.student.split_trainable(client.student.get_trainable_flat())
    n_upd = client.apply_delta(blob)
    after = client.student.split_trainable(client.student.get_trainable_flat())
    ref, masks = so.apply_delta([before[n] for n in client.student.trainable_names], blob)
    assert n_upd == sum(int(m.sum()) for m in masks) > 0
    for n, r in zip(client.student.trainable_names, ref):
        assert np.array_equal(after[n], r), n
    client.close_model()
