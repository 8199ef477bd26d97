"""C-ABI boundary behaviours the reference's callers rely on, straight through ctypes:
  * frozen hand-off: ams_export_frozen / ams_create_frozen (reference save_to_frozen_graph + the frozen branch of
    SemanticNetwork.__init__, SemanticNetwork.py:706-714, :80-118) -- the client handle predicts bit-identically to the
    server handle's own inference-mode run, refuses to train, and rejects a foreign file;
  * threading: a feeder thread inside ams_enqueue WHILE the trainer thread runs ams_train_step (reference `_fill_queue`
    thread vs `sess.run(train)`, SemanticNetwork.py:230-231, :701 vs :260) gives results bit-identical to the serial
    order enqueue-all-then-train;
  * a corrupt delta leaves the handle untouched; ams_queue_clear drops staged batches."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

import student_oracle as so
from _util import log
from ams_b200 import _native as nat
from ams_b200.student import Student

pytestmark = pytest.mark.gpu
H, W, N = 64, 128, 2
CLS = [0, 1, 2, 8, 10, 11, 13]


def _student(**kw):
    spec = so.load_spec('cityscapes')
    fr = so.synthetic_frames(N, H, W, seed=0)
    V = so.calibrate_moving_stats(spec, so.synthetic_variables(spec, 1), fr.astype(np.float32))
    st = Student(19, H, W, CLS, **kw)
    for k, v in V.items():
        st.set_tensor(k, v)
    return st, V, fr


def test_export_frozen_create_frozen_roundtrip(tmp_path):
    st, V, fr = _student()
    lab = so.synthetic_labels(N, H, W, seed=2, block=16)
    st.enqueue(fr, lab)
    st.train_step(1e-3, False)                       # move weights and moving statistics away from the checkpoint
    st.enqueue(fr, lab)
    pred_s, cm_s, loss_s = st.infer_metric(N, nat.BN_MOVING)
    logits_s = st.get_logits(N)
    path = os.path.join(str(tmp_path), 'client_final.pb')
    st.export_frozen(path)
    assert open(path, 'rb').read(8) == b'AMSFRZ01'
    client = Student(None, H, W, CLS, frozen_path=path)
    assert client.frozen and client.num_classes == 19 and nat.lib().ams_is_frozen(client._h) == 1
    for name, _, _, _ in st.variables:               # all 272 variables, moving statistics included, bit for bit
        assert np.array_equal(client.get_tensor(name), st.get_tensor(name)), name
    client.enqueue(fr, lab)
    pred_c, cm_c, loss_c = client.infer_metric(N, nat.BN_MOVING)
    assert np.array_equal(pred_c, pred_s) and np.array_equal(cm_c, cm_s) and np.float32(loss_c) == np.float32(loss_s)
    assert np.array_equal(client.get_logits(N), logits_s)
    client.enqueue(fr, lab)
    with pytest.raises(nat.NativeError, match="Can't train frozen graph"):
        client.train_step(1e-3, False)
    client.queue_clear()
    client.close()
    bad = os.path.join(str(tmp_path), 'tf_graph.pb')
    open(bad, 'wb').write(b'\x0a\x03abc' * 10)
    with pytest.raises(ValueError):
        Student(None, H, W, CLS, frozen_path=bad)
    with pytest.raises(nat.NativeError):             # exported from the 19-class graph, asked for as the 21-class one
        Student(21, H, W, [0, 7], frozen_path=path)
    st.close()
    log('frozen hand-off through the C ABI: client == server (272 variables, predictions, confusion matrix, logits)')


def test_feeder_thread_concurrent_with_train_step_is_bit_identical_to_serial_order():
    K = 6
    batches = [(so.synthetic_frames(N, H, W, seed=10 + i), so.synthetic_labels(N, H, W, seed=20 + i, block=16)) for i in range(K)]

    def run(threaded):
        st, _, _ = _student(queue_capacity=2)        # capacity < K: the feeder really blocks on the trainer
        losses = []
        if threaded:
            err = []

            def feed():
                try:
                    for f, l in batches:
                        st.enqueue(f, l)
                except BaseException as e:           # noqa: BLE001
                    err.append(e)
            th = threading.Thread(target=feed)
            th.start()
            for _ in range(K):
                losses.append(st.train_step(1e-3, False))
            th.join()
            assert not err, err
        else:
            for f, l in batches:
                st.enqueue(f, l)
                losses.append(st.train_step(1e-3, False))
        flat = st.get_trainable_flat()
        mv = np.concatenate([st.get_tensor(n).reshape(-1) for n, _, tr, _ in st.variables if not tr])
        st.close()
        return np.array(losses, dtype=np.float32), flat, mv

    l1, p1, m1 = run(False)
    l2, p2, m2 = run(True)
    assert np.array_equal(l1, l2) and np.array_equal(p1, p2) and np.array_equal(m1, m2)
    log('feeder thread in ams_enqueue concurrent with ams_train_step: %d steps bit-identical to the serial order' % K)


def test_corrupt_delta_leaves_the_handle_untouched_and_queue_clear():
    st, _, fr = _student()
    lab = so.synthetic_labels(N, H, W, seed=2, block=16)
    st.snapshot_before()
    st.enqueue(fr, lab)
    st.train_step(1e-3, True)
    st.select_topk(0.05)
    blob = st.pack_delta()
    mask0, params0 = st.get_mask().copy(), st.get_trainable_flat().copy()
    other, _, _ = _student()
    m_before, p_before = other.get_mask().copy(), other.get_trainable_flat().copy()
    with pytest.raises(nat.NativeError):
        other.apply_delta(blob[:-2])                 # truncated: rejected ...
    assert np.array_equal(other.get_mask(), m_before) and np.array_equal(other.get_trainable_flat(), p_before)   # ... and nothing changed
    assert other.apply_delta(blob) == int(mask0.sum())
    assert np.array_equal(other.get_mask(), mask0)
    sel = mask0.astype(bool)
    assert np.array_equal(other.get_trainable_flat()[sel], params0[sel].astype(np.float16).astype(np.float32))
    other.enqueue(fr, lab)
    other.enqueue(fr, lab)
    assert other.queue_size() == 2 and other.queue_clear() == 2 and other.queue_size() == 0
    other.close()
    st.close()
