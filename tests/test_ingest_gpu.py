"""Frame ingest on the device against the cv2.resize oracle (oracle/cv2_resize_oracle.py, pinned on
OpenCV's own outputs): the resize kernels are bit-exact, and a batch queued through ams_enqueue_raw gives exactly the
predictions / confusion matrix of the same batch resized by the oracle and queued through ams_enqueue."""
import numpy as np
import pytest
import torch

import cv2_resize_oracle as ro
import student_oracle as so
from _util import P, call, log, stream_ptr
from ams_b200 import _native as nat
from ams_b200.student import Student
from ams_b200.synthetic import synthetic_checkpoint

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('n,sh,sw,dh,dw,cn', [(2, 1080, 1920, 512, 1024, 3), (1, 720, 1280, 256, 512, 3), (1, 480, 854, 512, 1024, 3),
                                              (2, 1024, 2048, 512, 1024, 3), (1, 97, 61, 40, 33, 1), (3, 33, 47, 64, 128, 3),
                                              (1, 2160, 3840, 512, 1024, 3)])
def test_resize_kernels_bit_exact(n, sh, sw, dh, dw, cn):
    L = nat.lib()
    rng = np.random.default_rng(sh * 7 + dw)
    src = rng.integers(0, 256, size=(n, sh, sw, cn), dtype=np.uint8)
    lab = rng.integers(0, 21, size=(n, sh, sw), dtype=np.uint8)
    d_src, d_lab = torch.from_numpy(src).to(DEV), torch.from_numpy(lab).to(DEV)
    out = torch.zeros((n, dh, dw, cn), dtype=torch.uint8, device=DEV)
    call(L.ams_op_resize_u8, P(d_src), n, sh, sw, cn, P(out), dh, dw, 0, 0, stream_ptr())
    ref = np.stack([ro.resize_linear_u8(src[i] if cn == 3 else src[i, :, :, 0], dw, dh).reshape(dh, dw, cn) for i in range(n)])
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    nbad = int((got != ref).sum())
    log('resize linear %s -> %s x%d: %d / %d bytes differ' % ((sh, sw), (dh, dw), cn, nbad, ref.size))
    assert nbad == 0
    if cn == 3:
        call(L.ams_op_resize_u8, P(d_src), n, sh, sw, 3, P(out), dh, dw, 0, 1, stream_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref[..., ::-1])                  # BGR -> RGB on the way out
    outl = torch.zeros((n, dh, dw), dtype=torch.uint8, device=DEV)
    call(L.ams_op_resize_u8, P(d_lab), n, sh, sw, 1, P(outl), dh, dw, 1, 0, stream_ptr())
    torch.cuda.synchronize()
    refl = np.stack([ro.resize_nearest_u8(lab[i], dw, dh) for i in range(n)])
    assert np.array_equal(outl.cpu().numpy(), refl)


def test_enqueue_raw_equals_host_resize_then_enqueue():
    H, W, N = 64, 128, 2
    cls = [0, 1, 2, 8, 10, 11, 13]
    st = Student(19, H, W, cls)
    for k, v in synthetic_checkpoint('cityscapes', 1).items():
        st.set_tensor(k, v)
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 256, size=(N, 135, 240, 3), dtype=np.uint8)               # "decoded BGR camera frames"
    raw_lab = so.synthetic_labels(N, 270, 480, seed=4, block=24)                    # teacher maps at another size
    frames = np.stack([ro.ingest_frame(raw[i], H, W) for i in range(N)])
    labels = np.stack([ro.resize_nearest_u8(raw_lab[i], W, H) for i in range(N)])
    st.enqueue(frames, labels)
    p_ref, cm_ref, loss_ref = st.infer_metric(N, nat.BN_MOVING)
    st.enqueue_raw(raw, raw_lab, bgr=True)
    p_dev, cm_dev, loss_dev = st.infer_metric(N, nat.BN_MOVING)
    assert np.array_equal(p_ref, p_dev) and np.array_equal(cm_ref, cm_dev) and np.float32(loss_ref) == np.float32(loss_dev)
    # and a training step sees the same batch
    st.enqueue(frames, labels)
    l1 = st.train_step(0.0, masked=False)
    st.enqueue_raw(raw, raw_lab, bgr=True)
    l2 = st.train_step(0.0, masked=False)
    assert l1 == l2
    st.close()
