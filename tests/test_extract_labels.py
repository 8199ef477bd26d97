"""ams_b200.extract_labels: the frame loop of the reference's extract_labels.py (pad, teacher, crop, PNG dump) with a
plug-in teacher.  CPU part: a deterministic numpy teacher pins the pixel arithmetic and the file names; GPU part: the
MobileNetV2 graph behind the teacher interface at the padded (h+1) x (w+1) size."""
import os

import cv2
import numpy as np
import pytest

from ams_b200 import extract_labels as el
from ams_b200.utils.utils import colormap


class ColourTeacher:
    """label = f(pixel): lets the test predict every output byte, including the effect of the symmetric pad + crop"""

    def __init__(self):
        self.shapes = []

    def predict(self, frame_rgb):
        self.shapes.append(frame_rgb.shape)
        return ((frame_rgb[..., 0].astype(np.int32) // 16 + frame_rgb[..., 2].astype(np.int32) // 64) % 19).astype(np.int32)


def _frames(n, h, w, seed=0):
    return [np.random.default_rng(seed + i).integers(0, 256, size=(h, w, 3), dtype=np.uint8) for i in range(n)]


def test_pad_is_numpy_symmetric_top_left():
    f = _frames(1, 5, 7)[0]
    p = el.pad_top_left_symmetric(f)
    assert p.shape == (6, 8, 3)
    assert np.array_equal(p[1:, 1:], f) and np.array_equal(p[0, 1:], f[0]) and np.array_equal(p[1:, 0], f[:, 0])
    assert np.array_equal(p[0, 0], f[0, 0])


@pytest.mark.parametrize('height', [None, 24])
def test_frame_loop_files_and_bytes(tmp_path, height):
    flags = el.parse_flags(['--dump_path', str(tmp_path / 'labels'), '--teacher_checkpoint', 'unused'] +
                           (['--height', str(height)] if height else []))
    frames = _frames(3, 40, 72, seed=5)
    teacher = ColourTeacher()
    assert el.extract_labels(flags, teacher, frames_bgr=frames, log=lambda *_: None) == 3
    cm = colormap()
    for i, bgr in enumerate(frames):
        rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
        if height:
            rgb = cv2.resize(rgb, (2 * height, height))
        want = ColourTeacher().predict(np.pad(rgb, ((1, 0), (1, 0), (0, 0)), mode='symmetric'))[1:, 1:]
        assert teacher.shapes[i] == (rgb.shape[0] + 1, rgb.shape[1] + 1, 3)
        gt = cv2.imread(os.path.join(flags.dump_path, 'gt_%06d.png' % i), cv2.IMREAD_UNCHANGED)
        assert gt.dtype == np.uint8 and np.array_equal(gt, want.astype(np.uint8))
        annot = cv2.imread(os.path.join(flags.dump_path, 'annot_%06d.png' % i))
        assert np.array_equal(cv2.cvtColor(annot, cv2.COLOR_BGR2RGB), cm[want])
        vis = cv2.imread(os.path.join(flags.dump_path, 'vis_%06d.png' % i))
        assert np.array_equal(cv2.cvtColor(vis, cv2.COLOR_BGR2RGB), cv2.addWeighted(rgb, 0.5, cm[want], 0.5, 0))


def test_flags_are_the_reference_flags():
    f = el.parse_flags([])
    assert vars(f) == dict(dump_path=None, teacher_checkpoint=None, gpu=0, input_video=None, height=None)


@pytest.mark.gpu
def test_mobilenet_graph_behind_the_teacher_interface(tmp_path):
    import torch
    import student_oracle as so
    from _util import log
    spec = so.load_spec('cityscapes')
    h = 64
    frames = _frames(2, 96, 160, seed=11)
    cal = np.stack([np.pad(cv2.resize(cv2.cvtColor(f, cv2.COLOR_BGR2RGB), (2 * h, h)), ((1, 0), (1, 0), (0, 0)), mode='symmetric')
                    for f in frames])
    V = so.calibrate_moving_stats(spec, so.synthetic_variables(spec, 4), cal.astype(np.float32))
    prefix = os.path.join(str(tmp_path), 'teacher')
    np.save(prefix + '.npy', {'teacher/' + k: v for k, v in V.items()})
    flags = el.parse_flags(['--dump_path', str(tmp_path / 'out'), '--teacher_checkpoint', prefix, '--height', str(h)])
    teacher = el.StudentGraphTeacher(flags.teacher_checkpoint, flags.gpu)
    try:
        assert el.extract_labels(flags, teacher, frames_bgr=frames, log=lambda *_: None) == 2
    finally:
        teacher.close()
    params = {k: torch.tensor(v) for k, v in V.items()}
    agree = []
    for i in range(2):
        gt = cv2.imread(os.path.join(flags.dump_path, 'gt_%06d.png' % i), cv2.IMREAD_UNCHANGED)
        assert gt.shape == (h, 2 * h) and gt.max() < 19
        with torch.no_grad():
            sem, _ = so.forward(spec, params, cal[i:i + 1].astype(np.float32), bn_mode='moving', precision='fp32')
            ref = so.full_res_logits(sem, h + 1, 2 * h + 1).argmax(3).numpy()[0][1:, 1:]
        agree.append(float((gt == ref).mean()))
    log('extract_labels through the MobileNetV2 graph at %dx%d (+1 px pad): argmax agreement with the fp32 oracle %.4f'
        % (h, 2 * h, float(np.mean(agree))))
    assert np.mean(agree) >= 0.85                  # random-init stress level; the trained figure is in test_trained_gpu.py
