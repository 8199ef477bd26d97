"""Teacher label extraction: the orchestration of the reference's extract_labels.py.

For every frame of a video the reference (extract_labels.py:62-106) converts BGR->RGB, optionally resizes to
(2*height, height), pads ONE pixel at the top and at the left in 'symmetric' mode, runs the teacher, crops the
prediction back with [1:, 1:] and writes three PNGs per frame into `dump_path`: `gt_%06d.png` (uint8 label ids -- the
files `run.py` later reads as teacher labels, run.py:104-111), `annot_%06d.png` (colourised labels) and
`vis_%06d.png` (frame and colourised labels blended 50/50).  This module keeps that loop, the flags, the file names and
the pixel arithmetic, with two fix-forwards: the reference's `teacher_conf` line (extract_labels.py:86) reads a variable
that is never assigned and would raise NameError on the first frame -- dropped; frames come from any iterable, not only
cv2.VideoCapture.

The teacher is a plug-in: any object with `predict(frame_rgb_uint8[h, w, 3]) -> label ids [h, w]`.
  * `ams_b200.teacher.XceptionTeacher`: DeepLabv3+ / Xception-65 (config C5) on the device through the C ABI
    (ams_teacher_*, ams_b200/csrc/teacher.cu).  The teacher's graph (`<teacher_checkpoint>.meta`) and weights are not
    part of the reference repository (external download, README.md:45-46; `create_teacher` only names three of its
    tensors, utils/graph_utils.py:148-152), so the topology restates the public model-zoo definition and is checked
    against an independent torch restatement (oracle/teacher_oracle.py): unpinned and unsourced, and said so.
  * `StudentGraphTeacher`: the DeeplabV3-MobileNetV2 graph of this library (frozen BatchNorm, all 19 / 21 classes) behind the
    same interface, for checkpoints in the student layout.
`main()` picks the backend from the variable names of the checkpoint.
"""
import argparse
import os
import time

import numpy as np

from .exp_configs import test_length
from .utils.utils import colormap


def default_flags():
    """The reference's flag set (extract_labels.py:19-26)."""
    return argparse.Namespace(dump_path=None, teacher_checkpoint=None, gpu=0, input_video=None, height=None)


def parse_flags(argv=None):
    d = default_flags()
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--dump_path', type=str, default=d.dump_path, help='Directory of the path data')
    ap.add_argument('--teacher_checkpoint', type=str, default=d.teacher_checkpoint, help='Directory for teacher checkpoint')
    ap.add_argument('--gpu', type=int, default=d.gpu, help='GPU to use for this')
    ap.add_argument('--input_video', type=str, default=d.input_video, help='Video used in the test, optional')
    ap.add_argument('--height', type=int, default=d.height, help='height to extract labels')
    return ap.parse_args(argv)


def pad_top_left_symmetric(frame):
    """np.pad(frame, ((1, 0), (1, 0), (0, 0)), mode='symmetric') -- extract_labels.py:83: the first row / column repeated
    once in front (the teacher graph is fed (h+1) x (w+1) and its prediction is cropped back with [1:, 1:], :85)."""
    return np.pad(frame, ((1, 0), (1, 0), (0, 0)), mode='symmetric')


class StudentGraphTeacher:
    """DeeplabV3-MobileNetV2 (this library's graph, frozen BatchNorm, every class of the graph) behind the teacher
    interface.  Built lazily for the first frame's size: the network sees the padded (h+1) x (w+1) frame."""

    def __init__(self, checkpoint_prefix, gpu=0):
        path = checkpoint_prefix if checkpoint_prefix.endswith('.npy') else checkpoint_prefix + '.npy'
        ckpt = np.load(path, allow_pickle=True)
        self.variables = ckpt.item() if ckpt.dtype == object else dict(ckpt)
        # the reference prefixes the teacher's variables with 'teacher/' (extract_labels.py:58): accept both spellings
        self.variables = {(k[len('teacher/'):] if k.startswith('teacher/') else k): v for k, v in self.variables.items()}
        self.num_classes = int(np.asarray(self.variables['logits/semantic/biases:0']).size)
        self.gpu = int(gpu)
        self._student = None
        self._shape = None

    def _build(self, h, w):
        from .student import Student
        if self._student is not None:
            self._student.close()
        st = Student(self.num_classes, h, w, list(range(self.num_classes)), device=self.gpu, label_depth=max(19, self.num_classes))
        for name, value in self.variables.items():
            if name in ('global_step:0',) or 'Momentum' in name or 'Adam' in name:     # extract_labels.py:59-60 filter
                continue
            st.set_tensor(name, value)
        self._student, self._shape = st, (h, w)

    def predict(self, frame_rgb):
        from . import _native as nat
        frame_rgb = np.ascontiguousarray(frame_rgb, dtype=np.uint8)
        h, w = frame_rgb.shape[:2]
        if self._shape != (h, w):
            self._build(h, w)
        self._student.enqueue(frame_rgb[None], None)
        return self._student.infer(1, nat.BN_MOVING)[0]

    def close(self):
        if self._student is not None:
            self._student.close()
            self._student = None


def video_frames(path, max_frames=None):
    """BGR frames of a video file, at most max_frames (reference: test_length(exp_num) * fps, extract_labels.py:63-65)."""
    import cv2
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise IOError('Error opening video stream or file: %s' % path)
    try:
        n = 0
        while max_frames is None or n < max_frames:
            ret, frame = cap.read()
            if not ret:
                break
            yield frame
            n += 1
    finally:
        cap.release()


def video_frame_budget(path):
    """exp_num = leading integer of the file name; frames = test_length(exp_num) * round(fps) (extract_labels.py:45, :64-65)."""
    import cv2
    exp_num = int(os.path.basename(path).split('-')[0])
    cap = cv2.VideoCapture(path)
    fps = round(cap.get(cv2.CAP_PROP_FPS))
    cap.release()
    return int(test_length(exp_num) * fps)


def extract_labels(flags, teacher, frames_bgr=None, log=print):
    """The frame loop of extract_labels.py:62-106.  `frames_bgr`: iterable of BGR uint8 frames (default: flags.input_video).
    Returns the number of frames written."""
    import cv2
    os.makedirs(flags.dump_path, exist_ok=True)
    dump = flags.dump_path if flags.dump_path.endswith(os.sep) else flags.dump_path + os.sep
    colormap_ = colormap()
    max_length = None
    if frames_bgr is None:
        max_length = video_frame_budget(flags.input_video)
        frames_bgr = video_frames(flags.input_video, max_length)
        log('There are %d frames to extract' % max_length)
    index_frame = 0
    begin_time = time.time()
    for frame in frames_bgr:
        frame = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
        if flags.height is not None:
            frame = cv2.resize(frame, (flags.height * 2, flags.height))
        correct_shape = np.shape(frame)[:-1]
        padded = pad_top_left_symmetric(frame)
        teacher_out = np.asarray(teacher.predict(padded))[1:, 1:]
        assert np.shape(teacher_out) == correct_shape, (np.shape(teacher_out), correct_shape)
        cv2.imwrite('%sgt_%06d.png' % (dump, index_frame), np.array(teacher_out, dtype=np.uint8))
        label_colored = colormap_[teacher_out]
        cv2.imwrite('%sannot_%06d.png' % (dump, index_frame), cv2.cvtColor(np.array(label_colored, dtype=np.uint8), cv2.COLOR_RGB2BGR))
        colored_frame = cv2.addWeighted(np.array(frame, dtype=np.uint8), 0.5, np.array(label_colored, dtype=np.uint8), 0.5, 0)
        cv2.imwrite('%svis_%06d.png' % (dump, index_frame), cv2.cvtColor(np.array(colored_frame, dtype=np.uint8), cv2.COLOR_RGB2BGR))
        index_frame += 1
        if index_frame % 100 == 0 and max_length:
            t = (time.time() - begin_time) / index_frame * (max_length - index_frame)
            log('Have computed %d frames so far, ETF: %02d:%02d.%02d' % (index_frame, t // 60, t % 60, (t * 100) % 100))
    return index_frame


def make_teacher(checkpoint_prefix, gpu=0):
    """Backend by checkpoint layout: 'xception_65/...' variables -> XceptionTeacher, 'MobilenetV2/...' -> StudentGraphTeacher."""
    path = checkpoint_prefix if checkpoint_prefix.endswith('.npy') else checkpoint_prefix + '.npy'
    ckpt = np.load(path, allow_pickle=True)
    variables = ckpt.item() if ckpt.dtype == object else dict(ckpt)
    if any('xception_65/' in k for k in variables):
        from .teacher import XceptionTeacher
        nc = int(np.asarray(variables[[k for k in variables if k.endswith('logits/semantic/biases:0')][0]]).size)
        return XceptionTeacher(variables, nc, gpu)
    return StudentGraphTeacher(checkpoint_prefix, gpu)


def main(argv=None):
    flags = parse_flags(argv)
    print('Extracting labels...')
    teacher = make_teacher(flags.teacher_checkpoint, flags.gpu)
    try:
        print('Starting Teacher Inference')
        n = extract_labels(flags, teacher)
        print('Wrote %d label maps to %s' % (n, flags.dump_path))
    finally:
        teacher.close()


if __name__ == '__main__':
    main()
