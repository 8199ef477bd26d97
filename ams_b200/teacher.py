"""`XceptionTeacher`: the DeepLabv3+ / Xception-65 teacher of the label-extraction path (config C5) behind the plug-in
interface of ams_b200/extract_labels.py -- `predict(frame_rgb_uint8[h, w, 3]) -> label ids [h, w]` -- i.e. the network
the reference runs with `sess.run(teacher['predictions'], {teacher['images']: frame})` (extract_labels.py:84) after
importing the teacher's .meta (utils/graph_utils.py:129-152).

Thin typed wrapper over the C ABI (include/ams_b200.h: ams_teacher_*).  The teacher's graph and weights are not part
of the reference repository (README.md:45-46); the topology is the public model-zoo definition restated in
ams_b200/csrc/teacher.cu and variables are addressed by the TF names of that definition, so the checkpoint dict the
reference loads (`np.load('<teacher_checkpoint>.npy').item()`, optionally with the `teacher/` prefix of
extract_labels.py:58) loads here as is.  No CPU fallback."""
import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _native as nat


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def teacher_variables(num_classes=19):
    """[(name, shape)] of the restated graph in creation order -- host only, no device needed."""
    L = nat.lib()
    name = C.create_string_buffer(256)
    shape = (C.c_int * 4)()
    nd = C.c_int()
    out = []
    for i in range(L.ams_teacher_layout_num_tensors(int(num_classes))):
        nat.check(L.ams_teacher_layout_tensor_info(int(num_classes), i, name, 256, shape, C.byref(nd)))
        out.append((name.value.decode(), tuple(shape[:nd.value])))
    return out


def synthetic_teacher_checkpoint(num_classes=19, seed=1):
    """Random-init weights in the teacher's variable layout (BASELINE.json configs[4]: 'random-init weights'): He-normal
    kernels, BatchNorm gamma ~ U(0.5, 1), beta ~ N(0.3, 0.1), moving mean ~ N(0, 0.1), moving variance ~ U(0.5, 1.5)."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in teacher_variables(num_classes):
        if name.endswith('weights:0'):
            fan_in = shape[0] * shape[1] * (1 if 'depthwise' in name else shape[2])
            a = rng.normal(0.0, np.sqrt(2.0 / fan_in), size=shape)
        elif name.endswith('gamma:0'):
            a = rng.uniform(0.5, 1.0, size=shape)
        elif name.endswith('beta:0'):
            a = rng.normal(0.3, 0.1, size=shape)
        elif name.endswith('moving_variance:0'):
            a = rng.uniform(0.5, 1.5, size=shape)
        else:                                   # moving_mean, logits biases
            a = rng.normal(0.0, 0.1, size=shape)
        out[name] = a.astype(np.float32)
    return out


class XceptionTeacher:
    def __init__(self, checkpoint, num_classes=19, gpu=0):
        """checkpoint: dict {'<tf variable name>:0': ndarray} or the path prefix of '<prefix>.npy' holding such a dict."""
        self._h = None
        L = nat.lib()
        if isinstance(checkpoint, str):
            path = checkpoint if checkpoint.endswith('.npy') else checkpoint + '.npy'
            checkpoint = np.load(path, allow_pickle=True).item()
        h = L.ams_teacher_create(int(num_classes), int(gpu))
        if not h:
            raise nat.NativeError('ams_teacher_create failed: ' + nat.last_error())
        self._h, self._L, self.num_classes = C.c_void_p(h), L, int(num_classes)
        want = dict(teacher_variables(num_classes))
        seen = set()
        for key, value in checkpoint.items():
            name = key[len('teacher/'):] if key.startswith('teacher/') else key
            if name in ('global_step:0',) or 'Momentum' in name or 'Adam' in name:       # extract_labels.py:59-60 filter
                continue
            a = np.ascontiguousarray(value, dtype=np.float32)
            nat.check(L.ams_teacher_set_tensor(self._h, name.encode(), _ptr(a), a.size), 'teacher set_tensor(%s)' % name)
            seen.add(name)
        missing = [n for n in want if n not in seen]
        if missing:
            self.close()
            raise KeyError('teacher checkpoint lacks %d variables, e.g. %s' % (len(missing), missing[0]))

    def predict_batch(self, frames_rgb, want_logits=False):
        """frames [n,h,w,3] uint8 -> int32 label ids [n,h,w] (predictions:0) (+ fp32 logits at output stride 4)."""
        f = np.ascontiguousarray(frames_rgb, dtype=np.uint8)
        assert f.ndim == 4 and f.shape[3] == 3, f.shape
        n, h, w = f.shape[:3]
        out = np.empty((n, h, w), dtype=np.int32)
        logits = np.empty((n, -(-h // 4), -(-w // 4), self.num_classes), dtype=np.float32) if want_logits else None
        nat.check(self._L.ams_teacher_predict(self._h, _ptr(f), n, h, w, _ptr(out), _ptr(logits)), 'teacher predict')
        return (out, logits) if want_logits else out

    def predict(self, frame_rgb):
        """teacher plug-in interface of ams_b200.extract_labels: one (already padded) RGB frame -> label ids [h, w]."""
        return self.predict_batch(np.asarray(frame_rgb)[None])[0]

    def time_forward(self, reps=5):
        ms = C.c_float()
        nat.check(self._L.ams_teacher_time_forward(self._h, int(reps), C.byref(ms)), 'teacher time_forward')
        return float(ms.value)

    def close(self):
        if self._h is not None:
            self._L.ams_teacher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
