"""`Student`: the device-resident DeeplabV3-MobileNetV2 student behind `SemanticNetwork`.

Thin, typed wrapper over the C ABI (include/ams_b200.h): owns one `ams_net` handle (one GPU), converts numpy
buffers to the pointer/size arguments the library takes, and keeps the reference's variable-name table
(`utils/utils.py:10-49` SaveHelper semantics: names are TF variable names with ':0').
"""
import ctypes as C
import json
import os

import numpy as np

from . import _native as nat

_GRAPHS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'graphs')
GRAPH_TAGS = {19: 'cityscapes', 21: 'pascalvoc2012'}


def load_graph_spec(tag):
    with open(os.path.join(_GRAPHS, tag + '.json')) as f:
        return json.load(f)


def _low_res(size):
    """Spatial size of the logits: 1-px pad, then four stride-2 'SAME' convolutions (stem, blocks 1, 3, 6)."""
    v = size + 1
    for _ in range(4):
        v = -(-v // 2)
    return v


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Student:
    def __init__(self, num_classes, height, width, class_indices, device=0, label_depth=19, queue_capacity=4,
                 frozen_path=None):
        """frozen_path: build the inference-only client handle from a model written by export_frozen()
        (ams_create_frozen; num_classes may be None = take it from the file)."""
        self._h = None
        L = nat.lib()
        cfg = nat.AmsConfig()
        cfg.num_classes = int(num_classes) if num_classes is not None else 0
        cfg.graph_variant = 1 if cfg.num_classes == 21 else 0
        cfg.height, cfg.width, cfg.device = int(height), int(width), int(device)
        class_indices = [int(c) for c in class_indices]
        cfg.class_count = len(class_indices)
        for i, c in enumerate(class_indices):
            cfg.class_indices[i] = c
        cfg.label_depth = int(label_depth)
        cfg.queue_capacity = int(queue_capacity)
        if frozen_path is not None:
            h = L.ams_create_frozen(os.fsencode(frozen_path), C.byref(cfg))
            if not h:
                msg = nat.last_error()
                if 'not an ams_b200 frozen model' in msg or 'cannot open frozen model' in msg:
                    raise ValueError(msg)
                raise nat.NativeError('ams_create_frozen failed: ' + msg)
        else:
            h = L.ams_create(C.byref(cfg))
            if not h:
                raise nat.NativeError('ams_create failed: ' + nat.last_error())
        self._h = C.c_void_p(h)
        self._L = L
        self.frozen = frozen_path is not None
        self.height, self.width = int(height), int(width)
        self.class_indices = class_indices
        self.class_count = len(class_indices)
        self.variables = []          # [(name, shape, trainable, offset)] in tf.global_variables() order
        name = C.create_string_buffer(256)
        shape = (C.c_int * 4)()
        nd, tr, off = C.c_int(), C.c_int(), C.c_longlong()
        for i in range(L.ams_num_tensors(self._h)):
            nat.check(L.ams_tensor_info(self._h, i, name, 256, shape, C.byref(nd), C.byref(tr), C.byref(off)))
            self.variables.append((name.value.decode(), tuple(shape[:nd.value]), bool(tr.value), off.value))
        self.var_shapes = {n: s for n, s, _, _ in self.variables}
        self.num_classes = int(self.var_shapes['logits/semantic/biases:0'][0])
        self.trainable_names = [n for n, _, t, _ in self.variables if t]
        self.n_trainable = int(L.ams_trainable_count(self._h))
        self.low_res = (_low_res(height), _low_res(width))

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if self._h is not None:
            self._L.ams_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        nat.check(self._L.ams_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        nat.check(self._L.ams_synchronize(self._h))

    # ------------------------------------------------------------------ variables
    def set_tensor(self, name, value):
        a = np.ascontiguousarray(value, dtype=np.float32)
        nat.check(self._L.ams_set_tensor(self._h, name.encode(), _ptr(a), a.size), 'set_tensor(%s)' % name)

    def get_tensor(self, name):
        base = name
        for suf in ('/Adam_1:0', '/Adam:0'):
            if name.endswith(suf):
                base = name[:-len(suf)] + ':0'
        shape = () if name in ('beta1_power:0', 'beta2_power:0') else self.var_shapes.get(base)
        if shape is None:
            raise KeyError(name)
        a = np.empty(shape, dtype=np.float32)
        nat.check(self._L.ams_get_tensor(self._h, name.encode(), _ptr(a), a.size), 'get_tensor(%s)' % name)
        return a

    def get_trainable_flat(self):
        a = np.empty(self.n_trainable, dtype=np.float32)
        nat.check(self._L.ams_get_trainable(self._h, _ptr(a)))
        return a

    def set_trainable_flat(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        assert a.size == self.n_trainable
        nat.check(self._L.ams_set_trainable(self._h, _ptr(a)))

    def split_trainable(self, flat):
        """flat arena -> {name: array} in tf.trainable_variables() order."""
        out = {}
        for n, s, t, off in self.variables:
            if t:
                cnt = int(np.prod(s))
                out[n] = flat[off:off + cnt].reshape(s)
        return out

    def reset_optimizer(self):
        nat.check(self._L.ams_reset_optimizer(self._h))

    # ------------------------------------------------------------------ input queue
    def enqueue(self, frames, labels=None):
        frames = np.asarray(frames)
        assert frames.ndim == 4 and frames.shape[1:] == (self.height, self.width, 3), frames.shape
        if frames.dtype == np.uint8:
            f, dt = np.ascontiguousarray(frames), nat.FRAMES_U8
        else:
            f, dt = np.ascontiguousarray(frames, dtype=np.float32), nat.FRAMES_F32
        lab = None
        if labels is not None:
            labels = np.asarray(labels)
            assert labels.shape == frames.shape[:3], labels.shape
            if labels.dtype != np.uint8:
                # reference: tf.cast(labels, int32) truncates; ids outside [0,255) can never be a selected class
                li = np.trunc(labels).astype(np.int64)
                labels = np.where((li >= 0) & (li < 255), li, 255).astype(np.uint8)
            lab = np.ascontiguousarray(labels)
        nat.check(self._L.ams_enqueue(self._h, _ptr(f), dt, _ptr(lab), frames.shape[0]), 'enqueue')
        return frames.shape[0]

    def enqueue_raw(self, frames, labels=None, bgr=True):
        """Camera-size uint8 frames [n,h,w,3] (BGR as decoded by cv2 when bgr=True) and optional teacher label maps
        [n,lh,lw]: resized on the device exactly as the reference does on the host -- cv2.resize(frame, (W, H)) +
        cv2.cvtColor(BGR2RGB), labels with INTER_NEAREST (run.py:181-183, :415-421) -- then queued like enqueue()."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        assert frames.ndim == 4 and frames.shape[3] == 3, frames.shape
        lab, lh, lw = None, 0, 0
        if labels is not None:
            lab = np.ascontiguousarray(labels, dtype=np.uint8)
            assert lab.ndim == 3 and lab.shape[0] == frames.shape[0], lab.shape
            lh, lw = lab.shape[1:]
        nat.check(self._L.ams_enqueue_raw(self._h, _ptr(frames), frames.shape[1], frames.shape[2], 1 if bgr else 0, _ptr(lab),
                                          lh, lw, frames.shape[0]), 'enqueue_raw')
        return frames.shape[0]

    def export_frozen(self, path):
        """ams_export_frozen: the client model (reference save_to_frozen_graph, SemanticNetwork.py:711-714)."""
        nat.check(self._L.ams_export_frozen(self._h, os.fsencode(path)), 'export_frozen')

    def set_block_fusion(self, on):
        """frozen inference: one kernel per stride-1 inverted-residual block (default) or one kernel per layer"""
        nat.check(self._L.ams_set_block_fusion(self._h, 1 if on else 0), 'set_block_fusion')

    def set_infer_split(self, on):
        """frozen inference of an even batch >= 4 as two concurrent half batches (default on; bit-identical results)"""
        nat.check(self._L.ams_set_infer_split(self._h, 1 if on else 0), 'set_infer_split')

    def queue_size(self):
        return self._L.ams_queue_size(self._h)

    def queue_clear(self):
        return self._L.ams_queue_clear(self._h)

    # ------------------------------------------------------------------ inference
    def _label_buffer(self, n):
        """Fresh int32 [n,H,W] array in page-locked host memory (torch's caching host allocator recycles the blocks):
        the 2 MB/frame label map then leaves the device at PCIe speed instead of through a pageable staging copy."""
        import torch
        return torch.empty((n, self.height, self.width), dtype=torch.int32, pin_memory=True).numpy()

    def infer(self, n, bn_mode):
        out = self._label_buffer(n)
        nat.check(self._L.ams_infer(self._h, bn_mode, _ptr(out)), 'infer')
        return out

    def infer_metric(self, n, bn_mode):
        out = self._label_buffer(n)
        cm = np.zeros((self.class_count, self.class_count), dtype=np.int64)
        loss = C.c_float()
        nat.check(self._L.ams_infer_metric(self._h, bn_mode, _ptr(out), _ptr(cm), C.byref(loss)), 'infer_metric')
        return out, cm, np.float32(loss.value)

    def confmat_labels(self, before, after):
        b = np.ascontiguousarray(before, dtype=np.uint8)
        a = np.ascontiguousarray(after, dtype=np.uint8)
        assert a.shape == b.shape
        cm = np.zeros((self.class_count, self.class_count), dtype=np.int64)
        nat.check(self._L.ams_confmat_labels(self._h, _ptr(b), _ptr(a), a.size, _ptr(cm)), 'confmat_labels')
        return cm

    # ------------------------------------------------------------------ training
    def train_step(self, lr, masked):
        loss = C.c_float()
        nat.check(self._L.ams_train_step(self._h, float(lr), 1 if masked else 0, C.byref(loss)), 'train_step')
        return np.float32(loss.value)

    def train_step_async(self, lr, masked, loss_out=None):
        """The same step without the host round trip: everything is enqueued on the stream; `loss_out` (a 1-element
        float32 array in page-locked memory, e.g. `torch.empty(1, pin_memory=True).numpy()`) receives the pre-update
        loss when the stream gets there.  Call synchronize() before reading it."""
        nat.check(self._L.ams_train_step_async(self._h, float(lr), 1 if masked else 0, _ptr(loss_out)), 'train_step_async')

    def set_mask(self, mask_flat):
        if mask_flat is None:
            nat.check(self._L.ams_set_mask(self._h, None))
        else:
            m = np.ascontiguousarray(mask_flat, dtype=np.uint8)
            assert m.size == self.n_trainable
            nat.check(self._L.ams_set_mask(self._h, _ptr(m)))

    def get_mask(self):
        m = np.empty(self.n_trainable, dtype=np.uint8)
        nat.check(self._L.ams_get_mask(self._h, _ptr(m)))
        return m

    def snapshot_before(self):
        nat.check(self._L.ams_snapshot_before(self._h))

    def select_topk(self, coord_frac):
        kept, thr = C.c_longlong(), C.c_float()
        nat.check(self._L.ams_select_topk(self._h, float(coord_frac), C.byref(kept), C.byref(thr)), 'select_topk')
        return kept.value, np.float32(thr.value)

    def pack_delta(self):
        n = C.c_longlong()
        nat.check(self._L.ams_pack_delta(self._h, None, 0, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        nat.check(self._L.ams_pack_delta(self._h, _ptr(buf), buf.size, C.byref(n)))
        return buf.tobytes()

    # ------------------------------------------------------------------ data-parallel hooks
    def apply_delta(self, blob):
        """Client side: apply a delta produced by pack_delta() to the resident parameters; returns the number of updated
        coordinates."""
        buf = np.frombuffer(bytes(blob), dtype=np.uint8)
        n = C.c_longlong()
        nat.check(self._L.ams_apply_delta(self._h, _ptr(buf), len(buf), C.byref(n)), 'apply_delta')
        return int(n.value)

    def train_forward_backward(self):
        nv, ls = C.c_longlong(), C.c_double()
        nat.check(self._L.ams_train_forward_backward(self._h, C.byref(nv), C.byref(ls)), 'train_forward_backward')
        return nv.value, ls.value

    def train_forward_backward_async(self):
        """Forward/backward enqueued without synchronising; (n_valid, loss_sum) stay on the device (step_terms_ptr)."""
        nat.check(self._L.ams_train_forward_backward(self._h, None, None), 'train_forward_backward')

    def step_terms_ptr(self):
        """Device pointer to the two doubles (n_valid, loss_sum) of the last forward/backward."""
        return self._L.ams_step_terms_device(self._h)

    def apply_optimizer_device(self, lr, masked, loss_out=None):
        """Adam + mask with the gradient scale 1 / n_valid read from the (allreduced) device terms; `loss_out` as in
        train_step_async (global mean loss)."""
        nat.check(self._L.ams_apply_optimizer_device(self._h, float(lr), 1 if masked else 0, _ptr(loss_out)), 'apply_optimizer_device')

    # global-batch BatchNorm statistics over NVLink peer memory (include/ams_b200.h: ams_syncbn_*)
    def syncbn_init(self, world, rank):
        buf = np.zeros(64, dtype=np.uint8)
        nat.check(self._L.ams_syncbn_init(self._h, int(world), int(rank), _ptr(buf), buf.size), 'syncbn_init')
        return buf

    def syncbn_connect(self, handles):
        h = np.ascontiguousarray(handles, dtype=np.uint8)
        assert h.ndim == 2 and h.shape[1] == 64, h.shape
        nat.check(self._L.ams_syncbn_connect(self._h, _ptr(h), h.shape[0]), 'syncbn_connect')

    def syncbn_enable(self, on):
        nat.check(self._L.ams_syncbn_enable(self._h, 1 if on else 0), 'syncbn_enable')

    def syncbn_status(self):
        ep, er = C.c_uint(), C.c_uint()
        nat.check(self._L.ams_syncbn_status(self._h, C.byref(ep), C.byref(er)), 'syncbn_status')
        return ep.value, er.value

    def gradient_arena(self):
        n = C.c_longlong()
        p = self._L.ams_gradient_arena(self._h, C.byref(n))
        return p, n.value

    def gradient_bucket_split(self):
        """first float of the late-layer gradient bucket (final-resolution stage + ASPP + logits)"""
        return int(self._L.ams_gradient_bucket_split(self._h))

    def gradient_bucket_wait(self, cuda_stream_ptr):
        """make `cuda_stream_ptr` wait until the late bucket of the last enqueued forward/backward is complete"""
        nat.check(self._L.ams_gradient_bucket_wait(self._h, C.c_void_p(cuda_stream_ptr)), 'gradient_bucket_wait')

    def apply_optimizer(self, lr, masked, grad_scale):
        nat.check(self._L.ams_apply_optimizer(self._h, float(lr), 1 if masked else 0, float(grad_scale)))

    # ------------------------------------------------------------------ profiling
    def profile_enable(self, on=True):
        nat.check(self._L.ams_profile_enable(self._h, int(on)))

    def profile_report(self):
        """{tag: dict(launches, ms, algo_bytes)} accumulated since profile_enable(True)."""
        buf = C.create_string_buffer(1 << 19)
        self._L.ams_profile_report(self._h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            tag, n, ms, b = line.split()
            out[tag] = dict(launches=int(n), ms=float(ms), algo_bytes=float(b))
        return out

    # ------------------------------------------------------------------ parity hooks
    def get_logits(self, n):
        h, w = self.low_res
        a = np.empty((n, h, w, self.num_classes), dtype=np.float32)
        nat.check(self._L.ams_get_logits(self._h, _ptr(a), a.size), 'get_logits')
        return a

    def get_gradients(self):
        a = np.empty(self.n_trainable, dtype=np.float32)
        nat.check(self._L.ams_get_gradients(self._h, _ptr(a)))
        return a

    def layers(self):
        out = []
        name = C.create_string_buffer(256)
        iv = [C.c_int() for _ in range(7)]
        fv = [C.c_float(), C.c_float()]
        for i in range(self._L.ams_num_layers(self._h)):
            nat.check(self._L.ams_layer_info(self._h, i, name, 256, *[C.byref(v) for v in iv[:6]],
                                             C.byref(fv[0]), C.byref(fv[1]), C.byref(iv[6])))
            out.append(dict(name=name.value.decode(), kind=iv[0].value, cin=iv[1].value, cout=iv[2].value,
                            stride=iv[3].value, dilation=iv[4].value, act=iv[5].value, eps=fv[0].value,
                            one_minus_decay=fv[1].value, residual=iv[6].value))
        return out

    def get_activation(self, index, shape, which=0):
        a = np.empty(shape, dtype=np.uint16)
        nat.check(self._L.ams_get_activation(self._h, index, which, _ptr(a), a.size), 'get_activation')
        # forward tensors are IEEE fp16, gradients (which == 2) bf16
        return a.view(np.float16).astype(np.float32) if which != 2 else (a.astype(np.uint32) << 16).view(np.float32)
