"""Drop-in `SemanticNetwork` for the AMS student hot path, backed by libams_b200 (hand-written sm_100a kernels).

Same constructor arguments, method names, argument order, return tuples, public attributes (`curr_mask`,
`train_params`, `mask`, `class_indices_graph`, `take_array` ...) and assertion / NameError behaviour as the
reference class (`/root/reference/SemanticNetwork.py:24-755`); every `sess.run` of the reference maps to one
C-ABI call (see include/ams_b200.h).  What changes underneath:
  * the graph is not imported from `<meta_dir>.meta`; the library carries the same topology (checked against the
    shipped model.meta by tests/test_layout.py) and picks the 19- or 21-class variant from the checkpoint;
  * masks live on the device as one byte map in tf.trainable_variables() order and are uploaded when they
    change, not re-fed every iteration (reference: SemanticNetwork.py:255-257);
  * the coord_desc_auto selection (|after-before| percentile, SemanticNetwork.py:263-288) runs on the device;
  * the frozen client model `<dir>.pb` is this package's own container (np.savez of the 272 variables), not a
    TF GraphDef -- see INTEGRATION.md.
There is no CPU fallback: without the CUDA library / an sm_100a device construction fails.
"""
import threading
import time
from collections import OrderedDict, deque

import numpy as np

from . import _native as nat
from .student import Student
from .utils.utils import SaveHelper, calculate_miou, colormap, mini_batch

try:
    from termcolor import colored
except ImportError:                                     # termcolor is not part of this image
    def colored(text, *_a, **_k):
        return text


_FIRST = ['/Conv/', '/expanded_conv/'] + ['/expanded_conv_%d/' % i for i in range(1, 17)]


def _E(i, part):
    return ['MobilenetV2/expanded_conv_%d/%s/weights:0' % (i, part),
            'MobilenetV2/expanded_conv_%d/%s/BatchNorm/gamma:0' % (i, part),
            'MobilenetV2/expanded_conv_%d/%s/BatchNorm/beta:0' % (i, part)]


# (strategy, coord_frac) -> (substrings selecting whole tensors, exact names selecting whole tensors,
#                            {boundary tensor: Bernoulli P(True)})      -- reference SemanticNetwork.py:310-653
_CP_BN = ['concat_projection/BatchNorm/gamma:0', 'concat_projection/BatchNorm/beta:0']
_MASK_TABLES = {
    ('coord_desc_last', 0.1): ([], ['aspp0/BatchNorm/gamma:0', 'aspp0/BatchNorm/beta:0', 'concat_projection/weights:0'] + _CP_BN +
                               ['logits/semantic/weights:0', 'logits/semantic/biases:0'], {'aspp0/weights:0': 0.90728}),
    ('coord_desc_first', 0.1): (_FIRST[:10], _E(9, 'expand'), {'MobilenetV2/expanded_conv_9/depthwise/depthwise_weights:0': 0.25231}),
    ('coord_desc_both', 0.1): (_FIRST[:8] + ['logits/semantic/'],
                               _E(7, 'expand') + ['MobilenetV2/expanded_conv_7/depthwise/depthwise_weights:0'] + _CP_BN,
                               {'MobilenetV2/expanded_conv_7/depthwise/BatchNorm/gamma:0': 0.80208, 'concat_projection/weights:0': 0.76490}),
    ('coord_desc_last', 0.05): (['logits/semantic/'], _CP_BN, {'concat_projection/weights:0': 0.76490}),
    ('coord_desc_first', 0.05): (_FIRST[:8], _E(7, 'expand') + ['MobilenetV2/expanded_conv_7/depthwise/depthwise_weights:0'],
                                 {'MobilenetV2/expanded_conv_7/depthwise/BatchNorm/gamma:0': 0.80208}),
    ('coord_desc_both', 0.05): (_FIRST[:6] + ['/expanded_conv_5/expand/', '/expanded_conv_5/depthwise/', 'logits/semantic/'], _CP_BN,
                                {'MobilenetV2/expanded_conv_5/project/weights:0': 0.42285, 'concat_projection/weights:0': 0.36187}),
    ('coord_desc_last', 0.01): (['logits/semantic/', 'concat_projection/BatchNorm/'], [], {'concat_projection/weights:0': 0.12005}),
    ('coord_desc_first', 0.01): (_FIRST[:4] + ['/expanded_conv_3/depthwise/', '/expanded_conv_3/expand/'], [],
                                 {'MobilenetV2/expanded_conv_3/project/weights:0': 0.00217}),
    ('coord_desc_both', 0.01): (_FIRST[:3] + ['logits/semantic/', 'concat_projection/BatchNorm/'],
                                ['MobilenetV2/expanded_conv_2/expand/weights:0', 'MobilenetV2/expanded_conv_2/expand/BatchNorm/gamma:0'],
                                {'MobilenetV2/expanded_conv_2/expand/BatchNorm/beta:0': 0.03472, 'concat_projection/weights:0': 0.03944}),
    ('coord_desc_last', 0.2): (['logits/semantic/', 'concat_projection/', 'aspp0/', 'image_pooling/',
                                'MobilenetV2/expanded_conv_16/project/BatchNorm'], [],
                               {'MobilenetV2/expanded_conv_16/project/weights:0': 0.39270}),
    ('coord_desc_first', 0.2): (_FIRST[:12] + ['/expanded_conv_11/expand/', '/expanded_conv_11/depthwise/'], [],
                                {'MobilenetV2/expanded_conv_11/project/weights:0': 0.97367}),
    ('coord_desc_both', 0.2): (_FIRST[:10] + ['concat_projection/', 'aspp0/BatchNorm/', 'logits/semantic/'], _E(9, 'expand'),
                               {'MobilenetV2/expanded_conv_9/depthwise/depthwise_weights:0': 0.25231, 'aspp0/weights:0': 0.90728}),
    ('coord_desc_last', 0.02): (['logits/semantic/', 'concat_projection/BatchNorm/'], [], {'concat_projection/weights:0': 0.7187}),
    ('coord_desc_first', 0.02): (_FIRST[:6], [], {'MobilenetV2/expanded_conv_5/expand/weights:0': 0.7367}),
    ('coord_desc_both', 0.02): (_FIRST[:4] + ['/expanded_conv_3/depthwise/', '/expanded_conv_3/expand/', 'logits/semantic/',
                                              'concat_projection/BatchNorm/'], [],
                                {'MobilenetV2/expanded_conv_3/project/weights:0': 0.00217, 'concat_projection/weights:0': 0.12005}),
}


class SemanticNetwork(object):
    OPT_FILTER = ['Adam', 'Momentum']
    OP_FILTER = ['image_cache:0', 'global_step:0']
    THREAD_SLEEP_INTERVAL = 1 / 1000.
    TOTAL_CLASSES = 19
    WHITE = np.array([255, 255, 255], dtype=np.uint8)
    BLACK = np.array([0, 0, 0], dtype=np.uint8)

    def __init__(self, meta_dir, class_weights_exp=None, height=None, gpu_id='0', frozen=False,
                 scale=None, mini_batch_size=None, lr=None, mem_frac=1, coord_frac=0.1, cross_miou_compat=False,
                 filter_out=None, over_ride_total_classes=None, **kwargs):
        assert height is not None, "No height is given"
        assert class_weights_exp is not None, "No class weights specified"
        assert frozen or None not in [scale, mini_batch_size, lr], "Training parameters must be specified for " \
                                                                   "non-frozen graph"
        self.lr = lr
        self.mini_batch_size = mini_batch_size
        self.scale = scale
        if over_ride_total_classes is not None:
            print(colored('Overriding default number of classes', 'cyan'))
            self.TOTAL_CLASSES = over_ride_total_classes
        self.coord_frac = coord_frac
        self.class_weights_graph = np.asarray(class_weights_exp)
        self.class_indices_graph = np.where(self.class_weights_graph == 1)[0]
        assert self.class_weights_graph.shape == (self.TOTAL_CLASSES, 1)
        self.class_count = len(self.class_indices_graph)
        assert self.class_indices_graph.shape == (self.class_count,)
        assert self.class_count > 0
        self.cross_miou_compat = cross_miou_compat
        self.color_map_reduced_ = np.take(colormap(), self.class_indices_graph, axis=0)
        self.take_array = np.cumsum(self.class_weights_graph).reshape(
            self.TOTAL_CLASSES) * self.class_weights_graph.reshape(self.TOTAL_CLASSES)
        self.take_array = np.where(self.take_array != 0, self.take_array - 1, self.take_array)
        self.take_array = self.take_array.astype(int)
        assert self.take_array.shape == (self.TOTAL_CLASSES,)
        self.frozen = frozen
        self.height = height
        assert self.height > 0
        self.meta_dir = meta_dir
        self.process_lock = threading.Lock()
        for flag in ('train_biases_only', 'regularize', 'soft_teacher'):
            if kwargs.get(flag):
                raise NotImplementedError('%s=True is outside the hot path (never set by run.py)' % flag)
        self.masked_gradients = bool(kwargs.get('masked_gradients', False))
        device = int(str(gpu_id).split(',')[0]) if str(gpu_id) != '' else 0

        # one_hot depth for teacher labels is the reference's NUM_CLASSES constant (utils/graph_utils.py:15)
        if self.frozen:
            # reference :80-93: import `<meta_dir>.pb`; here the container ams_export_frozen wrote (C ABI: ams_create_frozen)
            checkpoint = None
            self.student = Student(None, self.height, 2 * self.height, self.class_indices_graph, device=device, label_depth=19,
                                   frozen_path=meta_dir + ".pb")
        else:
            checkpoint = np.load("%s.npy" % meta_dir, allow_pickle=True).item()
            num_classes = int(np.asarray(checkpoint['logits/semantic/biases:0']).shape[0])
            self.student = Student(num_classes, self.height, 2 * self.height, self.class_indices_graph, device=device,
                                   label_depth=19)
        self.bn_mode = nat.BN_MOVING if self.frozen else nat.BN_BATCH
        self.saver = SaveHelper(self.student, map_fun=lambda x: x)
        self.save_vars = [n for n, _, _, _ in self.student.variables]
        self.save_vars += [n[:-2] + '/Adam:0' for n in self.student.trainable_names]
        self.save_vars += [n[:-2] + '/Adam_1:0' for n in self.student.trainable_names]
        self.save_vars += ['beta1_power:0', 'beta2_power:0']
        if filter_out is not None:
            self.OPT_FILTER = list(self.OPT_FILTER) + list(filter_out)
        self.filter = lambda elem: elem if all(
            keyword not in elem for keyword in self.OPT_FILTER) and elem not in self.OP_FILTER else None
        if checkpoint is not None:
            self.saver.restore_vars(None, checkpoint, self.filter)
        self.mask = None
        self._mask_on_device = None
        self._fill_thr = None
        self._stop_feeders = False
        print("Semantic Network is ready!!!")

    # ------------------------------------------------------------------ checkpoint surface
    def restore_initial(self):
        self.saver.restore_vars(None, "%s.npy" % self.meta_dir, self.filter)

    def restore(self, chk):
        self.saver.restore_vars(None, chk, self.filter)

    def get_vars(self):
        return self.saver.save_vars(None, self.save_vars, lambda x: x)

    # ------------------------------------------------------------------ inference
    def predict_input(self, frames):
        self.process_lock.acquire()
        frames = np.asarray(frames)
        n = self.student.enqueue(frames, None)
        labels_ = self.student.infer(n, self.bn_mode)
        assert labels_.shape == frames.shape[:-1]
        self.process_lock.release()
        return labels_

    def calc_cross_miou(self, labels):
        assert not self.frozen or self.cross_miou_compat
        assert labels.shape == (2, self.height, 2 * self.height)
        self.process_lock.acquire()
        lab = np.asarray(labels)
        lab = np.where((lab >= 0) & (lab < 255), lab, 255).astype(np.uint8)
        conf_mat_ = self.student.confmat_labels(lab[0], lab[1]).astype(np.float64)
        iou_ = calculate_miou(conf_mat_, nan=True)
        miou_ = np.nanmean(iou_)
        self.process_lock.release()
        return conf_mat_, iou_, miou_

    def predict_with_metric(self, frames, labels_teacher):
        self.process_lock.acquire()
        frames = np.asarray(frames)
        n = self.student.enqueue(frames, labels_teacher)
        labels_student, cm, loss_ = self.student.infer_metric(n, self.bn_mode)
        conf_mat_ = cm.astype(np.float64)                # tf.metrics.mean_iou keeps a float64 matrix of exact integers
        assert labels_student.shape == frames.shape[:-1]
        iou_ = calculate_miou(conf_mat_, nan=True)
        miou_ = np.nanmean(iou_)
        self.process_lock.release()
        return labels_student, conf_mat_, iou_, miou_, loss_

    # ------------------------------------------------------------------ training
    def train_with_deque(self, frame_deque, label_deque, num_of_iterations, train_strategy='full_model',
                         keep_mask=False):
        assert not self.frozen, "Can't train frozen graph!!!"
        if not keep_mask:
            self.mask = None
        self.process_lock.acquire()
        batch_deque = deque()
        batch_thr = threading.Thread(target=self._fill_batch, args=(batch_deque, frame_deque, label_deque,
                                                                    num_of_iterations,))
        batch_thr.start()
        self._stop_feeders = False
        try:
            self._train(batch_deque, num_of_iterations, train_strategy)
        except BaseException:
            # stop both feeder threads and drop what they staged: the next predict_input must not dequeue training batches
            self._stop_feeders = True
            batch_thr.join()
            if self._fill_thr is not None:
                self._fill_thr.join()
            self.student.queue_clear()
            raise
        finally:
            self.process_lock.release()          # held by THIS call since the acquire above; _train never releases it

    def _split(self, flat, dtype=None):
        out = OrderedDict()
        for name, arr in self.student.split_trainable(flat).items():
            out[name] = arr.astype(dtype) if dtype is not None else arr
        return out

    def _upload_mask(self, train_mask_):
        if train_mask_ is None:
            self.student.set_mask(None)
            self._mask_on_device = None
            return
        if self._mask_on_device is train_mask_:
            return
        flat = np.concatenate([np.asarray(train_mask_[n], dtype=np.uint8).reshape(-1) for n in self.student.trainable_names])
        self.student.set_mask(flat)
        self._mask_on_device = train_mask_

    def _train(self, batch_deque, num_of_iterations, train_strategy):
        signal_deque = deque()
        fill_thr = threading.Thread(target=self._fill_queue, args=(batch_deque, num_of_iterations, signal_deque))
        self._fill_thr = fill_thr
        fill_thr.start()
        masked = 'coord_desc_' in train_strategy
        assert not masked or self.masked_gradients, "coord_desc_* strategies need masked_gradients=True at build time"
        _before, train_mask_ = self.get_train_mask(train_strategy)
        for it in range(num_of_iterations):
            signal = None
            while signal is None:
                try:
                    signal = signal_deque.popleft()
                except IndexError:
                    time.sleep(self.THREAD_SLEEP_INTERVAL)
            if isinstance(signal, BaseException):
                raise signal
            t1 = time.time()
            if masked:
                self._upload_mask(train_mask_)
            loss = self.student.train_step(self.lr, masked)
            print('Loss is %.3f at iteration %d and took %.1f ms' % (loss, it, (time.time() - t1) * 1000.0))
            if train_strategy == 'coord_desc_auto':
                if it == 0 and self.mask is None:
                    kept, _thr = self.student.select_topk(self.coord_frac)
                    train_mask_ = self._split(self.student.get_mask(), dtype=bool)
                    self._mask_on_device = train_mask_
                    all_vars = self.student.n_trainable
                    print("Using auto mode, Training %.3f%% of variables" % (100 * kept / all_vars))
                    self.mask = train_mask_
        fill_thr.join()
        if masked:
            self.curr_mask = [train_mask_[name] for name in self.student.trainable_names]
            after = self._split(self.student.get_trainable_flat())
            self.train_params = [after[name] for name in self.student.trainable_names]
        else:
            _after_train = self.saver.save_vars(None, self.save_vars, self.filter)
            self.train_params = [_after_train[name] for name in _after_train.keys()]
            self.curr_mask = [np.ones_like(_after_train[name], dtype=bool) for name in _after_train.keys()]

    def delta_bytes(self):
        """The `<save_dir>_mask.dat` payload run.py:316-328 writes, packed on the device (masked strategies):
        per variable np.packbits(mask), then per variable params[mask].astype(float16)."""
        return self.student.pack_delta()

    def apply_delta(self, blob):
        """Client side of the model stream (an extension: the reference rebuilds a TF session from a whole frozen graph
        at every update, run.py:401-411, and only SIZES the delta file): applies the bytes of delta_bytes() /
        `<save_dir>_mask.dat` to the resident model; returns the number of updated coordinates.  BatchNorm moving
        statistics are not in the delta -- the client keeps the ones of its last full hand-off."""
        self.process_lock.acquire()
        try:
            # the delta's mask becomes the device mask: whatever this object uploaded before is no longer there
            self._mask_on_device = None
            return self.student.apply_delta(blob)
        finally:
            self.process_lock.release()

    def get_train_mask(self, train_strategy):
        names = self.student.trainable_names
        shapes = self.student.var_shapes
        if train_strategy == 'coord_desc_auto':
            self.student.snapshot_before()                       # `_before` stays on the device
            _before = 'device-snapshot'
            if self.mask is None:
                train_mask_ = OrderedDict((n, np.ones(shapes[n], dtype=bool)) for n in names)
            else:
                train_mask_ = self.mask
        elif (train_strategy, self.coord_frac) in _MASK_TABLES:
            substr, exact, bern = _MASK_TABLES[(train_strategy, self.coord_frac)]
            _before = 'device-snapshot'
            train_mask_ = OrderedDict((n, np.zeros(shapes[n], dtype=bool)) for n in names)
            for key in names:
                if any(k in key for k in substr) or key in exact:
                    train_mask_[key] = np.ones(shapes[key], dtype=bool)
                elif key in bern:
                    p = bern[key]
                    train_mask_[key] = np.random.choice([True, False], size=shapes[key], p=[p, 1 - p]).astype(bool)
            all_vars, train_vars_len = self.train_vars_count(train_mask_)
            tag = train_strategy.split('_')[-1] + ('%g' % (100 * self.coord_frac))
            print("Using %s mode, Training %.3f%% of variables" % (tag, 100 * train_vars_len / all_vars))
        elif train_strategy == 'coord_desc_rand':
            _before = 'device-snapshot'
            train_mask_ = OrderedDict(
                (n, np.random.choice([True, False], size=shapes[n], p=[self.coord_frac, 1 - self.coord_frac]).astype(bool))
                for n in names)
            all_vars, train_vars_len = self.train_vars_count(train_mask_)
            print("Using rand mode, Training %.3f%% of variables" % (100 * train_vars_len / all_vars))
        elif train_strategy == 'full_model':
            _before = None
            train_mask_ = None
        else:
            raise NameError('train_strategy %s is not implemented.' % train_strategy)
        return _before, train_mask_

    def train_vars_count(self, train_mask_):
        all_vars = 0
        train_vars_len = 0
        for var_name in self.student.trainable_names:
            train_vars_len += np.sum(train_mask_[var_name])
            all_vars += train_mask_[var_name].size
        return all_vars, train_vars_len

    def _fill_batch(self, batch_deque, frame_deque, label_deque, number_of_batches):
        for batch_index in range(number_of_batches):
            if self._stop_feeders:
                return
            image_batch, label_batch = mini_batch(frame_deque, label_deque, [self.height, self.height * 2], self.scale,
                                                  self.mini_batch_size, 1, flip=False)
            assert np.shape(label_batch) == (1, self.mini_batch_size, self.height, self.height * 2)
            assert np.shape(image_batch) == (1, self.mini_batch_size, self.height, self.height * 2, 3)
            batch_deque.append({'frames': image_batch[0], 'labels': label_batch[0]})

    def _fill_queue(self, batch_deque, number_of_batches, signal_deque):
        for batch_index in range(number_of_batches):
            batch = None
            while batch is None:
                if self._stop_feeders:
                    return
                try:
                    batch = batch_deque.popleft()
                except IndexError:
                    time.sleep(self.THREAD_SLEEP_INTERVAL)
            try:
                frames = batch['frames']
                # mini_batch hands out float64 copies of uint8 frames; ship them as bytes when that is lossless
                f8 = frames.astype(np.uint8)
                if np.array_equal(f8, frames):
                    frames = f8
                self.student.enqueue(frames, batch['labels'])
            except BaseException as e:                  # surface the failure in the trainer thread
                signal_deque.append(e)
                return
            signal_deque.append(1)

    # ------------------------------------------------------------------ frozen hand-off
    def get_frozen_graph(self):
        """The client model: every variable of the graph, to be run with inference-mode BN on the moving statistics
        (reference: trim_graph_frozen(kill_norms=True), utils/graph_utils.py:79-126)."""
        out = OrderedDict()
        for name, _, _, _ in self.student.variables:
            out[name] = self.student.get_tensor(name)
        return out

    def save_to_frozen_graph(self, save_dir):
        """reference :711-714 writes `<save_dir>.pb` (a TF GraphDef); same file name, this library's container
        (C ABI: ams_export_frozen), read back by SemanticNetwork(meta_dir=save_dir, frozen=True)."""
        self.student.export_frozen(save_dir + ".pb")

    def close_model(self):
        self.student.close()

    # ------------------------------------------------------------------ visualisation helpers
    def colorize(self, frame=None, label=None):
        import cv2
        assert frame is not None or label is not None, "At least a label or frame must be given"
        assert frame is None or frame.shape == (self.height, self.height * 2, 3)
        if label is None:
            label = self.predict_input(np.expand_dims(frame, axis=0))[0]
        assert label.shape == (self.height, self.height * 2)
        label_colored = self.color_map_reduced_[label]
        if frame is not None:
            return label_colored, cv2.addWeighted(frame, 0.5, label_colored, 0.5, 0)
        return label_colored

    def colorize_teacher(self, label, frame=None):
        import cv2
        assert frame is None or frame.shape == (self.height, self.height * 2, 3)
        assert label.shape == (self.height, self.height * 2)
        label_colored = colormap()[label]
        if frame is not None:
            return label_colored, cv2.addWeighted(frame, 0.5, label_colored, 0.5, 0)
        return label_colored

    def cross_ignore(self, label_teacher, label_student=None, frame_student=None):
        assert label_student is not None or frame_student is not None, \
            "At least a label or frame from student must be given"
        assert label_teacher.shape == (self.height, self.height * 2)
        label_teacher_reduced = self.take_array[label_teacher]
        if label_student is None:
            label_student = self.predict_input(np.expand_dims(frame_student, axis=0))[0]
        assert label_student.shape == (self.height, self.height * 2)
        ignore_mask = np.where(np.expand_dims(label_teacher_reduced, axis=-1) == 0, self.WHITE, self.BLACK)
        colorized_label_teacher = self.colorize(label=label_teacher_reduced)
        cross_cond = np.logical_and(np.logical_not(ignore_mask[:, :, :1]),
                                    np.expand_dims(np.not_equal(label_teacher_reduced, label_student), axis=-1))
        cross_mask = np.where(cross_cond, colorized_label_teacher, self.BLACK)
        assert ignore_mask.shape == cross_mask.shape
        assert ignore_mask.shape == (self.height, self.height * 2, 3)
        return cross_mask, ignore_mask
