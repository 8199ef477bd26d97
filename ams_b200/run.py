"""Server / client orchestration of AMS on top of ams_b200.SemanticNetwork.

Mirrors the control flow of the reference's run.py -- `train_model` (server: frame sampling at `send_period`, bounded
frame/label memory, periodic distillation phases, ASR sampling-rate control from the teacher-label phi-score, ATR
training-period control, down-link accounting from the gzipped delta file, model hand-off per update) and
`infer_output` (client: load the model scheduled for each second, per-frame prediction + mIoU, sliding 10 s mIoU),
and the `simple` / `early` / `pretrained` / `horizon` drivers -- with the section-4 bugs of the reference fixed forward
(flags read from one namespace, frames appended to the label memory as frames, integer ranges).

Differences, all at the edges of the hot path:
  * frames come from a FrameSource (cv2.VideoCapture + `gt_%06d.png` files like the reference, or a synthetic
    generator for tests and benchmarks); nothing else in the loop knows where frames come from;
  * the delta file is written from `SemanticNetwork.delta_bytes()` (packed on the device, byte-identical to the host
    writer of run.py:316-328); the client can apply it in place (`apply_delta`) or reload the exported model;
  * H.264 up-link emulation (`compress_uplink`: two-pass ffmpeg encode + re-decode of the sampled frames, reference
    run.py:196-260) is NOT built -- it is outside the hot path (out of scope) -- and the flag is refused with
    NotImplementedError rather than silently falling back; the PNG accounting path (the reference's default) is what
    runs, and frame_memory only ever holds RGB frames at the network size; plotting is out of scope.
"""
import argparse
import os
import subprocess as sp
import tempfile
import time
from collections import deque

import numpy as np

from .SemanticNetwork import SemanticNetwork
from .exp_configs import class_weights, coco_class_converter, is_coco, test_length
from .utils.utils import calculate_miou, choose_frames, string_class_iou


# --------------------------------------------------------------------------------------------- configuration
def default_flags():
    """The reference's flag set with its defaults (run.py:18-70)."""
    return argparse.Namespace(
        input_video=None, gt_video=None, student_checkpoint=None, output_dir=None, gpu='0',
        initial_fill=False, memory_len=250, batch_size=10, iter=200, height=256, lr=1e-3, send_period=30, train_period=10,
        only_results=False, compress_uplink=False, uplink_bw=200, no_restore=False, save_pic=False, enable_ASR=False,
        enable_ATR=False, train_strategy='full_model', coord_fraction='0.1', mode=None, early_cutoff_time=60,
        client_applies_delta=False)


def parse_flags(argv=None):
    d = default_flags()
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    for k, v in vars(d).items():
        if isinstance(v, bool):
            ap.add_argument('--' + k, action='store_true', default=v)
        else:
            ap.add_argument('--' + k, type=type(v) if v is not None else str, default=v)
    return ap.parse_args(argv)


# --------------------------------------------------------------------------------------------- frame sources
class VideoSource:
    """cv2.VideoCapture + ground-truth label PNGs named gt_%06d.png (run.py:104-111, :160-165)."""

    def __init__(self, video_path, gt_path):
        import cv2
        self._cv2 = cv2
        self.cap = cv2.VideoCapture(video_path)
        if not self.cap.isOpened():
            raise IOError('Error opening video stream or file: %s' % video_path)
        self.fps = int(round(self.cap.get(cv2.CAP_PROP_FPS)))
        self.gt_path = gt_path
        self.pos = 0

    def seek(self, frame_index):
        self.cap.set(self._cv2.CAP_PROP_POS_FRAMES, frame_index)
        self.pos = frame_index

    def read(self):
        ret, frame = self.cap.read()
        if not ret:
            raise EOFError('Premature end of video')
        gt = self._cv2.imread('%sgt_%06d.png' % (self.gt_path, self.pos), self._cv2.IMREAD_GRAYSCALE)
        self.pos += 1
        return frame, gt

    def close(self):
        self.cap.release()


class SyntheticSource:
    """Deterministic stand-in for a camera + teacher: BGR uint8 frames and piecewise-constant label maps that drift
    slowly with time (so consecutive teacher maps overlap, as the phi-score expects)."""

    def __init__(self, height, width, fps=5, num_ids=19, block=32, seed=0):
        self.h, self.w, self.fps, self.num_ids, self.block, self.seed = height, width, fps, num_ids, block, seed
        self.pos = 0

    def seek(self, frame_index):
        self.pos = frame_index

    def read(self):
        i = self.pos
        self.pos += 1
        scene = i // (4 * self.fps)                                   # the label layout changes every 4 s
        rng = np.random.default_rng(self.seed + 7919 * scene)
        coarse = rng.integers(0, self.num_ids, size=(-(-self.h // self.block), -(-self.w // self.block)), dtype=np.uint8)
        gt = np.repeat(np.repeat(coarse, self.block, axis=0), self.block, axis=1)[:self.h, :self.w].copy()
        frng = np.random.default_rng(self.seed + 104729 * i)
        frame = (gt[..., None].astype(np.int32) * 13 + frng.integers(0, 64, size=(self.h, self.w, 3))).astype(np.uint8)
        return frame, gt

    def close(self):
        pass


# --------------------------------------------------------------------------------------------- ASR / ATR controllers
def asr_update(send_rate, phi_scores):
    """Adaptive sampling rate (run.py:279-291): phi = mean cross-mIoU of consecutive teacher maps since the last phase;
    send_rate -= 0.2 * tanh(20 * (phi - 0.6)), clipped to [0.1, 1]."""
    send_rate = send_rate - 0.2 * np.tanh((np.mean(phi_scores) - 0.6) * 20)
    return float(np.clip(send_rate, 0.1, 1))


def atr_update(hibernate, train_period_current, train_period_reset, send_rate_history):
    """Adaptive training rate (run.py:293-303): hibernate below a mean send rate of 0.25, wake above 0.35; while
    hibernating the training period grows by 2 s per phase up to 6x its nominal value.  Returns (hibernate, period)."""
    mean_rate = np.mean(list(send_rate_history)) if len(send_rate_history) else 1.0
    if mean_rate < 0.25:
        hibernate = True
    if mean_rate > 0.35 and hibernate:
        hibernate = False
        train_period_current = train_period_reset
    if hibernate:
        train_period_current = min(train_period_current + 2, 6 * train_period_reset)
    return hibernate, train_period_current


def reschedule(save_range, now, train_end, train_period_current):
    """ATR: keep the past save points, re-space the future ones (run.py:304-308)."""
    idx = save_range.index(now)
    out = save_range[:idx]
    out.extend(range(now, train_end, train_period_current))
    assert now in out
    return out


def simple_event_list(length, train_period, memory_len, initial_fill):
    """Model-update times of mode 'simple' (run.py:593-598; the float range of the reference made integer)."""
    first_train = int(np.ceil(100 / train_period) * train_period)
    events = [0]
    events.extend(t for t in range(first_train, length, train_period) if t == 0 or t >= memory_len or not initial_fill)
    return events


# --------------------------------------------------------------------------------------------- helpers
def get_save_dir(flags, prepend):
    """run.py:563-573: '<output_dir><prepend>_<video file>_<checkpoint dir>_<height>'."""
    video = (flags.input_video or 'synthetic').split('/')[-1]
    ckpt = (flags.student_checkpoint or 'checkpoint/model').rstrip('/').split('/')
    return os.path.join(flags.output_dir, '%s_%s_%s_%d' % (prepend, video, ckpt[-2] if len(ckpt) > 1 else ckpt[-1], flags.height))


def _resize_pair(flags, frame, label):
    import cv2
    size = (flags.height * 2, flags.height)
    # reference run.py:181-183 (PNG path): network-size RGB frame, nearest-neighbour label map
    frame = cv2.cvtColor(cv2.resize(frame, size), cv2.COLOR_BGR2RGB)
    return frame, cv2.resize(label, size, interpolation=cv2.INTER_NEAREST)


def _png_kilobytes(frame):
    import cv2
    with tempfile.NamedTemporaryFile(suffix='.png', delete=False) as f:
        path = f.name
    try:
        cv2.imwrite(path, frame)
        return os.path.getsize(path) / 1024
    finally:
        os.remove(path)


def write_delta_file(semantic_network, path):
    """`<save_dir>_mask.dat` (run.py:316-328) + `gzip -9 -f -k`; returns the down-link bits of this update."""
    blob = semantic_network.delta_bytes()
    with open(path, 'wb') as f:
        f.write(blob)
    sp.Popen(['gzip', '-9', '-f', '-k', path]).wait()
    return os.path.getsize(path + '.gz') * 8, blob


# --------------------------------------------------------------------------------------------- server
def train_model(flags, source, train_start, train_end, sampling_period, gpu_id, run_label, exp_num, save_range,
                sample_send_period, log=print):
    """Server side (run.py:78-361).  Returns a dict of the per-period logs it also writes to `<run_label>_results_*`."""
    assert train_end - train_start != 0, 'There should be at least one set of data points'
    if flags.compress_uplink:
        raise NotImplementedError('compress_uplink (two-pass H.264 encode / re-decode of the up-link through ffmpeg, reference '
                                  'run.py:196-260) is outside the B200 hot path and not built: run without it (PNG accounting)')
    fps = source.fps
    train_end_frame = train_end * fps
    i = train_start * fps
    source.seek(i)
    save_range = list(save_range)
    update_count = 0
    send_rate = sampling_period / fps
    sample_per_period, up_bw_per_period, down_bw_per_period = [], [], []
    frame_label_bucket = []
    num_unseen_frames = 0
    model_save_times = [0]
    train_period_reset = train_period_current = (save_range[2] - save_range[1]) if len(save_range) > 2 else flags.train_period
    send_rate_deq = deque(maxlen=5)
    hibernate = False
    map_coco = coco_class_converter() if is_coco(exp_num) else None
    mem = max(1, int(flags.memory_len / sampling_period * fps))
    frame_memory, label_memory, pending = deque(maxlen=mem), deque(maxlen=mem), deque(maxlen=mem)
    net = SemanticNetwork(meta_dir=flags.student_checkpoint, class_weights_exp=class_weights(exp_num), height=flags.height,
                          gpu_id=gpu_id, scale=[1], mini_batch_size=flags.batch_size, lr=flags.lr, mem_frac=1,
                          coord_frac=float(flags.coord_fraction), train_biases_only=False, regularize=False,
                          masked_gradients=flags.train_strategy not in ['full_model'], cross_miou_compat=flags.enable_ASR)
    save_dir = get_save_dir(flags, run_label + '_%d' % train_start)
    net.save_to_frozen_graph(save_dir + '_final')
    log('Saved model to %s_final.pb' % save_dir)
    deltas = []
    while i < train_end_frame:
        frame_label_bucket.append(source.read())
        i += 1
        if i % fps == 0 and i // fps % sample_send_period == 0:        # once per send period (the reference re-enters per frame)
            frames_chosen, labels_chosen = choose_frames(frame_label_bucket, send_rate)
            for frame, label in zip(frames_chosen, labels_chosen):
                frame, label = _resize_pair(flags, frame, label)
                pending.append(frame)
                label_memory.append(map_coco[label] if map_coco is not None else label)
            frame_label_bucket.clear()
            num_frames = len(pending)
            sample_per_period.append(num_frames)
            num_unseen_frames += num_frames
            size_images = 0.0
            while pending:
                f = pending.popleft()
                size_images += _png_kilobytes(f)
                frame_memory.append(f)
            up_bw_per_period.append(size_images * 8)
        if i % fps == 0 and i // fps in save_range and len(frame_memory) > 0:
            now = i // fps
            if flags.enable_ASR and len(label_memory) > 1:
                i_start = max(0, len(label_memory) - num_unseen_frames - 1)
                phi = [net.calc_cross_miou(np.array([label_memory[k], label_memory[k + 1]]))[2]
                       for k in range(i_start, len(label_memory) - 1)]
                if phi:
                    send_rate = asr_update(send_rate, phi)
                    send_rate_deq.append(send_rate)
                    log('Send rate updated to %.2f' % send_rate)
                num_unseen_frames = 0
            if flags.enable_ATR:
                hibernate, train_period_current = atr_update(hibernate, train_period_current, train_period_reset, send_rate_deq)
                save_range = reschedule(save_range, now, train_end, train_period_current)
            if not flags.no_restore:
                net.restore_initial()
            t1 = time.time()
            net.train_with_deque(frame_memory, label_memory, flags.iter, flags.train_strategy)
            log('Training for %d iterations took %d ms' % (flags.iter, 1000 * (time.time() - t1)))
            bits, blob = write_delta_file(net, save_dir + '_mask.dat')
            deltas.append((now, blob))
            down_bw_per_period.append(bits)
            update_count += 1
            save_dir = get_save_dir(flags, run_label + '_%d' % now)
            net.save_to_frozen_graph(save_dir + '_final')
            model_save_times.append(i / fps)
    net.close_model()
    final = get_save_dir(flags, run_label + '_results')
    np.save(final + '_fps_client.npy', sample_per_period)
    np.save(final + '_bw_uplink.npy', up_bw_per_period)
    np.save(final + '_bw_downlink.npy', down_bw_per_period)
    np.save(final + '_model_update_times.npy', model_save_times)
    with open(final + '_update.txt', 'w') as f:
        f.write('%d\n%d\n%d\n%d\n%d' % (sum(down_bw_per_period), sum(up_bw_per_period), update_count, train_end - train_start,
                                        sum(sample_per_period)))
    return {'samples': sample_per_period, 'uplink_bits': up_bw_per_period, 'downlink_bits': down_bw_per_period,
            'model_update_times': model_save_times, 'update_count': update_count, 'deltas': deltas, 'save_range': save_range}


# --------------------------------------------------------------------------------------------- client
def infer_output(flags, source, inf_start, inf_end, gpu_id, run_label, exp_num, load_range, log=print):
    """Client side (run.py:364-461): per-frame prediction against the ground-truth maps with the model in force."""
    import cv2
    assert inf_end - inf_start != 0, 'There should be at least one set of data points'
    fps = source.fps
    i = inf_start * fps
    source.seek(i)
    size = (flags.height * 2, flags.height)
    net = None
    conf_mem = deque(maxlen=10 * fps)
    loss_s, miou_cats, miou_s, miou_mem_s = [], [], [], []
    final = get_save_dir(flags, run_label + '_results')
    while i < inf_end * fps:
        if i % fps == 0 and i // fps in load_range:
            save_dir = get_save_dir(flags, run_label + '_%d' % (i // fps))
            if net is not None:
                net.close_model()
            net = SemanticNetwork(meta_dir=save_dir + '_final', class_weights_exp=class_weights(exp_num), height=flags.height,
                                  gpu_id=gpu_id, mem_frac=1, frozen=True)
        frame, gt = source.read()
        frame = cv2.cvtColor(cv2.resize(frame, size), cv2.COLOR_BGR2RGB)
        gt = cv2.resize(gt, size, interpolation=cv2.INTER_NEAREST)
        _, conf_mat, _, miou, loss = net.predict_with_metric(np.expand_dims(frame, 0), np.expand_dims(gt, 0))
        loss_s.append(loss)
        miou_cats.append(np.array(conf_mat))
        miou_s.append(miou)
        conf_mem.append(conf_mat)
        miou_mem_s.append(np.nanmean(calculate_miou(np.sum(list(conf_mem), axis=0), nan=True)))
        i += 1
        if i % fps == 0:
            m = np.nanmean(calculate_miou(np.sum(miou_cats[-fps:], axis=0), nan=True))
            log('miou at %03d secs: %.1f%%' % (i / fps, float(m) * 100))
    np.save('%s_loss.npy' % final, loss_s)
    np.save('%s_mioucats.npy' % final, miou_cats)
    np.save('%s_mious.npy' % final, miou_s)
    np.save('%s_mioumems.npy' % final, miou_mem_s)
    if net is not None:
        net.close_model()
    return {'loss': loss_s, 'miou': miou_s, 'miou_mem': miou_mem_s, 'conf_mats': miou_cats}


# --------------------------------------------------------------------------------------------- drivers
def main(argv=None, source_factory=None):
    flags = parse_flags(argv)
    os.makedirs(flags.output_dir, exist_ok=True)
    vid_num = int(flags.input_video.split('/')[-1].split('-')[0])
    make = source_factory or (lambda: VideoSource(flags.input_video, flags.gt_video))
    length = test_length(vid_num)

    def run(train_args, infer_args):
        if flags.only_results:
            return
        src = make()
        out = train_model(flags, src, *train_args)
        src.close()
        events = out['model_update_times'] if flags.enable_ATR else infer_args[-1]
        src = make()
        infer_output(flags, src, *infer_args[:-1], [int(e) for e in events])
        src.close()

    if flags.mode == 'simple':
        label = '%d__%d_tp%d_f%d' % (0, length, flags.train_period, flags.send_period)
        events = simple_event_list(length, flags.train_period, flags.memory_len, flags.initial_fill)
        run((0, length, flags.send_period, flags.gpu, label, vid_num, events, flags.train_period),
            (0, length, flags.gpu, label, vid_num, events))
    elif flags.mode == 'early':
        label = 'early%d_f%d' % (flags.early_cutoff_time, flags.send_period)
        events = [0, flags.early_cutoff_time]
        run((0, flags.early_cutoff_time, flags.send_period, flags.gpu, label, vid_num, events, flags.train_period),
            (0, length, flags.gpu, label, vid_num, events))
    elif flags.mode == 'pretrained':
        run((0, 1, flags.send_period, flags.gpu, 'pretrained', vid_num, [0], flags.train_period),
            (0, length, flags.gpu, 'pretrained', vid_num, [0]))
    elif flags.mode == 'horizon':
        k1s, k2, points = [16, 32, 64, 128, 256, 512], 256, 3
        step = (length - k2 - k1s[-1]) // (points - 1)
        run((0, 1, flags.send_period, flags.gpu, 'pretrained', vid_num, [0], flags.train_period),
            (0, length, flags.gpu, 'pretrained', vid_num, [0]))
        for p in range(points):
            t = k1s[-1] + p * step
            for k1 in k1s:
                label = '%d__%d__%d_f%d' % (t - k1, t, t + k2, flags.send_period)
                run((t - k1, t, flags.send_period, flags.gpu, label, vid_num, [t], flags.train_period),
                    (t, t + k2, flags.gpu, label, vid_num, [t]))
    else:
        raise ValueError('mode must be one of simple, horizon, early, pretrained')
    print('Process [Main]: Done!!!')


if __name__ == '__main__':
    main()
