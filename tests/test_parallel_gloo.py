"""world_size-2 `gloo` test (CPU) of the data-parallel host logic: the exchange step (gradient-arena allreduce +
n_valid/loss_sum allreduce + 1/n_valid scaling) makes two ranks, each holding half of the batch, produce exactly the
single-process loss and (up to fp32 summation order) gradients of the reference's mean over all valid pixels.
The oracle stands in for the device model here (BN on moving statistics, so the halves do not couple)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
        sys.path.insert(0, p)
    import student_oracle as so
    from ams_b200.parallel import allreduce_step_terms, shard_streams
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    spec = so.load_spec('cityscapes')
    V = so.synthetic_variables(spec, 5)
    frames = so.synthetic_frames(4, 32, 48, 9).astype(np.float32)
    labels = so.synthetic_labels(4, 32, 48, 9, block=8)
    cls = np.array([0, 1, 2, 8, 10, 11, 13])
    mine = shard_streams(4, world, rank)
    ts = so.TrainState(spec, V)
    # local SUM loss and its gradient (what ams_train_forward_backward leaves in the arena)
    params = {k: torch.tensor(v, requires_grad=(k in ts.m)) for k, v in ts.vars.items()}
    sem, _ = so.forward(spec, params, frames[mine], bn_mode='moving')
    h = so.head(so.full_res_logits(sem, 32, 48), labels[mine], cls)
    loss_sum = h['loss'] * h['n_valid']
    grads = torch.autograd.grad(loss_sum, [params[k] for k in ts.trainable], allow_unused=True)
    flat = torch.cat([(torch.zeros_like(params[k]) if g is None else g).reshape(-1) for k, g in zip(ts.trainable, grads)])
    scale, loss = allreduce_step_terms(flat, h['n_valid'], float(loss_sum), None)
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank), (flat * scale).numpy())
    if rank == 0:
        # single-process reference on the whole batch
        params = {k: torch.tensor(v, requires_grad=(k in ts.m)) for k, v in ts.vars.items()}
        sem, _ = so.forward(spec, params, frames, bn_mode='moving')
        h = so.head(so.full_res_logits(sem, 32, 48), labels, cls)
        g = torch.autograd.grad(h['loss'], [params[k] for k in ts.trainable], allow_unused=True)
        ref = torch.cat([(torch.zeros_like(params[k]) if t is None else t).reshape(-1) for k, t in zip(ts.trainable, g)])
        np.save(os.path.join(out_dir, 'ref.npy'), ref.numpy())
        np.save(os.path.join(out_dir, 'loss.npy'), np.array([loss, float(h['loss'])]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1, ref = (np.load(os.path.join(tmp_path, f)) for f in ('rank0.npy', 'rank1.npy', 'ref.npy'))
    loss = np.load(os.path.join(tmp_path, 'loss.npy'))
    assert np.array_equal(r0, r1)                                   # every rank ends with the same gradient
    assert abs(loss[0] - loss[1]) < 1e-5 * max(1.0, abs(loss[1]))
    denom = np.abs(ref).max()
    assert np.abs(r0 - ref).max() < 2e-5 * denom


class _StubStudent:
    """Host-side stand-in for `Student`: the 'device' arena and terms are CPU tensors, the kernels are two lines."""

    def __init__(self, rank):
        self.rank = rank
        self.grad = torch.zeros(6, dtype=torch.float32)
        self.terms = torch.zeros(2, dtype=torch.float64)
        self.params = torch.ones(6, dtype=torch.float32)
        self.step = 0

    def train_forward_backward_async(self):
        self.step += 1
        self.grad[:] = torch.arange(6, dtype=torch.float32) * (self.rank + 1) * self.step
        self.terms[0], self.terms[1] = 10.0 * (self.rank + 1), 3.0 * (self.rank + 1) * self.step

    def apply_optimizer_device(self, lr, masked, loss_out):
        scale = 1.0 / float(self.terms[0]) if float(self.terms[0]) > 0 else 0.0
        self.params -= lr * self.grad * scale
        loss_out[0] = float(self.terms[1] / self.terms[0])

    def synchronize(self):
        pass

    def gradient_bucket_split(self):
        return 4                                  # two buckets: [0, 4) early layers, [4, 6) late layers


def _dp_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from ams_b200.parallel import DataParallelStudent
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    st = _StubStudent(rank)
    dp = DataParallelStudent(st, tensors=(st.grad, st.terms))
    for _ in range(3):
        dp.train_step_async(0.5, False)
    losses = dp.losses()
    assert dp.losses() == []                                        # drained
    one = dp.train_step(0.5, False)                                 # synchronous form returns this step's loss
    np.save(os.path.join(out_dir, 'dp%d.npy' % rank), np.concatenate([st.params.numpy(), np.array(losses + [one], np.float32)]))
    dp.close()
    dist.destroy_process_group()


def test_data_parallel_student_host_logic(tmp_path):
    """order of the exchange step (backward -> allreduce of the late and the early gradient bucket, allreduce(terms) ->
    Adam reading the global terms), loss slots filled per step and drained by losses(), identical state on both ranks"""
    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (np.load(os.path.join(tmp_path, 'dp%d.npy' % r)) for r in range(2))
    assert np.array_equal(a, b)
    # global terms: n = 10 + 20, loss_sum = 3 s + 6 s -> loss = 0.3 s; gradient sum = arange * 3 s, scaled by 1/30
    want_p = np.ones(6, np.float32)
    for s_ in (1, 2, 3, 4):
        want_p = want_p - np.float32(0.5) * (np.arange(6, dtype=np.float32) * 3 * s_) * np.float32(1.0 / 30.0)
    assert np.allclose(a[:6], want_p, rtol=1e-6)
    assert np.allclose(a[6:], [0.3, 0.6, 0.9, 1.2], rtol=1e-6)


def test_shard_streams_partition():
    from ams_b200.parallel import shard_streams
    for world in (1, 2, 4, 8):
        got = sorted(s for r in range(world) for s in shard_streams(8, world, r))
        assert got == list(range(8))
    assert shard_streams(8, 4, 1) == [1, 5]


def test_single_process_terms_without_group():
    from ams_b200.parallel import allreduce_step_terms
    g = torch.ones(4)
    scale, loss = allreduce_step_terms(g, 8, 4.0)
    assert scale == 0.125 and loss == 0.5
    scale, loss = allreduce_step_terms(g, 0, 0.0)
    assert scale == 0.0 and np.isnan(loss)
