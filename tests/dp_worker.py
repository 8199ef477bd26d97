"""Worker of tests/test_dp_gpu.py: one rank of a 2-GPU data-parallel distillation job (launched with torch.distributed.run).

What is checked (SURVEY 8e caveat 2): with the BatchNorm statistics summed over the ranks inside the BN kernels
(NVLink peer memory, ams_syncbn_*), a step on 2 GPUs x B frames equals the single-process step on the same 2B frames
-- which is what the reference computes (FusedBatchNormV3(is_training=True) over the whole batch in one process):
  * EXACT: when every rank holds the SAME B frames, all sums double and so does n, so the step must be bit-identical to
    the single-GPU step on those B frames (loss, every gradient coordinate x world, every BN moving mean) -- a
    chaos-free check of the whole exchange (forward, backward, image-pooling branch);
  * different frames per rank: BN moving statistics and gradients agree with a single-GPU run of the global batch up to
    the summation order of the fp64 partial sums, which a random-init BN/ReLU stack amplifies, and are far
    closer to it than the per-replica mode;
  * parameters and moving statistics stay BIT-IDENTICAL across ranks over eager, captured and graph-replayed steps;
  * the per-replica mode (sync off) is measurably different from the global-batch result (the test can tell them apart).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

H, W, B = 192, 384, 2
LR = 1e-3
# argv[1]: 'cityscapes' (19-class graph, every class) or 'pascalvoc2012' (config C4: 21-class graph, BN decay 0.98,
# class vector of reference experiment 40, exp_configs.py:152-154)
TAG = sys.argv[1] if len(sys.argv) > 1 else 'cityscapes'
NUM_CLASSES = 21 if TAG == 'pascalvoc2012' else 19
CLASSES = [0, 7, 12, 15] if TAG == 'pascalvoc2012' else list(range(19))


def rel(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    import torch
    import torch.distributed as dist
    from ams_b200.parallel import DataParallelStudent
    from ams_b200.student import Student
    from ams_b200.synthetic import synthetic_checkpoint, synthetic_frames, synthetic_labels

    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    ckpt = synthetic_checkpoint(TAG, 1)
    steps = 4
    frames = [synthetic_frames(B * world, H, W, seed=10 + i) for i in range(steps)]
    labels = [synthetic_labels(B * world, H, W, seed=10 + i, block=16) for i in range(steps)]
    moving_names = [k for k in ckpt if k.endswith('moving_mean:0') or k.endswith('moving_variance:0')]

    def make():
        st = Student(NUM_CLASSES, H, W, CLASSES, device=local, queue_capacity=steps + 2)
        st.set_stream(stream.cuda_stream)
        for k, v in ckpt.items():
            st.set_tensor(k, v)
        return st

    def moving(st):
        return np.concatenate([st.get_tensor(k).ravel() for k in moving_names])

    def one_step_grads(sync_bn, tag, duplicate=False):
        """one exchange step from the checkpoint: (gradient sums / n_valid, moving statistics, loss)"""
        st = make()
        dp = DataParallelStudent(st, sync_bn=sync_bn)
        lo = 0 if duplicate else rank * B
        st.enqueue(frames[0][lo:lo + B], labels[0][lo:lo + B])
        st.train_forward_backward_async()
        dist.all_reduce(dp.grad)
        dist.all_reduce(dp.terms)
        st.synchronize()
        nv, ls = [float(x) for x in dp.terms.cpu()]
        g = st.get_gradients() if duplicate else st.get_gradients() / nv
        mv = moving(st)
        if sync_bn:
            ep, err = st.syncbn_status()
            assert err == 0 and ep == 1, (tag, ep, err)
        dp.close()
        st.close()
        return g, mv, ls / nv

    g_sync, mv_sync, loss_sync = one_step_grads(True, 'sync')
    g_rep, mv_rep, loss_rep = one_step_grads(False, 'replica')
    g_dup, mv_dup, loss_dup = one_step_grads(True, 'duplicate', duplicate=True)

    ok = True
    if rank == 0:
        # exact: the same B frames on every rank == the single-GPU step on those B frames
        one = make()
        one.enqueue(frames[0][:B], labels[0][:B])
        nv1, ls1 = one.train_forward_backward()
        g1 = one.get_gradients()
        mv1 = moving(one)
        one.close()
        is_mean = np.concatenate([np.full(ckpt[k].size, k.endswith('moving_mean:0')) for k in moving_names])
        n_bad_g = int(np.count_nonzero(g_dup != np.float32(world) * g1))
        n_bad_m = int(np.count_nonzero(mv_dup[is_mean] != mv1[is_mean]))
        loss_same = (loss_dup == ls1 / nv1)
        print('[dp] ' + TAG + ': duplicate frames on %d ranks vs one GPU: %d of %d gradient coordinates differ, %d of %d moving means '
              'differ, loss identical: %s' % (world, n_bad_g, g1.size, n_bad_m, int(is_mean.sum()), loss_same), flush=True)
        ok &= n_bad_g == 0 and n_bad_m == 0 and bool(loss_same)
        ref = make()
        ref.enqueue(frames[0], labels[0])
        loss_ref = float(ref.train_step(LR, False))
        g_ref = ref.get_gradients()
        mv_ref = moving(ref)
        ref.close()
        e_mv, e_g = rel(mv_sync, mv_ref), rel(g_sync, g_ref)
        r_mv, r_g = rel(mv_rep, mv_ref), rel(g_rep, g_ref)
        print('[dp] ' + TAG + ': global batch %d @ %dx%d on %d GPUs vs one GPU: moving stats rel-L2 %.3e (per-replica BN: %.3e), '
              'gradients rel-L2 %.3e (per-replica BN: %.3e), loss %.6f vs %.6f (per-replica %.6f)'
              % (B * world, H, W, world, e_mv, r_mv, e_g, r_g, loss_sync, loss_ref, loss_rep), flush=True)
        ok &= e_mv < 2e-3 and abs(loss_sync - loss_ref) < 5e-3 * abs(loss_ref)        # measured: 3e-4 (cityscapes), 1.6e-3 (VOC, 4 classes)
        ok &= r_mv > 10 * e_mv and r_g > 3 * e_g                  # the per-replica mode is much further away

    # ---- several full steps (eager, capture, replay): ranks stay bit-identical, no exchange error
    st = make()
    dp = DataParallelStudent(st, sync_bn=True)
    for i in range(steps):
        st.enqueue(frames[i][rank * B:(rank + 1) * B], labels[i][rank * B:(rank + 1) * B])
    for i in range(steps):
        dp.train_step_async(LR, False)
    losses = dp.losses()
    ep, err = st.syncbn_status()
    params = torch.from_numpy(st.get_trainable_flat()).cuda()
    mv = torch.from_numpy(moving(st)).cuda()
    allp = [torch.empty_like(params) for _ in range(world)]
    allm = [torch.empty_like(mv) for _ in range(world)]
    dist.all_gather(allp, params)
    dist.all_gather(allm, mv)
    same = all(torch.equal(allp[0], t) for t in allp) and all(torch.equal(allm[0], t) for t in allm)
    if rank == 0:
        print('[dp] %d steps with the exchange: losses %s, epoch %d, error %d, ranks bit-identical: %s'
              % (steps, ['%.5f' % x for x in losses], ep, err, same), flush=True)
    ok &= same and err == 0 and ep == steps and len(losses) == steps and bool(np.all(np.isfinite(losses)))
    # bucketed exchange (late-layer bucket reduced on the communication stream WHILE backward continues, released by an event
    # recorded in the middle of the step / an external event node of the replayed graph) == one flat allreduce after backward
    st_f = make()
    dp_f = DataParallelStudent(st_f, sync_bn=True, buckets=False)
    for i in range(steps):
        st_f.enqueue(frames[i][rank * B:(rank + 1) * B], labels[i][rank * B:(rank + 1) * B])
    for i in range(steps):
        dp_f.train_step_async(LR, False)
    losses_f = dp_f.losses()
    p_flat = st_f.get_trainable_flat()
    p_buck = params.cpu().numpy()
    d_rel = rel(p_buck, p_flat)
    n_diff = int(np.count_nonzero(p_buck != p_flat))
    if rank == 0:
        print('[dp] bucketed (split at %d of %d floats, overlapped) vs flat allreduce after %d steps: %d parameters differ, rel-L2 %.2e, '
              'losses equal: %s' % (dp.split, p_flat.size, steps, n_diff, d_rel, losses == losses_f), flush=True)
    ok &= dp.split > 0 and dp.comm_stream is not None and d_rel < 1e-6 and (world > 2 or n_diff == 0)
    dp_f.close()
    st_f.close()
    # switching the exchange off and on again re-captures the step
    st.syncbn_enable(False)
    st.enqueue(frames[0][rank * B:(rank + 1) * B], labels[0][rank * B:(rank + 1) * B])
    dp.train_step_async(LR, False)
    st.syncbn_enable(True)
    st.enqueue(frames[1][rank * B:(rank + 1) * B], labels[1][rank * B:(rank + 1) * B])
    dp.train_step_async(LR, False)
    l2 = dp.losses()
    ok &= bool(np.all(np.isfinite(l2)))
    dp.close()
    st.close()

    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        print('[dp] rank %d FAILED' % rank, flush=True)
        sys.exit(1)
    if rank == 0:
        print('[dp] OK', flush=True)


if __name__ == '__main__':
    main()
