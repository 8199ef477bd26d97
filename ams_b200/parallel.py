"""Multi-GPU plumbing for the two places where the AMS hot path shards.  The reference is single-GPU
(`visible_device_list`, SemanticNetwork.py:74); this is the B200 8-GPU extension of the same step.

  * inference: streams are independent -> `shard_streams` assigns stream s to rank s % world, no collective;
  * distillation: data parallel, one process per GPU.  Every rank runs forward/backward on its 8 frames with the
    loss left as a SUM over its valid pixels, and the flat fp32 gradient arena (2,113,043 floats = 8.45 MB) is summed
    over the ranks in TWO buckets (NCCL over NVLink): the late-layer bucket (final-resolution stage + ASPP + logits,
    ~90 % of the coordinates) is complete after ~45 % of the backward pass and its allreduce runs on a communication
    stream WHILE the high-resolution layers are still in backward; the small early-layer bucket and the
    (n_valid, loss_sum) pair follow when backward ends.  Adam then runs identically on every rank with gradients
    scaled by 1 / global n_valid, which reproduces the reference's `reduce_mean` over all valid
    pixels of the global batch (utils/graph_utils.py:408).  BatchNorm batch statistics are summed over the ranks inside
    the BN finalize kernels through NVLink peer memory (sync_bn=True: the step equals the reference's single-process
    step on the global batch) or stay per replica (sync_bn=False, a documented deviation).
torch.distributed is plumbing only; on CPU the same code runs over gloo (tests/test_parallel_gloo.py).
"""
import torch
import torch.distributed as dist


def shard_streams(num_streams, world_size, rank):
    """Streams handled by `rank`: s % world_size == rank (no communication between shards)."""
    return [s for s in range(num_streams) if s % world_size == rank]


def allreduce_step_terms(grad, n_valid, loss_sum, group=None):
    """Sum the gradient arena and the (n_valid, loss_sum) pair over ranks, in place.
    Returns (grad_scale, global_mean_loss): grad_scale = 1 / global n_valid (0 valid pixels -> scale 0, NaN loss)."""
    terms = torch.tensor([float(n_valid), float(loss_sum)], dtype=torch.float64, device=grad.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(terms, op=dist.ReduceOp.SUM, group=group)
    nv, ls = float(terms[0]), float(terms[1])
    if nv <= 0:
        return 0.0, float('nan')
    return 1.0 / nv, ls / nv


class SyncBnError(RuntimeError):
    """A rank did not reach the same BatchNorm layer within the timeout (raised on EVERY rank of the group)."""


class _DeviceArena:
    """Zero-copy torch view of a device buffer owned by libams_b200 (CUDA array interface)."""

    def __init__(self, ptr, count, typestr='<f4'):
        self.__cuda_array_interface__ = {'shape': (count,), 'typestr': typestr, 'data': (ptr, False), 'version': 2}


def exchange_ipc_handles(handle, group=None):
    """All-gather of the 64-byte CUDA IPC handles of the ranks' SyncBN receive buffers -> uint8 [world, 64]."""
    world = dist.get_world_size(group)
    mine = torch.from_numpy(handle.copy()).cuda()
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return torch.stack(out).cpu().numpy()


class DataParallelStudent:
    """One rank of a data-parallel distillation job: wraps a `Student` and the process group.

    The step is enqueued without any host round trip: forward/backward (CUDA graph), allreduce of the gradient arena
    and of the (n_valid, loss_sum) pair in place on the device, Adam reading 1 / n_valid from that pair.  The loss of
    step i lands in a page-locked slot and is read after a later synchronisation (the reference only prints it,
    SemanticNetwork.py:261).

    sync_bn=True: BatchNorm batch statistics are summed over the ranks inside the BN finalize kernels through NVLink
    peer memory (ams_syncbn_*), which makes the job equivalent to the reference's single-process step on the global
    batch; sync_bn=False keeps per-replica statistics (faster, a documented deviation)."""

    def __init__(self, student, group=None, sync_bn=False, strict=True, tensors=None, buckets=True):
        self.student = student
        self.group = group
        self.comm_stream = None
        self.split = 0
        if tensors is not None:
            # (gradient arena, terms) supplied by the caller: the gloo / CPU test of this class's host logic
            self.grad, self.terms = tensors
            if buckets and hasattr(student, 'gradient_bucket_split'):
                self.split = int(student.gradient_bucket_split())
        else:
            ptr, n = student.gradient_arena()
            self.grad = torch.as_tensor(_DeviceArena(ptr, n), device='cuda')
            self.terms = torch.as_tensor(_DeviceArena(student.step_terms_ptr(), 2, '<f8'), device='cuda')
            if buckets:
                self.split = int(student.gradient_bucket_split())
                self.comm_stream = torch.cuda.Stream()
        # two views of the arena: [0, split) early layers, [split, n) late layers (complete first in backward)
        self.grad_early = self.grad[:self.split] if self.split > 0 else None
        self.grad_late = self.grad[self.split:] if self.split > 0 else self.grad
        self._loss_slots = torch.zeros(256, dtype=torch.float32)
        if torch.cuda.is_available():
            self._loss_slots = self._loss_slots.pin_memory()
        self._loss_np = self._loss_slots.numpy()
        self._pending = 0
        self._backlog = []              # losses drained early because the slot ring was full
        self.sync_bn = False
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.sync_bn_error = None
        if sync_bn and self.world > 1:
            # strict=False: if peer memory cannot be mapped on ANY rank (no P2P / IPC in this container) all ranks agree
            # to fall back to per-replica statistics and say so in sync_bn_error
            rank = dist.get_rank(group)
            handles = exchange_ipc_handles(student.syncbn_init(self.world, rank), group)
            try:
                student.syncbn_connect(handles)
            except Exception as e:                       # noqa: BLE001 -- reported, and re-raised when strict
                self.sync_bn_error = str(e)
            okf = torch.tensor([0 if self.sync_bn_error else 1], device='cuda')
            dist.all_reduce(okf, op=dist.ReduceOp.MIN, group=group)     # also: every rank has mapped every zeroed buffer
            if int(okf.item()) == 1:
                self.sync_bn = True
            else:
                if not self.sync_bn_error:
                    student.syncbn_enable(False)
                    self.sync_bn_error = 'a peer rank could not map the receive buffers'
                if strict:
                    raise RuntimeError('SyncBN peer-memory setup failed: ' + self.sync_bn_error)

    def _slot(self):
        if self._pending >= self._loss_np.size:
            self._backlog = self.losses()            # ring full: synchronise once, keep what was read
        i = self._pending
        self._pending += 1
        return self._loss_np[i:i + 1]

    def train_step_async(self, lr, masked):
        """Enqueue one step; returns nothing.  losses() synchronises and returns the losses of the steps since the last call."""
        # the library and torch must share one stream (Student.set_stream(torch.cuda.current_stream().cuda_stream)):
        # the allreduces are ordered after backward and before Adam by stream order alone
        self.student.train_forward_backward_async()
        if self.world > 1:
            late = None
            if self.comm_stream is not None:
                # late-layer bucket: its allreduce waits (on the communication stream) for the event the library records
                # in the MIDDLE of the step, and overlaps the rest of the backward pass
                with torch.cuda.stream(self.comm_stream):
                    self.student.gradient_bucket_wait(self.comm_stream.cuda_stream)
                    late = dist.all_reduce(self.grad_late, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            else:
                dist.all_reduce(self.grad_late, op=dist.ReduceOp.SUM, group=self.group)
            if self.grad_early is not None:
                dist.all_reduce(self.grad_early, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.terms, op=dist.ReduceOp.SUM, group=self.group)
            if late is not None:
                late.wait()                              # the step's stream waits for the late bucket before Adam reads it
        self.student.apply_optimizer_device(lr, masked, self._slot())

    def losses(self):
        self.student.synchronize()
        out = self._backlog + [float(x) for x in self._loss_np[:self._pending]]
        self._pending = 0
        self._backlog = []
        if self.sync_bn and self.world > 1:
            _, err = self.student.syncbn_status()
            flag = torch.tensor([int(err)], dtype=torch.int32, device='cuda')
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)      # every rank learns about it: same control flow
            if int(flag.item()):
                raise SyncBnError('SyncBN exchange timed out: a rank did not reach the same BatchNorm layer (ranks must run '
                                  'the same sequence of training steps); the statistics of this phase are invalid')
        return out

    def disable_sync_bn(self, reason):
        """Fall back to per-replica statistics (call on every rank, e.g. after SyncBnError)."""
        self.student.syncbn_enable(False)
        self.sync_bn = False
        self.sync_bn_error = reason

    def train_step(self, lr, masked):
        """Synchronous form: returns this step's global mean loss."""
        self.train_step_async(lr, masked)
        return self.losses()[-1]

    def close(self):
        """Ranks leave together: nobody unmaps a receive buffer a peer may still push into."""
        if self.world > 1:
            self.student.synchronize()
            dist.barrier(group=self.group)
