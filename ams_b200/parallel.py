"""Multi-GPU plumbing for the two places where the AMS hot path shards.  The reference is single-GPU
(`visible_device_list`, SemanticNetwork.py:74); this is the B200 8-GPU extension of the same step.

  * inference: streams are independent -> `shard_streams` assigns stream s to rank s % world, no collective;
  * distillation: data parallel, one process per GPU.  Every rank runs forward/backward on its 8 frames with the
    loss left as a SUM over its valid pixels, then ONE exchange step: allreduce(sum) of the flat fp32 gradient arena
    (2,113,043 floats = 8.45 MB, NCCL over NVLink) and of (n_valid, loss_sum); Adam then runs identically on every
    rank with gradients scaled by 1 / global n_valid, which reproduces the reference's `reduce_mean` over all valid
    pixels of the global batch (utils/graph_utils.py:408).  BatchNorm statistics stay per replica (documented
    deviation).
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


class _DeviceArena:
    """Zero-copy torch view of a device buffer owned by libams_b200 (CUDA array interface)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {'shape': (count,), 'typestr': '<f4', 'data': (ptr, False), 'version': 2}


class DataParallelStudent:
    """One rank of a data-parallel distillation job: wraps a `Student` and the process group."""

    def __init__(self, student, group=None):
        self.student = student
        self.group = group
        ptr, n = student.gradient_arena()
        self.grad = torch.as_tensor(_DeviceArena(ptr, n), device='cuda')

    def train_step(self, lr, masked):
        # the library and torch must share one stream (Student.set_stream(torch.cuda.current_stream().cuda_stream)):
        # the allreduce is ordered after backward and before Adam by stream order alone
        n_valid, loss_sum = self.student.train_forward_backward()
        scale, loss = allreduce_step_terms(self.grad, n_valid, loss_sum, self.group)
        self.student.apply_optimizer(lr, masked, scale)
        return loss
