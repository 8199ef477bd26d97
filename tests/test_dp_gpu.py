"""Data-parallel distillation on 2 GPUs of one box: asynchronous exchange step and global-batch BatchNorm
statistics over NVLink peer memory.  The work is done by tests/dp_worker.py, one process per GPU under
torch.distributed.run; skipped on a single-GPU box (run with `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from _util import log

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs on one box')
@pytest.mark.parametrize('tag', ['cityscapes', 'pascalvoc2012'])
def test_dp_syncbn_matches_single_process_global_batch(tag):
    world = 8 if torch.cuda.device_count() >= 8 else (4 if torch.cuda.device_count() >= 4 else 2)      # powers of two: exact doubling
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'dp_worker.py'), tag]
    env = dict(os.environ, AMS_SYNCBN_TIMEOUT_MS='5000')
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    for line in r.stdout.splitlines():
        if line.startswith('[dp]'):
            log(line)
    assert r.returncode == 0, r.stdout[-4000:]
    assert '[dp] OK' in r.stdout
