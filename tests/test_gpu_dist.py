"""GPU, >= 2 devices: the N-sharded path over NCCL against the single-GPU path
(tests/dist_check.py under torchrun).  Skipped on a one-GPU box; the ring plumbing itself
is covered on CPU by tests/test_ring_cpu.py."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 4])
def test_sharded_equals_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(29540 + world),
           os.path.join(ROOT, 'tests', 'dist_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert 'DIST_CHECK_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
