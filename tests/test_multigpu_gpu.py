"""NCCL hit-table allgather across GPUs (skipped on single-GPU boxes)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_allgather_hits_two_ranks():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'tools', 'mgpu_search.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert 'MGPU OK' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
