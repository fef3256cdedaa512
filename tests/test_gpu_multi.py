"""Two-GPU test of the sharded driver (skipped on a one-GPU box): Calculator.run() under torch.distributed
with NCCL -- cells sharded round-robin, one packed all_gather on the devices -- must give every rank the
complete, ordered result, bit for bit equal to the single-rank solve of the same batch
(/root/reference/catint/calculator.py:209-212, catint_io.py:154-178 is what it replaces); the two-wave continuation
run over the same two ranks must end on the same steady states (1e-6)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_run_equals_single_rank_on_two_gpus():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('GPU tests need a B200')
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, CATINT_QUIET='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29631', os.path.join(ROOT, 'scripts', 'dist_check.py')]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:]
    assert p.stdout.count('sharded == single-rank: True') == 2, p.stdout[-3000:]
    assert p.stdout.count('continuation == plain: True') == 2, p.stdout[-3000:]
