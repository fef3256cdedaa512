"""world_size-2 gloo test of the N>1 host path: round-robin sharding of the cell batch and the
single final all_gather.  The solve itself is replaced by a deterministic stand-in (this tests
plumbing; the product has no CPU solver)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden, batch_from_setup
from catint_b200 import distributed as D


def fake_solve(sub):
    """any deterministic function of the per-cell parameters, shaped like the real result"""
    B, n, S = sub.B, sub.nx_max, sub.S
    key = sub.par[:, S:2 * S].sum(axis=1)               # depends on the cell's fluxes
    c = np.zeros((2, B, n, S))
    c += key[None, :, None, None] * 1e6
    c += np.arange(n)[None, None, :, None] + 0.01 * np.arange(S)[None, None, None, :] + np.arange(2)[:, None, None, None] * 100
    return {'c': c, 'phi': c[..., 0] * 2, 'g': c[..., 1] * 3, 'flux': sub.par[:, S:2 * S].copy(),
            'status': (key > np.median(key)).astype(np.int32), 'n_steps': (np.arange(B) + 7).astype(np.int32) * 0 + sub.nx,
            'n_newton': sub.nx * 2, 'n_setups': sub.nx * 3}


def make_batch(B=7):
    su = load_golden('ref_c1.npz')
    rng = np.random.default_rng(3)
    fl = su['flux'][None, :] * rng.uniform(0.1, 3.0, (B, 1))
    batch = batch_from_setup(su, B=B, fluxes=fl)
    batch.nx[::3] = 100                                  # ragged
    return batch


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        batch = make_batch()
        full = D.solve_sharded(None, batch, solve_fn=fake_solve)
        q.put((rank, {k: v for k, v in full.items()}))
    finally:
        dist.destroy_process_group()


def test_shard_map_is_round_robin():
    assert list(D.shard_indices(7, 0, 2)) == [0, 2, 4, 6]
    assert list(D.shard_indices(7, 1, 2)) == [1, 3, 5]
    parts = np.concatenate([D.shard_indices(1024, r, 8) for r in range(8)])
    assert sorted(parts) == list(range(1024))


def test_two_rank_gather_restores_cell_order():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=60) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = fake_solve(make_batch())
    for r in range(2):
        for k in ('c', 'phi', 'g', 'flux', 'status', 'n_steps', 'n_newton', 'n_setups'):
            assert np.array_equal(got[r][k], want[k]), (r, k)


def test_single_process_passthrough():
    batch = make_batch(5)
    full = D.solve_sharded(None, batch, solve_fn=fake_solve)
    assert np.array_equal(full['c'], fake_solve(batch)['c'])
