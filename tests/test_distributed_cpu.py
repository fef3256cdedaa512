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
            'status': (np.floor(key * 1e7) % 3).astype(np.int32), 'n_steps': (np.arange(B) + 7).astype(np.int32) * 0 + sub.nx,
            'n_newton': sub.nx * 2, 'n_setups': sub.nx * 3}


def make_batch(B=7):
    su = load_golden('ref_c1.npz')
    rng = np.random.default_rng(3)
    fl = su['flux'][None, :] * rng.uniform(0.1, 3.0, (B, 1))
    batch = batch_from_setup(su, B=B, fluxes=fl)
    batch.nx[::3] = 100                                  # ragged
    return batch


def _worker(rank, world, port, q, n_cells=7, root_only=False):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        batch = make_batch(n_cells)
        calls = []

        def solve(sub):
            calls.append(sub.B)
            return fake_solve(sub)
        full = D.solve_sharded(None, batch, solve_fn=solve, n_out=2, root_only=root_only)
        if full is None:
            q.put((rank, None))
        else:
            full['solve_calls'] = np.array(calls)
            q.put((rank, {k: v for k, v in full.items()}))
    finally:
        dist.destroy_process_group()


def _run_ranks(world, n_cells, root_only=False):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, n_cells, root_only)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def test_shard_map_is_round_robin():
    assert list(D.shard_indices(7, 0, 2)) == [0, 2, 4, 6]
    assert list(D.shard_indices(7, 1, 2)) == [1, 3, 5]
    parts = np.concatenate([D.shard_indices(1024, r, 8) for r in range(8)])
    assert sorted(parts) == list(range(1024))


def test_two_rank_gather_restores_cell_order():
    got = _run_ranks(2, 7)
    want = fake_solve(make_batch())
    for r in range(2):
        for k in ('c', 'phi', 'g', 'flux', 'status', 'n_steps', 'n_newton', 'n_setups'):
            assert np.array_equal(got[r][k], want[k]), (r, k)
            assert got[r][k].dtype == want[k].dtype, (r, k)
        # ONE packed collective: [2 ranks x 4 cells, width] doubles
        n_out, n, S = 2, 101, 8
        assert got[r]['gather_bytes'] == 2 * 4 * (n_out * n * S + 2 * n_out * n + S + 4) * 8


def test_fewer_cells_than_ranks():
    """a rank whose round-robin shard is empty skips the solve and still joins the gather (the reference's
    `itask % size != rank: continue`, calculator.py:209-212)"""
    got = _run_ranks(3, 2)
    want = fake_solve(make_batch(2))
    for r in range(3):
        for k in ('c', 'phi', 'g', 'flux', 'status', 'n_steps', 'n_newton', 'n_setups'):
            assert np.array_equal(got[r][k], want[k]), (r, k)
    assert list(got[0]['solve_calls']) == [1] and list(got[1]['solve_calls']) == [1]
    assert list(got[2]['solve_calls']) == []


def test_single_process_passthrough():
    batch = make_batch(5)
    full = D.solve_sharded(None, batch, solve_fn=fake_solve, n_out=2)
    assert np.array_equal(full['c'], fake_solve(batch)['c'])


def _transport_worker(rank, world, port, q, resultsdir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['CATINT_QUIET'] = '1'
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from catint_b200 import workloads
        from catint_b200.transport import Transport
        tp = Transport(resultsdir=resultsdir, model_name='shared', **workloads.c2(n_potentials=3))
        tp.save()
        dist.barrier()
        q.put((rank, (tp.mpi_rank, tp.mpi_size, tp.outputfoldername, sorted(os.listdir(tp.outputfoldername)))))
    finally:
        dist.destroy_process_group()


def test_results_folder_is_created_once_and_saved_by_rank_zero(tmp_path):
    """every rank used to create its own numbered folder and pickle into the same files.  Now rank 0
    creates the folder and broadcasts its name, log files are per rank, only rank 0 saves."""
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_transport_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][:2] == (0, 2) and got[1][:2] == (1, 2)
    assert got[0][2] == got[1][2]
    assert sorted(os.listdir(str(tmp_path))) == ['shared_results']
    files = got[1][3]
    assert 'transport_id000.log' in files and 'transport_id001.log' in files
    assert sum(1 for f in files if f.startswith('alldata')) == 1


def test_root_only_gather_hands_the_results_to_rank_zero():
    """root_only=True: every rank joins the gather, only rank 0 copies the complete result to its host"""
    got = _run_ranks(2, 7, root_only=True)
    want = fake_solve(make_batch())
    assert got[1] is None
    for k in ('c', 'phi', 'g', 'flux', 'status', 'n_steps', 'n_newton', 'n_setups'):
        assert np.array_equal(got[0][k], want[k]), k


# ---- continuation (two waves) over two ranks: plumbing with a stand-in solve ---------------------------------
def _fake_device_solve(sub, y0=None, max_steps=None):
    """stand-in for Calculator.solve_batch_device (CPU tensors): the 'steady state' is a function of the cell's
    fluxes; the result also records where the cell started from (c of output 0 = y0, bulk state when none)"""
    r = fake_solve(sub)
    B, n, S = sub.B, sub.nx_max, sub.S
    r['c'] = r['c'][:2].copy()
    start = np.broadcast_to(sub.par[:, None, 0:S], (B, n, S)) if y0 is None else y0.numpy()
    r['c'][0] = start
    r['status'] = np.zeros(B, dtype=np.int32)
    r['status'][sub.par[:, 3 * S + 4] == -1.0] = 2       # cells flagged by a Stern capacitance of -1 "fail"
    if y0 is not None:                                   # ... flagged -2: fail when warm-started, converge from the bulk state
        r['status'][sub.par[:, 3 * S + 4] == -2.0] = 3
        assert max_steps == max(1000, 3 * int(sub.nx_max))          # cap = 3 x the slowest cold cell (fake n_steps = nx)
    return {k: torch.as_tensor(np.ascontiguousarray(v)) for k, v in r.items()}


def _continuation_worker(rank, world, port, q, n_cells, k, root_only):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from catint_b200.calculator import Calculator
        batch = make_batch(n_cells)
        S = batch.S
        batch.par[k, 3 * S + 4] = -1.0                   # the second cold cell fails: its neighbours start cold
        batch.par[10, 3 * S + 4] = -2.0                  # warm cell 10 fails from its neighbour's state: solved again cold
        batch.par[18, 0] *= 1.5                          # warm cell 18 has another bulk composition than its neighbour
        calc = Calculator.__new__(Calculator)            # plumbing only: no Transport, no device
        calc.continuation = k
        calc.max_steps = 100000
        calc.solve_batch_device = _fake_device_solve
        calc.output_times = lambda: [1.0, 2.0]
        res = calc.run_continuation(batch, root_only=root_only)
        q.put((rank, None if res is None else dict(res, stats=calc.continuation_stats)))
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.parametrize('world,root_only', [(1, False), (2, False), (2, True)])
def test_continuation_waves_start_from_the_nearest_cold_neighbour(world, root_only):
    n_cells, k = 23, 4
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_continuation_worker, args=(r, world, port, q, n_cells, k, root_only)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = make_batch(n_cells)
    S, n = batch.S, batch.nx_max
    batch.par[18, 0] *= 1.5
    want = fake_solve(batch)
    cold = sorted(set(range(0, n_cells, k)) | {n_cells - 1})
    for rank in range(world):
        res = got[rank]
        if root_only and rank != 0:
            assert res is None
            continue
        assert res['stats']['cold_cells'] == len(cold) and res['stats']['warm_cells'] == n_cells - len(cold)
        # final states and per-cell fields: those of a plain run, in cell order
        assert np.array_equal(res['c'][1], want['c'][1]) and np.array_equal(res['flux'], want['flux'])
        assert np.array_equal(res['n_steps'], want['n_steps']) and res['status'].dtype == np.int32
        assert res['stats']['rerun_cold_cells'] == 1 and res['status'][10] == 0 and res['status'][k] == 2
        assert np.count_nonzero(res['status']) == 1
        for cell in range(n_cells):
            start = res['c'][0, cell]
            bulk = np.broadcast_to(batch.par[cell, None, 0:S], (n, S))
            if cell in cold:
                assert np.array_equal(start, bulk)
                continue
            near = min(cold, key=lambda c: (abs(c - cell), c))
            left, right = max(c for c in cold if c < cell), min(c for c in cold if c > cell)
            nxc = batch.nx
            both = k not in (left, right) and nxc[left] == nxc[cell] == nxc[right]
            if cell in (10, 18):                         # re-run after a failed warm start / other bulk composition
                assert np.array_equal(start, bulk)
            elif both:                                   # both bracketing cold cells converged, same node count
                w = (cell - left) / float(right - left)
                mix = want['c'][1, left] + (want['c'][1, right] - want['c'][1, left]) * w
                assert np.array_equal(start, mix), (cell, left, right)
            elif near == k or nxc[near] != nxc[cell]:    # failed nearest neighbour / other node count -> bulk state
                assert np.array_equal(start, bulk)
            else:                                        # only the nearest one fits: its state
                assert np.array_equal(start, want['c'][1, near]), (cell, near)
        assert res['stats']['interpolated_starts'] == sum(
            1 for cell in range(n_cells) if cell not in cold and cell != 18
            and k not in (max(c for c in cold if c < cell), min(c for c in cold if c > cell))
            and batch.nx[max(c for c in cold if c < cell)] == batch.nx[cell] == batch.nx[min(c for c in cold if c > cell)])
        assert 0 < res['stats']['interpolated_starts'] < res['stats']['warm_started_cells']
