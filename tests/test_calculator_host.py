"""Host logic of the Calculator (time mesh quirks, batch assembly, result scatter) -- no GPU."""
import numpy as np
import pytest

from conftest import load_golden
from catint_b200.transport import Transport
from catint_b200.calculator import Calculator, build_cell_batch, shard_indices
from catint_b200 import backend as be
from catint_b200 import workloads


def make_tp(resultsdir, **kw):
    tp = Transport(resultsdir=resultsdir, **kw)
    tp.set_calculator('odeint')
    return tp


def test_time_mesh_and_output_indices(resultsdir):
    """calculator.py:105-138: tmesh = arange(0,tmax+dt,dt); only the last index for ntout=1;
    `it>1 and it % int(nt/ntout) == 0` otherwise."""
    tp = make_tp(resultsdir, **workloads.c1())
    c = Calculator(transport=tp, dt=0.1, tmax=10, ntout=1)
    assert tp.nt == 101 and tp.itout == [100] and tp.ntout == 1
    c = Calculator(transport=tp, dt=0.1, tmax=10, ntout=5)
    assert tp.itout == [20, 40, 60, 80, 100]
    assert c.output_times() == pytest.approx([2., 4., 6., 8., 10.])
    c = Calculator(transport=tp)                      # no time mesh: arange(0,1,0.1)
    assert tp.nt == 10 and tp.itout == [9]


def test_calc_names(resultsdir):
    tp = make_tp(resultsdir, **workloads.c1())
    assert Calculator(transport=tp, dt=1, tmax=2, calc='lsoda').calc == 'lsoda'
    c = Calculator(transport=tp, dt=1, tmax=2, calc='vode--bdf')
    assert (c.calc, c.calc_method) == ('vode', 'bdf')
    assert Calculator(transport=tp, dt=1, tmax=2, calc='FTCS').calc == 'FTCS'        # K4 (tests/test_explicit.py)
    for bad in ('nonsense', 'comsol', 'dopri5', 'odeint--LF'):
        with pytest.raises(SystemExit):
            Calculator(transport=tp, dt=1, tmax=2, calc=bad)
    with pytest.raises(SystemExit):
        Calculator(transport=None)


def test_batch_of_a_potential_sweep(resultsdir):
    kw = workloads.c2(n_potentials=16)
    tp = make_tp(resultsdir, **kw)
    batch, models = build_cell_batch(tp)
    su = load_golden('ref_c1.npz')
    S = 8
    assert batch.B == 16 and batch.S == S and batch.R == 5 and batch.nx_max == 101 and batch.b == 9
    assert np.array_equal(batch.z, su['z'])
    assert np.all(batch.nx == 101)
    # everything but the fluxes and the wall potential is the C1 fixture
    for c in range(16):
        p = batch.par[c]
        assert np.array_equal(p[0:S], su['c_bulk']) and np.array_equal(p[2 * S:3 * S], su['D'])
        assert p[3 * S] == float(su['beta']) and p[3 * S + 1] == float(su['eps']) and p[3 * S + 5] == float(su['dx'])
        phi = tp.descriptors['phiM'][c]
        assert p[3 * S + 2] == phi
        i_co = min(10. * 10 ** (-(phi + 0.9) / 0.12), 150.)
        assert p[S + 3] == pytest.approx(i_co / 2 / 96485.33289, rel=1e-13)       # CO
        assert p[S + 1] == pytest.approx(-p[S + 3], rel=1e-13)                       # CO2 = -CO
    assert np.array_equal(batch.nu, be.stoichiometry(S, batch.reactions, 'summed'))
    leg, _ = build_cell_batch(tp, rate_mode='legacy_overwrite')
    assert not np.array_equal(leg.nu, batch.nu)


def test_batch_of_a_ragged_thickness_sweep(resultsdir):
    kw = workloads.co2r_inputs()
    kw['descriptors'] = {'boundary thickness': [30e-6, 50e-6], 'bulk_pH': [6.8, 7.5]}
    tp = make_tp(resultsdir, **kw)
    batch, models = build_cell_batch(tp)
    assert batch.B == 4 and list(batch.nx) == [102, 102, 101, 101] and batch.nx_max == 102
    a, b = load_golden('ref_c1_L30.npz'), load_golden('ref_c1_pH7p5.npz')
    assert batch.par[0, 3 * 8 + 5] == float(a['dx'])
    assert np.array_equal(batch.par[0, 0:8], a['c_bulk'])
    assert np.array_equal(batch.par[3, 0:8], b['c_bulk'])
    sub = batch.select(shard_indices(4, 1, 2))
    assert sub.B == 2 and list(sub.nx) == [102, 101] and sub.nx_max == 102


def test_scatter_results_conventions(resultsdir):
    """result containers as the reference's tools read them (comsol_reader.py:186-300, plot.py:51-56)."""
    kw = workloads.c2(n_potentials=3)
    tp = make_tp(resultsdir, **kw)
    calc = Calculator(transport=tp, dt=0.1, tmax=10, ntout=2)
    batch, models = build_cell_batch(tp)
    rng = np.random.default_rng(0)
    n_out, B, n, S = 2, 3, 101, 8
    res = {'c': rng.uniform(0.1, 2.0, (n_out, B, n, S)), 'g': rng.normal(size=(n_out, B, n)),
           'phi': rng.normal(size=(n_out, B, n)), 'flux': batch.par[:, S:2 * S].copy(),
           'status': np.zeros(B, dtype=np.int32), 'n_steps': np.ones(B, dtype=np.int32),
           'n_newton': np.ones(B, dtype=np.int32), 'n_setups': np.ones(B, dtype=np.int32)}
    calc.scatter_results(batch, models, res)
    names = list(tp.species)
    ad = tp.alldata[1]
    k = names.index('CO')
    assert ad['species']['CO']['surface_concentration'] == res['c'][-1, 1, 0, k]
    assert np.array_equal(ad['species']['CO']['concentration'], res['c'][-1, 1, :, k])
    j = res['flux'][1, k]
    assert ad['species']['CO']['electrode_flux'] == j
    assert ad['species']['CO']['electrode_current_density'] == pytest.approx(j * 2 * 96485.33289 / 1 / 10.)
    assert 'electrode_current_density' not in ad['species']['K+']
    h = names.index('H+')
    assert ad['system']['surface_pH'] == pytest.approx(-np.log10(res['c'][-1, 1, 0, h] / 1000.))
    assert np.array_equal(ad['system']['efield'], -res['g'][-1, 1])
    assert ad['system']['surface_potential'] == res['phi'][-1, 1, 0]
    q = np.array([tp.species[s]['charge'] for s in names]) * 96485.33289
    assert ad['system']['charge_density'][5] == pytest.approx(float(res['c'][-1, 1, 5] @ q))
    # serial-path containers hold the LAST cell; cout is species-major flat per output time
    assert len(tp.cout) == n_out and tp.cout[0].shape == (S * n,)
    assert np.array_equal(tp.cout[1][k * n:(k + 1) * n], res['c'][1, 2, :, k])
    assert np.array_equal(tp.efield, -res['g'][-1, 2]) and np.array_equal(tp.potential, res['phi'][-1, 2])


def test_get_rates_host_helper(resultsdir):
    from oracle.fixtures import system_from_setup
    tp = make_tp(resultsdir, **workloads.c1())
    su = load_golden('ref_c1.npz')
    C = su['rhs_states'][3].reshape(8, 101)
    for mode in ('summed', 'legacy_overwrite'):
        calc = Calculator(transport=tp, dt=1, tmax=2, rate_mode=mode)
        want = system_from_setup(su, mode).rates(C)
        assert np.allclose(calc.get_rates(C), want, rtol=1e-12, atol=1e-12 * np.max(np.abs(want)))


def test_run_without_gpu_fails_loudly(resultsdir):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    tp = make_tp(resultsdir, **workloads.c1())
    calc = Calculator(transport=tp, dt=0.1, tmax=1.0)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        calc.run()


def test_continuation_plan_splits_rows_and_finds_the_nearest_cold_cell():
    """Calculator(continuation=k): every k-th cell and the last one cold; on a 2D descriptor grid per row of the inner
    descriptor, neighbours never across a row end"""
    from catint_b200.calculator import continuation_plan
    cold, warm, near = continuation_plan(23, 4)
    assert list(cold) == [0, 4, 8, 12, 16, 20, 22] and len(warm) == 16
    for w, p in zip(warm, near):
        d = np.abs(cold - w)
        assert d[p] == d.min() and (p == 0 or d[p - 1] > d[p])           # nearest; ties go to the lower neighbour
    # 3 rows of 10 cells
    cold, warm, near = continuation_plan(30, 4, row_length=10)
    assert list(cold) == [0, 4, 8, 9, 10, 14, 18, 19, 20, 24, 28, 29]
    assert sorted(list(cold) + list(warm)) == list(range(30))
    for w, p in zip(warm, near):
        assert cold[p] // 10 == w // 10                                  # same row
        same_row = cold[cold // 10 == w // 10]
        assert abs(cold[p] - w) == np.min(np.abs(same_row - w))
    # degenerate inputs fall back to one row; k larger than the sweep: first and last cell cold
    assert list(continuation_plan(7, 4, row_length=3)[0]) == [0, 4, 6]
    cold, warm, near = continuation_plan(5, 16)
    assert list(cold) == [0, 4] and list(near) == [0, 0, 1]
    assert len(continuation_plan(2, 2)[1]) == 0


def test_auto_continuation_only_where_it_pays(resultsdir):
    from catint_b200 import workloads
    from catint_b200.calculator import auto_continuation_k, Calculator
    from catint_b200.transport import Transport
    assert auto_continuation_k(1024) is None and auto_continuation_k(4095) is None        # a plain run is as fast
    assert auto_continuation_k(16384) == 16 and auto_continuation_k(65536) == 64           # one resident cold wave
    assert auto_continuation_k(16384, world_size=8) is None and auto_continuation_k(131072, world_size=8) == 16
    assert auto_continuation_k(65536, row_length=256) is None                               # 2D grid: measured slower
    tp = Transport(resultsdir=resultsdir, **workloads.c1())
    tp.set_calculator('odeint')
    assert Calculator(transport=tp, dt=0.5, tmax=200, mode='stationary', continuation='auto').continuation == 'auto'
    for bad in ('always', 1):
        with pytest.raises(SystemExit):
            Calculator(transport=tp, dt=0.5, tmax=200, mode='stationary', continuation=bad)
    with pytest.raises(SystemExit):
        Calculator(transport=tp, dt=0.5, tmax=200, continuation='auto')                    # stationary mode only


def test_continuation_brackets_interpolation_weights():
    from catint_b200.calculator import continuation_plan, continuation_brackets
    cold, warm, _ = continuation_plan(30, 4, row_length=10)
    left, right, w = continuation_brackets(cold, warm, row_length=10)
    for cell, l, r, wt in zip(warm, left, right, w):
        assert cold[l] < cell < cold[r] and cold[l] // 10 == cell // 10 == cold[r] // 10
        assert not np.any((cold > cold[l]) & (cold < cold[r]))                  # adjacent cold cells
        assert wt == (cell - cold[l]) / float(cold[r] - cold[l]) and 0.0 < wt < 1.0
    cold, warm, _ = continuation_plan(23, 4)
    left, right, w = continuation_brackets(cold, warm)
    assert list(cold[left][-1:]) == [20] and list(cold[right][-1:]) == [22] and w[-1] == 0.5


def test_run_with_continuation_end_to_end_on_a_stand_in_solver(resultsdir):
    """Calculator.run() with continuation: descriptor grid -> plan -> waves -> result containers, with the device solve
    replaced by a stand-in (plumbing only; the product has no CPU solver)."""
    import torch
    tp = Transport(resultsdir=resultsdir, model_name='cont', **workloads.c2(n_potentials=12))
    tp.set_calculator('odeint')
    calc = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary', continuation=4)
    calls = []

    def stand_in(sub, backend=None, pinned=None, y0=None, max_steps=None):
        B, n, S = sub.B, sub.nx_max, sub.S
        calls.append((B, y0 is not None))
        c = torch.as_tensor(np.ascontiguousarray(np.broadcast_to(sub.par[:, None, 0:S], (B, n, S)))).clone()[None]
        if y0 is not None:
            assert tuple(y0.shape) == (B, n, S) and y0.dtype == torch.float64
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64)
        i32 = lambda v: torch.full((B,), v, dtype=torch.int32)
        return {'c': c, 'phi': z(1, B, n), 'g': z(1, B, n), 'flux': torch.as_tensor(sub.par[:, S:2 * S].copy()),
                'status': i32(0), 'n_steps': i32(7 if y0 is not None else 70), 'n_newton': i32(9), 'n_setups': i32(3),
                'h2d_bytes': 1}

    calc.solve_batch_device = stand_in
    res = calc.run()
    assert calls == [(4, False), (8, True)]                       # cells 0,4,8,11 cold; the other 8 warm
    assert calc.stats['converged'] == 12 and res['c'].shape == (1, 12, tp.nx, len(tp.species))
    assert list(res['n_steps']) == [70, 7, 7, 7, 70, 7, 7, 7, 70, 7, 7, 70]
    st = calc.continuation_stats
    assert st['cold_cells'] == 4 and st['warm_cells'] == 8 and st['interpolated_starts'] == 8 and st['rerun_cold_cells'] == 0
    assert len(tp.alldata) == 12 and 'surface_concentration' in tp.alldata[5]['species']['CO2']
    # 'auto' on a sweep this small is a plain run
    calc2 = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary', continuation='auto')
    calls.clear()
    calc2.solve_batch_device = stand_in
    calc2.solve_batch = lambda batch, **kw: {k: (v.numpy() if hasattr(v, 'numpy') else v) for k, v in stand_in(batch).items()}
    calc2.run()
    assert calls == [(12, False)]
