"""The reference's Poisson routine for every pb_bound pair and its fixed-step steppers:
/root/reference/catint/calculator_old.py:680-819 (get_potential_and_gradient), :976-1029 (integrate_FTCS),
:457-564 (integrate_Crank_Nicolson), dispatch :1121-1140.

CPU part: the literal restatement (oracle/explicit_oracle.py) against the oracle that is pinned by the
reference-executed ode_func fixtures (default pair), discrete identities for the other pairs, host plumbing.
GPU part (-m gpu): catint_pnp_potential_batch and catint_pnp_step_batch (K4) against the restatement.
Tolerances: the device uses prefix sums / a Thomas sweep where the reference has sequential loops / a dense solve:
same arithmetic up to summation order -> 1e-11 relative to the largest entry."""
import numpy as np
import pytest

from conftest import load_golden, batch_from_setup

UNIT_F = 96485.33289
PAIRS = {
    0: dict(potential=dict(wall=-0.07, bulk=None), gradient=dict(wall=None, bulk=4.0e5)),   # a field well above the
    #                                      rounding floor (~1e-4 V/m) of the cancelling charge sum
    2: dict(potential=dict(wall=-0.07, bulk=0.01), gradient=dict(wall=None, bulk=None)),
    3: dict(potential=dict(wall=-0.07, bulk=None), gradient=dict(wall=3.0e4, bulk=None)),
    4: dict(potential=dict(wall=None, bulk=0.01), gradient=dict(wall=3.0e4, bulk=None)),
    5: dict(potential=dict(wall=None, bulk=0.01), gradient=dict(wall=None, bulk=-2.0e3)),
}


def model_from_setup(su, pb_bound, rate_mode='summed', with_reactions=True):
    from catint_b200 import backend as be
    from oracle.explicit_oracle import ExplicitModel
    from oracle.fixtures import parse_rx
    S = len(su['z'])
    rx = parse_rx(su) if with_reactions else []
    return ExplicitModel(su['z'], su['D'], su['c_bulk'], su['flux'], float(su['dx']), int(su['nx']), float(su['beta']),
                         float(su['eps']), pb_bound, reactions=rx, nu=be.stoichiometry(S, rx, rate_mode),
                         use_migration=bool(su['use_migration']))


def batch_with_pair(su, code, B=1):
    batch = batch_from_setup(su, B=B)
    S = batch.S
    pb = PAIRS[code]
    f = lambda x: 0.0 if x is None else x
    batch.poisson_bc = code
    batch.par[:, 3 * S + 2] = f(pb['potential']['wall'])
    batch.par[:, 3 * S + 6] = f(pb['potential']['bulk'])
    batch.par[:, 3 * S + 7] = f(pb['gradient']['wall'])
    batch.par[:, 3 * S + 3] = f(pb['gradient']['bulk'])
    return batch


# ---------------------------------------------------------------- host ------------------------------------
def test_restated_poisson_routine_default_pair_matches_pinned_oracle():
    """wall potential + bulk gradient: the literal loops == the oracle whose RHS reproduces the reference's executed
    ode_func bit for bit (tests/test_oracle.py)"""
    from oracle.explicit_oracle import potential_and_gradient
    from oracle.fixtures import system_from_setup
    su = load_golden('ref_c1.npz')
    s = system_from_setup(su, 'summed', False)
    S, n = len(su['z']), int(su['nx'])
    pb = dict(potential=dict(wall=float(su['phi_wall']), bulk=None), gradient=dict(wall=None, bulk=float(su['g_bulk'])))
    for st in su['rhs_states']:
        C = st.reshape(S, n)
        v, g, lp = potential_and_gradient(C, float(su['dx']), s.q, float(su['eps']), pb)
        _, v2, g2, lp2 = s.rhs(st, with_field=True)
        assert np.max(np.abs(g - g2)) <= 1e-12 * max(np.max(np.abs(g2)), 1e-300)
        assert np.max(np.abs(v - v2)) <= 1e-12 * max(np.max(np.abs(v2)), 1e-300)
        assert np.max(np.abs(lp - lp2)) <= 1e-13 * max(np.max(np.abs(lp2)), 1e-300)


@pytest.mark.parametrize('code', [2, 3, 4, 5])
def test_restated_poisson_routine_other_pairs_satisfy_their_equations(code):
    from oracle.explicit_oracle import potential_and_gradient
    su = load_golden('ref_c1.npz')
    S, n = len(su['z']), int(su['nx'])
    dx = float(su['dx'])
    C = su['rhs_states'][3].reshape(S, n)
    pb = PAIRS[code]
    v, g, lp = potential_and_gradient(C, dx, su['z'] * UNIT_F, float(su['eps']), pb)
    if code == 2:
        assert v[0] == pb['potential']['wall'] and v[-1] == pb['potential']['bulk']
        res = (v[2:] - 2 * v[1:-1] + v[:-2]) / dx ** 2 - lp[1:-1]
        assert np.max(np.abs(res)) <= 1e-9 * np.max(np.abs(lp))
        assert np.allclose(g[1:-1], (v[2:] - v[:-2]) / (2 * dx), rtol=1e-14, atol=0)
    else:
        if pb['gradient']['wall'] is not None:
            assert g[0] == pb['gradient']['wall'] and np.allclose(np.diff(g[:-1]), lp[1:-1] * dx, rtol=1e-9, atol=1e-9 * np.max(np.abs(lp * dx)))
        else:
            assert g[-1] == pb['gradient']['bulk'] and np.allclose(np.diff(g[1:]), lp[1:-1] * dx, rtol=1e-9, atol=1e-9 * np.max(np.abs(lp * dx)))
        if pb['potential']['wall'] is not None:
            assert v[0] == pb['potential']['wall'] and np.allclose(np.diff(v[:-1]), g[1:-1] * dx, rtol=1e-9, atol=1e-12)
        else:
            assert v[-1] == pb['potential']['bulk'] and np.allclose(np.diff(v[1:]), g[1:-1] * dx, rtol=1e-9, atol=1e-12)


def test_pb_bound_pairs_map_to_backend_codes():
    from catint_b200.calculator import poisson_bc_code
    for code, pb in PAIRS.items():
        got = poisson_bc_code(pb)
        assert got[0] == code
    assert poisson_bc_code(dict(potential=dict(wall=0.0, bulk=None), gradient=dict(wall=1.0, bulk=1.0))) is None
    assert poisson_bc_code(dict(potential=dict(wall=None, bulk=None), gradient=dict(wall=None, bulk=1.0))) is None


def test_calculator_accepts_the_fixed_step_calcs(resultsdir):
    from catint_b200 import workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import Calculator
    tp = Transport(resultsdir=resultsdir, **workloads.c1())
    for name, lf in (('FTCS', False), ('Crank-Nicolson', False), ('FTCS--LF', True), ('Crank-Nicolson--LF', True)):
        c = Calculator(transport=tp, dt=1e-11, tmax=1e-9, ntout=4, calc=name)
        assert c.calc == name.split('--')[0] and c.use_lax_friedrich == lf
    with pytest.raises(SystemExit):
        Calculator(transport=tp, dt=1e-11, tmax=1e-9, ntout=4, calc='odeint--LF')
    with pytest.raises(SystemExit):                       # never silently ignored
        Calculator(transport=tp, dt=1e-11, tmax=1e-9, ntout=4, calc='odeint', scale_pb_grid='log')


# ---------------------------------------------------------------- device ----------------------------------
@pytest.fixture(scope='module')
def bk():
    import torch
    from catint_b200 import backend as be
    if not torch.cuda.is_available():
        pytest.skip('GPU tests need a B200')
    return be.PnpBackend('cuda:0')


def to_dev(a):
    import torch
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')


@pytest.mark.gpu
@pytest.mark.parametrize('code', [0, 2, 3, 4, 5])
@pytest.mark.parametrize('name', ['c1', 'c1_L30', 'c4'])
def test_potential_kernel_matches_restatement(bk, name, code):
    from oracle.explicit_oracle import potential_and_gradient
    su = load_golden('ref_%s.npz' % name)
    S, n = len(su['z']), int(su['nx'])
    states = su['rhs_states']
    batch = batch_with_pair(su, code, B=len(states))
    db = bk.upload(batch)
    c = np.stack([st.reshape(S, n).T for st in states])
    v, g, lp = [a.cpu().numpy() for a in bk.potential(db, to_dev(c))]
    for k, st in enumerate(states):
        v2, g2, lp2 = potential_and_gradient(st.reshape(S, n), float(su['dx']), su['z'] * UNIT_F, float(su['eps']), PAIRS[code])
        # the charge density is a cancelling sum: bounds relative to the sum of the absolute terms
        rho_abs = (np.abs(su['z'] * UNIT_F)[:, None] * np.abs(st.reshape(S, n))).sum(axis=0) / float(su['eps'])
        assert np.max(np.abs(lp[k] - lp2)) <= 1e-14 * np.max(rho_abs)
        gs = np.sum(rho_abs) * float(su['dx']) + np.max(np.abs(g2))
        assert np.max(np.abs(g[k] - g2)) <= 1e-13 * gs * (n if code == 2 else 1), (k, code)
        vs = gs * float(su['dx']) * n + np.max(np.abs(v2))
        assert np.max(np.abs(v[k] - v2)) <= 1e-13 * vs * (n if code == 2 else 1), (k, code)


@pytest.mark.gpu
@pytest.mark.parametrize('stepper,lf', [('ftcs', False), ('ftcs', True), ('cn', False), ('cn', True)])
@pytest.mark.parametrize('name,code', [('c1', 0), ('c1', 2), ('c1_nomig', 0), ('c4', 5)])
def test_fixed_step_steppers_match_restatement(bk, name, code, stepper, lf):
    """K4 against the literal restatement: 6 steps of 2e-11 s from the bulk state, outputs after three of them.
    (These schemes are unstable on the reference's problem -- the restatement itself overflows within ~60 steps at
    any step size above 1e-10 s -- so parity is checked where both sides are finite; errors are relative to the
    largest concentration, which carries the growth.)"""
    from catint_b200 import backend as be
    from oracle import explicit_oracle as eo
    su = load_golden('ref_%s.npz' % name)
    S, n = len(su['z']), int(su['nx'])
    batch = batch_with_pair(su, code, B=3)
    # three cells with different fluxes
    batch.par[1, S:2 * S] *= 0.3
    batch.par[2, S:2 * S] *= 2.0
    db = bk.upload(batch)
    dt = 2e-11
    nt, itout = 6, [1, 3, 5]
    out = bk.step(db, be.STEPPER_FTCS if stepper == 'ftcs' else be.STEPPER_CRANK_NICOLSON, dt, nt, itout, lax_friedrich=lf)
    got_c, got_v, got_g = out['c'].cpu().numpy(), out['phi'].cpu().numpy(), out['g'].cpu().numpy()
    for cell in range(3):
        m = model_from_setup(su, PAIRS[code])
        m.flux = batch.par[cell, S:2 * S].copy()
        fn = eo.integrate_ftcs if stepper == 'ftcs' else eo.integrate_crank_nicolson
        ref = fn(m, dt, nt, itout, lax_friedrich=lf)
        assert len(ref) == 3
        for k, (C, v, g) in enumerate(ref):
            assert np.all(np.isfinite(C))
            cs = np.max(np.abs(C))
            # With migration the field is a cancelling sum (terms ~1e12 V/m, rounding ~1e-4 V/m whatever the
            # summation order) and enters the Robin wall value with a factor mu/D = 39 per volt: the two summation
            # orders differ by ~1e-8 relative there, and the scheme amplifies it from step to step.
            tol = 1e-6 if bool(su['use_migration']) else 1e-11
            assert np.max(np.abs(got_c[k, cell, :n].T - C)) <= tol * cs, (cell, k)
            if bool(su['use_migration']):
                gs = np.sum((np.abs(su['z'] * UNIT_F)[:, None] * np.abs(C)).sum(axis=0)) / float(su['eps']) * float(su['dx'])
                assert np.max(np.abs(got_g[k, cell, :n] - g)) <= 1e-12 * gs * n + 1e-6 * np.max(np.abs(g)), (cell, k)
                assert np.max(np.abs(got_v[k, cell, :n] - v)) <= (1e-12 * gs * n + 1e-6 * np.max(np.abs(g))) * float(su['dx']) * n, (cell, k)


@pytest.mark.gpu
def test_calculator_runs_the_fixed_step_calcs(bk, resultsdir):
    """Calculator(calc='FTCS') end to end: same containers as the implicit path, state == the restatement"""
    from catint_b200 import workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import Calculator, build_cell_batch
    from oracle import explicit_oracle as eo
    kw = workloads.c2(n_potentials=3, phi_min=-0.8, phi_max=-1.0)
    tp = Transport(resultsdir=resultsdir, **kw)
    calc = Calculator(transport=tp, dt=2e-11, tmax=2.4e-10, ntout=2, calc='FTCS')
    res = calc.run()
    assert calc.stats['converged'] == 3
    batch, models = build_cell_batch(tp)
    S, n = batch.S, int(batch.nx[0])
    m0 = models[2]
    pb = m0.pb_bound
    model = eo.ExplicitModel(batch.z, batch.par[2, 2 * S:3 * S], batch.par[2, :S], batch.par[2, S:2 * S],
                             batch.par[2, 3 * S + 5], n, batch.par[2, 3 * S], batch.par[2, 3 * S + 1], pb,
                             reactions=batch.reactions, nu=batch.nu, use_migration=batch.use_migration)
    ref = eo.integrate_ftcs(model, tp.dt, tp.nt, [tp.nt - 1])
    got = res['c'][-1, 2, :n].T
    assert np.max(np.abs(got - ref[-1][0])) <= 1e-6 * np.max(np.abs(ref[-1][0]))
    assert len(tp.cout) == len([i for i in tp.itout if i < tp.nt])


@pytest.mark.gpu
def test_entry_points_reject_what_they_do_not_handle(bk):
    """error codes, not exits or silent fallbacks: the implicit kernels take the default pb_bound pair (and Stern),
    the fixed-step steppers every pair but Stern and uniform meshes only, bad control values are EINVAL"""
    import torch
    from catint_b200 import backend as be
    su = load_golden('ref_c1.npz')
    S, n = len(su['z']), int(su['nx'])
    c = to_dev(np.broadcast_to(su['c_bulk'][None, None, :], (1, n, S)))
    batch = batch_with_pair(su, be.BC_DIRICHLET_BOTH)
    db = bk.upload(batch)
    with pytest.raises(RuntimeError, match='default Poisson boundary'):
        bk.rhs(db, c)
    with pytest.raises(RuntimeError, match='Poisson boundary must be the default pair, the bulk/bulk pair or Stern'):
        bk.solve(db, [1.0])
    v, g, lp = bk.potential(db, c)                       # ... but the Poisson routine and the steppers take it
    assert torch.isfinite(v).all()
    out = bk.step(db, be.STEPPER_FTCS, 1e-11, 3, [2])
    assert out['c'].shape == (1, 1, n, S) and torch.isfinite(out['c']).all()
    with pytest.raises(RuntimeError, match='dt > 0'):
        bk.step(db, be.STEPPER_FTCS, 0.0, 3, [2])
    with pytest.raises(RuntimeError, match='unknown stepper'):
        bk.step(db, 7, 1e-11, 3, [2])
    stern = batch_with_pair(su, be.BC_DIRICHLET_BOTH)
    stern.poisson_bc = be.BC_STERN_ROBIN
    with pytest.raises(RuntimeError, match='not Stern'):
        bk.step(bk.upload(stern), be.STEPPER_FTCS, 1e-11, 3, [2])
    graded = batch_with_pair(su, 0)
    graded.mesh_id = np.zeros(1, dtype=np.int32)
    graded.mesh_xi = np.linspace(0.0, 1.0, n)[None, :]
    with pytest.raises(RuntimeError, match='uniform mesh'):
        bk.step(bk.upload(graded), be.STEPPER_CRANK_NICOLSON, 1e-11, 3, [2])
    # outputs only at the listed steps; Crank-Nicolson counts its steps from 1 like the reference
    out = bk.step(bk.upload(batch_with_pair(su, 0)), be.STEPPER_CRANK_NICOLSON, 1e-11, 4, [1, 3])
    assert out['c'].shape[0] == 2 and not torch.equal(out['c'][0], out['c'][1])


@pytest.mark.gpu
@pytest.mark.parametrize('graded', [False, True])
def test_implicit_integrator_takes_the_bulk_bulk_pair(bk, graded):
    """pb_bound = potential and gradient in the bulk (calculator_old.py:787-790 + :795-797): the field is the default
    pair's backward sum, so concentrations and gradient of catint_pnp_solve_batch must equal the default-pair run bit
    for bit; the potential must be the restatement's backward sum of that gradient (on the graded mesh: the same
    recursion with h_i = x_{i+1}-x_i)."""
    from catint_b200 import backend as be
    from oracle.explicit_oracle import potential_and_gradient
    su = load_golden('ref_c1.npz')
    S, n = len(su['z']), int(su['nx'])
    pair = dict(potential=dict(wall=None, bulk=0.013), gradient=dict(wall=None, bulk=float(su['g_bulk'])))
    res = {}
    for code in (be.BC_DIRICHLET_WALL_NEUMANN_BULK, be.BC_DIRICHLET_BULK_NEUMANN_BULK):
        batch = batch_from_setup(su, B=2)
        batch.par[1, S:2 * S] *= 0.5
        batch.poisson_bc = code
        batch.par[:, 3 * S + 3] = pair['gradient']['bulk']
        if code == be.BC_DIRICHLET_BULK_NEUMANN_BULK:
            batch.par[:, 3 * S + 2] = 0.0
            batch.par[:, 3 * S + 6] = pair['potential']['bulk']
        if graded:
            sx = np.linspace(0.0, 1.0, n)
            xi = sx + 0.02 * np.sin(np.pi * sx) ** 2          # spacings within +-6 % of the uniform ones
            batch.mesh_id = np.zeros(2, dtype=np.int32)
            batch.mesh_xi = xi[None, :]
            batch.par[:, 3 * S + 5] = float(su['dx']) * (n - 1)
        out = bk.solve(bk.upload(batch), [1e-7, 1e-6] if graded else [1e-3, 10.0], mode=be.MODE_TRANSIENT)
        assert out['status'].tolist() == [0, 0]
        res[code] = {k: out[k].cpu().numpy() for k in ('c', 'g', 'phi')}
    a, b = res[be.BC_DIRICHLET_WALL_NEUMANN_BULK], res[be.BC_DIRICHLET_BULK_NEUMANN_BULK]
    assert np.array_equal(a['c'], b['c']) and np.array_equal(a['g'], b['g'])
    for k in range(2):
        for cell in range(2):
            g = b['g'][k, cell, :n]
            v = b['phi'][k, cell, :n]
            if not graded:
                C = b['c'][k, cell, :n].T
                v2, g2, _ = potential_and_gradient(C, float(su['dx']), su['z'] * UNIT_F, float(su['eps']), pair)
                scale = np.max(np.abs(g2)) * float(su['dx']) * n + np.max(np.abs(v2))
                # the integrator carries g as an unknown solved to the Newton tolerance, the restatement sums the charge
                assert np.max(np.abs(g - g2)) <= 1e-6 * np.max(np.abs(g2))
                assert np.max(np.abs(v - v2)) <= 1e-6 * scale
            x = (batch.mesh_xi[0] * batch.par[cell, 3 * S + 5]) if graded else np.arange(n) * float(su['dx'])
            want = np.zeros(n)
            want[n - 1] = pair['potential']['bulk']
            for i in range(n - 2, 0, -1):
                want[i] = want[i + 1] - g[i] * (x[i + 1] - x[i])
            want[0] = want[1] + (want[1] - want[2]) * (x[1] - x[0]) / (x[2] - x[1])
            assert np.max(np.abs(v - want)) <= 1e-13 * (np.max(np.abs(want)) + np.sum(np.abs(g[1:n - 1] * np.diff(x)[1:])))
    assert not np.array_equal(a['phi'], b['phi'])
