"""Parity of the CUDA path (through the C ABI) with the oracle.  Run on the GPU box:
    python -m pytest tests -m gpu

Tolerances (stated per north_star): K1/K2 are the same arithmetic up to summation order ->
1e-12 relative to the largest entry; converged profiles / potentials / fluxes 1e-6 relative
(atol floor 1e-12*max c_bulk for values that cross zero)."""
import numpy as np
import pytest

from conftest import load_golden, RHS_CASES, batch_from_setup, oracle_system_of_cell

pytestmark = pytest.mark.gpu

RTOL_PROFILE = 1e-6


@pytest.fixture(scope='module')
def bk():
    import torch
    from catint_b200 import backend as be
    if not torch.cuda.is_available():
        pytest.skip('GPU tests need a B200')
    return be.PnpBackend('cuda:0')


def relerr(got, want, cscale, floor=1e-12):
    return float(np.max(np.abs(got - want) / (np.abs(want) + floor * cscale)))


def to_dev(a):
    import torch
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')


# ---------------------------------------------------------------- K1 ----------------------
@pytest.mark.parametrize('name', RHS_CASES)
@pytest.mark.parametrize('mode,literal', [('summed', False), ('legacy_overwrite', True)])
def test_rhs_kernel_matches_oracle(bk, name, mode, literal):
    from oracle.fixtures import system_from_setup
    su = load_golden('ref_%s.npz' % name)
    S, n = len(su['z']), int(su['nx'])
    batch = batch_from_setup(su, B=len(su['rhs_states']), rate_mode=mode, literal_sign=literal)
    db = bk.upload(batch)
    s = system_from_setup(su, mode, literal)
    c = np.stack([st.reshape(S, n).T for st in su['rhs_states']])
    dcdt, g, phi = bk.rhs(db, to_dev(c))
    dcdt, g, phi = dcdt.cpu().numpy(), g.cpu().numpy(), phi.cpu().numpy()
    eps_m = np.finfo(float).eps
    for k, st in enumerate(su['rhs_states']):
        ref, v, gg, _ = s.rhs(st, with_field=True)
        ref = ref.reshape(S, n).T
        C = st.reshape(S, n)
        # "same arithmetic up to summation order": the charge density is a cancelling sum, so the bound is
        # relative to the sum of the absolute terms (the bulk state cancels to 1e-11 of them), not to the result
        rho_abs = (np.abs(s.q)[:, None] * np.abs(C)).sum(axis=0) / s.eps * s.dx
        g_tol = 64 * eps_m * np.concatenate([np.cumsum(rho_abs[::-1])[::-1][1:], [0.0]]) + 1e-13 * np.max(np.abs(gg))
        g_tol[0] = 3 * g_tol[1]
        if s.use_migration:
            assert np.all(np.abs(g[k] - gg) <= g_tol)
            v_tol = np.cumsum(g_tol) * s.dx + 1e-13 * max(np.max(np.abs(v)), 1e-300)
            assert np.all(np.abs(phi[k] - v) <= 2 * v_tol + 2 * v_tol[-2])
        mig = (s.D * np.abs(s.beta * s.q))[None, :] * np.abs(C.T) / s.dx            # |d(dc/dt)/dg| per node
        d_tol = 1e-12 * np.max(np.abs(ref)) + mig * (np.roll(g_tol, 1) + np.roll(g_tol, -1) + g_tol)[:, None]
        assert np.all(np.abs(dcdt[k] - ref) <= d_tol)
        if literal and mode == 'legacy_overwrite':
            # the reference's own ode_func on the same state
            r2 = su['rhs_ref'][k].reshape(S, n).T
            assert np.all(np.abs(dcdt[k] - r2) <= d_tol)


# ---------------------------------------------------------------- K2 ----------------------
@pytest.mark.parametrize('name', ['c1', 'c1_norx', 'c4', 'c1_L30', 'c1_nomig'])
def test_jacobian_blocks_match_oracle(bk, name):
    from oracle.fixtures import system_from_setup
    from oracle.pnp_local import LocalForm
    su = load_golden('ref_%s.npz' % name)
    S, n = len(su['z']), int(su['nx'])
    s = system_from_setup(su, 'summed')
    lf = LocalForm(s)
    batch = batch_from_setup(su, B=2)
    db = bk.upload(batch)
    for st in (su['rhs_states'][2], su['rhs_states'][4]):
        y = lf.y_from_c(st.reshape(S, n))
        F, L, Dg, U, E0 = lf.residual(y, blocks=True)
        Fg, Lg, Dgg, Ug = [a[1].cpu().numpy() for a in bk.jacobian(db, to_dev(np.stack([y, y])))]
        for got, want in ((Fg, F), (Dgg, Dg), (Ug, U), (Lg[1:], L[1:]), (Lg[0], E0)):
            assert np.max(np.abs(got - want)) <= 1e-12 * max(np.max(np.abs(want)), 1e-300)


# ---------------------------------------------------------------- K3 ----------------------
GOLD = [('c1', 'summed', False), ('c1', 'legacy_overwrite', True), ('c1_nomig', 'summed', False),
        ('c1_norx', 'summed', False), ('c1_pH7p5', 'summed', False)]


@pytest.mark.parametrize('name,mode,literal', GOLD)
def test_steady_state_matches_odeint_golden(bk, name, mode, literal):
    """integrate to t=200 s + Newton polish == Newton root of the oracle's odeint end state
    (and == the odeint end state itself) to 1e-6 relative; wall flux == imposed flux."""
    from catint_b200 import backend as be
    from oracle.fixtures import system_from_setup
    su = load_golden('ref_%s.npz' % name)
    go = load_golden('oracle_%s_%s.npz' % (name, mode))
    S, n = len(su['z']), int(su['nx'])
    batch = batch_from_setup(su, B=2, rate_mode=mode, literal_sign=literal)
    out = bk.solve(bk.upload(batch), [200.0], mode=be.MODE_STEADY)
    assert out['status'].tolist() == [0, 0]
    cs = np.max(np.abs(su['c_bulk']))
    got = out['c'][-1, 0].cpu().numpy()
    assert np.array_equal(got, out['c'][-1, 1].cpu().numpy())           # identical cells, identical bits
    assert relerr(got, go['newton_c'].reshape(S, n).T, cs) < RTOL_PROFILE
    assert relerr(got, go['c_end'].reshape(S, n).T, cs) < RTOL_PROFILE
    s = system_from_setup(su, mode, literal)
    if s.use_migration:
        gsc = np.max(np.abs(go['newton_grad']))
        assert np.max(np.abs(out['g'][-1, 0].cpu().numpy() - go['newton_grad'])) < RTOL_PROFILE * gsc
        psc = max(np.max(np.abs(go['newton_potential'])), 1e-300)
        assert np.max(np.abs(out['phi'][-1, 0].cpu().numpy() - go['newton_potential'])) < RTOL_PROFILE * psc
    J = batch.par[0, S:2 * S]
    fl = out['flux'][0].cpu().numpy()
    assert np.max(np.abs(fl - J)) < RTOL_PROFILE * np.max(np.abs(J))


@pytest.mark.parametrize('name,mode,literal', [('c1', 'summed', False), ('c1_norx', 'summed', False)])
def test_transient_outputs_match_odeint(bk, name, mode, literal):
    """time-dependent mode at matching rtol/atol (scipy defaults): state at t=10 s and t=200 s.
    Both integrators carry their own global error of order 100*tol, hence 1e-5 at t=10 s."""
    from catint_b200 import backend as be
    su = load_golden('ref_%s.npz' % name)
    go = load_golden('oracle_%s_%s.npz' % (name, mode))
    S, n = len(su['z']), int(su['nx'])
    batch = batch_from_setup(su, B=1, rate_mode=mode, literal_sign=literal)
    out = bk.solve(bk.upload(batch), [10.0, 200.0], mode=be.MODE_TRANSIENT)
    assert out['status'].tolist() == [0]
    cs = np.max(np.abs(su['c_bulk']))
    assert relerr(out['c'][0, 0].cpu().numpy(), go['c_t10'].reshape(S, n).T, cs, floor=1e-9) < 1e-5
    assert relerr(out['c'][1, 0].cpu().numpy(), go['c_end'].reshape(S, n).T, cs, floor=1e-9) < RTOL_PROFILE


def test_potential_sweep_cells_match_odeint(bk, resultsdir):
    """8 cells of the 1024-point C2 sweep (currents 0.005 ... 150 A/m^2) against their odeint goldens:
    the discrete steady problem has several roots, the integrator must land on odeint's."""
    from catint_b200 import backend as be, workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    go = load_golden('oracle_c2_sweep.npz')
    tp = Transport(resultsdir=resultsdir, **workloads.c2())
    batch, _ = build_cell_batch(tp)
    cells = [int(c) for c in go['cells']]
    assert np.array_equal(batch.par[cells], go['par'])
    sub = batch.select(cells)
    out = bk.solve(bk.upload(sub), [200.0], mode=be.MODE_STEADY)
    assert out['status'].tolist() == [0] * len(cells)
    S, n = sub.S, sub.nx_max
    cs = np.max(np.abs(sub.par[0, :S]))
    for k, c in enumerate(cells):
        got = out['c'][-1, k].cpu().numpy()
        assert relerr(got, go['newton_c_%d' % c].reshape(S, n).T, cs) < RTOL_PROFILE, c
        assert relerr(got, go['c_end_%d' % c].reshape(S, n).T, cs, floor=1e-9) < RTOL_PROFILE, c


def test_forty_more_sweep_cells_land_on_the_odeint_root(bk, resultsdir):
    """Root selection over the WHOLE sweep: 40 further cells of the 1024-point C2 sweep (every 32nd cell and, more
    densely, the band where the Tafel current saturates and the step counts scatter) against scipy odeint's
    end state and its Newton root (tests/golden/make_sweep_golden.py dense)."""
    from catint_b200 import backend as be, workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    go = load_golden('oracle_c2_sweep_dense.npz')
    tp = Transport(resultsdir=resultsdir, **workloads.c2())
    batch, _ = build_cell_batch(tp)
    cells = [int(c) for c in go['cells']]
    assert len(cells) == 40 and np.array_equal(batch.par[cells], go['par'])
    sub = batch.select(cells)
    out = bk.solve(bk.upload(sub), [200.0], mode=be.MODE_STEADY)
    assert out['status'].tolist() == [0] * len(cells)
    S, n = sub.S, sub.nx_max
    cs = np.max(np.abs(sub.par[0, :S]))
    worst = 0.0
    for k, c in enumerate(cells):
        got = out['c'][-1, k].cpu().numpy()
        e1 = relerr(got, go['newton_c_%d' % c].reshape(S, n).T, cs)
        e2 = relerr(got, go['c_end_%d' % c].reshape(S, n).T, cs, floor=1e-9)
        assert e1 < RTOL_PROFILE and e2 < RTOL_PROFILE, (c, e1, e2)
        worst = max(worst, e1)
        psc = max(np.max(np.abs(go['newton_potential_%d' % c])), 1e-300)
        assert np.max(np.abs(out['phi'][-1, k].cpu().numpy() - go['newton_potential_%d' % c])) < RTOL_PROFILE * psc, c
    print('worst relative deviation from the odeint roots over 40 cells: %.2e' % worst)


def test_ragged_and_mixed_batch(bk):
    """cells with 101 and 102 nodes, different bulk compositions and temperatures in one launch;
    every cell must equal its own single-cell solve bit for bit, padding must stay untouched."""
    from catint_b200 import backend as be
    a, b, c = load_golden('ref_c1.npz'), load_golden('ref_c1_L30.npz'), load_golden('ref_c1_T320.npz')
    parts = [batch_from_setup(x, B=1) for x in (a, b, c, b, a)]
    par = np.concatenate([p.par for p in parts])
    nx = np.concatenate([p.nx for p in parts])
    mixed = be.CellBatch(parts[0].z, parts[0].reactions, parts[0].nu, par, nx, nx_max=104)
    out = bk.solve(bk.upload(mixed), [50.0], mode=be.MODE_STEADY)
    assert out['status'].tolist() == [0] * 5
    cm = out['c'][-1].cpu().numpy()
    assert np.all(cm[:, 102:, :] == 0.0) and np.all(cm[0, 101:, :] == 0.0)
    for k, p in enumerate(parts):
        single = be.CellBatch(p.z, p.reactions, p.nu, p.par, p.nx, nx_max=104)
        o1 = bk.solve(bk.upload(single), [50.0], mode=be.MODE_STEADY)
        assert np.array_equal(o1['c'][-1, 0].cpu().numpy(), cm[k])
    assert not np.array_equal(cm[0, :101], cm[1, :101])


def test_steady_state_is_a_root_of_the_rhs_kernel(bk):
    """size-independent property at the full C2 size: every converged cell zeroes K1's dc/dt and
    carries exactly the imposed wall flux; CO2 depletion grows monotonically with the current."""
    import torch
    from catint_b200 import backend as be, workloads
    su = load_golden('ref_c1.npz')
    S = 8
    names = [str(s) for s in su['species']]
    phis = np.linspace(-0.5, -1.5, 1024)
    fl = np.zeros((1024, S))
    for k, phi in enumerate(phis):
        f = 10.0 ** (-(phi + 0.9) / 0.12)
        jco, jh2 = min(10 * f, 150.) / 2 / 96485.33289, min(5 * f, 150.) / 2 / 96485.33289
        fl[k, names.index('CO')] = jco; fl[k, names.index('H2')] = jh2
        fl[k, names.index('CO2')] = -jco; fl[k, names.index('OH-')] = 2 * jco + 2 * jh2
    batch = batch_from_setup(su, B=1024, fluxes=fl)
    db = bk.upload(batch)
    out = bk.solve(db, [200.0], mode=be.MODE_STEADY)
    assert int((out['status'] == 0).sum()) == 1024
    c = out['c'][-1].contiguous()
    dcdt, _, _ = bk.rhs(db, c)
    # scale: size of the individual stencil terms D_k*c_k/dx^2 that cancel in dc/dt (rounding floor ~1e-16 of it)
    D = torch.tensor(su['D'], device='cuda:0')
    scale = (c.abs().amax(dim=1) * D[None, :] / float(su['dx']) ** 2).amax(dim=1)
    ratio = dcdt.abs().amax(dim=(1, 2)) / scale
    # the oracle's own Newton root sits at 2e-11 of this scale (cancellation in the migration terms)
    assert float(ratio.max()) < 5e-10, float(ratio.max())
    flux = out['flux'].cpu().numpy()
    assert np.max(np.abs(flux - fl)) <= 1e-9 * np.max(np.abs(fl))
    co2 = c[:, 0, names.index('CO2')].cpu().numpy()
    assert np.all(np.diff(co2) <= 1e-9)
    assert int(out['n_steps'].max()) < 5000


@pytest.mark.parametrize('S', [1, 2, 3, 5])
def test_small_block_sizes(bk, S):
    """block sizes 2..6 (one kernel instantiation each): reaction-free electrolytes with 1..5 species on a
    ragged pair of cells; the converged state must be a root of K1's dc/dt, carry the imposed wall fluxes and
    agree with the CPU Newton root of the same discrete system (oracle/pnp_local.py)."""
    import torch
    from catint_b200 import backend as be
    from oracle.pnp_local import LocalForm
    su = load_golden('ref_c1.npz')
    zs = {1: [0.0], 2: [1.0, -1.0], 3: [1.0, -1.0, 0.0], 5: [1.0, -1.0, 0.0, 2.0, -1.0]}[S]
    cb = {1: [30.0], 2: [100.0, 100.0], 3: [100.0, 100.0, 30.0], 5: [100.0, 140.0, 30.0, 20.0, 0.0]}[S]
    cb = np.array(cb); z = np.array(zs)
    if S == 5:
        cb[4] = 0.0; cb[1] = 100.0 + 2 * 20.0          # electroneutral: 100 + 2*20 = 140
    J = np.zeros(S); J[0] = -2e-5                       # species 0 consumed at the wall
    if S >= 3: J[2] = +2e-5                             # neutral product released
    D = np.array(su['D'][:S], dtype=float)
    B = 2
    par = np.zeros((B, be.npar(S)))
    for c in range(B):
        par[c, 0:S] = cb; par[c, S:2 * S] = J * (1 + c); par[c, 2 * S:3 * S] = D
        par[c, 3 * S + 0] = su['beta']; par[c, 3 * S + 1] = su['eps']; par[c, 3 * S + 2] = su['phi_wall']
        par[c, 3 * S + 3] = su['g_bulk']; par[c, 3 * S + 4] = 0.2; par[c, 3 * S + 5] = su['dx']
    nx = np.array([41, 38], dtype=np.int32)
    mig = S > 1
    batch = be.CellBatch(z, [], be.stoichiometry(S, [], 'summed'), par, nx, use_migration=mig,
                         species=['s%d' % k for k in range(S)])
    assert batch.b == S + 1
    db = bk.upload(batch)
    out = bk.solve(db, [50.0], mode=be.MODE_STEADY)
    assert out['status'].tolist() == [0, 0]
    c = out['c'][-1].contiguous()
    dcdt, _, _ = bk.rhs(db, c)
    for k in range(B):
        n = int(nx[k])
        scale = float((c[k, :n].abs().amax(dim=0) * torch.tensor(D, device=c.device)).max()) / float(su['dx']) ** 2
        assert float(dcdt[k, :n].abs().max()) < 5e-10 * scale
        assert np.max(np.abs(out['flux'][k].cpu().numpy() - par[k, S:2 * S])) <= 1e-9 * np.max(np.abs(par[k, S:2 * S]))
        lf = LocalForm(oracle_system_of_cell(batch, k))
        yr, info = lf.solve_steady(y0=lf.y_from_c(c[k, :n].cpu().numpy().T), pure_newton=True)
        assert info['converged']
        assert relerr(c[k, :n].cpu().numpy().T, lf.unpack(yr)[0], np.max(cb)) < RTOL_PROFILE


def test_calculator_run_end_to_end(bk, resultsdir):
    """the reference-facing call: Transport -> set_calculator -> Calculator.run() on a small sweep."""
    from catint_b200 import workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import Calculator
    kw = workloads.c2(n_potentials=8, phi_min=-0.8, phi_max=-1.0)
    tp = Transport(resultsdir=resultsdir, model_name='sweep', **kw)
    tp.set_calculator('odeint')
    calc = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary')
    res = calc.run()
    assert calc.stats['converged'] == 8
    go = load_golden('oracle_c1_summed.npz')
    # phiM=-0.9 is not on this grid, but cell fluxes are monotone: compare conventions instead
    ad = tp.alldata[3]
    assert len(ad['species']['CO2']['concentration']) == 101
    assert ad['species']['CO']['electrode_current_density'] == pytest.approx(
        tp.derive_for(phiM=tp.alldata_names[3][0]).species['CO']['current density'] / (-10.), rel=1e-6)
    assert ad['system']['status'] == 'converged'
    assert len(tp.cout) == 1 and tp.cout[0].shape == (8 * 101,)
    assert np.array_equal(tp.cout[0][101:202], np.array(tp.alldata[7]['species']['CO2']['concentration']))
    import os
    assert os.path.isfile(os.path.join(tp.outputfoldername, 'alldata.pkl'))


def _batch_from_par(par, nx, template_batch):
    from catint_b200 import backend as be
    return be.CellBatch(template_batch.z, template_batch.reactions, template_batch.nu, par, nx,
                        use_migration=template_batch.use_migration)


def test_ten_species_ragged_sweep_cells(bk, resultsdir):
    """C4-type cells: 10 species (CH4 from a third electrode reaction, inert Cl-), bulk_pH x boundary
    thickness sweep (101/102 nodes), block size 11, against goldens from the REFERENCE'S integrator (scipy
    odeint, tests/golden/make_c4_golden.py): end state at t=200 s and the Newton root started from it.
    The three thin-layer cells (dx = 0.1 um) are cells where odeint itself leaves every physical range
    at t ~ 1.2e-4 s (the discrete ODE blows up in finite time; growth record in the golden file): the GPU
    must report a failure there, not invent a result."""
    from catint_b200 import backend as be, workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    go = load_golden('oracle_c4_cells.npz')
    tp = Transport(resultsdir=resultsdir, **workloads.c4(n_pH=4, n_L=4))
    batch, _ = build_cell_batch(tp)
    assert batch.S == 10 and batch.b == 11 and batch.B == 16
    cells = [int(c) for c in go['cells']]
    assert cells == list(range(16))
    assert np.array_equal(batch.par[cells], go['par']) and np.array_equal(batch.nx[cells], go['nx'])
    out = bk.solve(bk.upload(batch), [200.0], mode=be.MODE_STEADY, max_steps=20000)
    status = out['status'].cpu().numpy()
    odeint_failed = [c for c in cells if not bool(go['ok_%d' % c])]
    assert odeint_failed == [0, 4, 8]
    for c in cells:
        if c in odeint_failed:
            assert 'blow-up guard' in str(go['msg_%d' % c]) and float(go['fail_cmax_%d' % c]) > 1e3
            assert status[c] == be_status('step_underflow'), (c, status[c])
            continue
        assert status[c] == 0, c
        n = int(batch.nx[c])
        cs = np.max(np.abs(batch.par[c, :10]))
        got = out['c'][-1, c, :n].cpu().numpy()
        assert relerr(got, go['newton_c_%d' % c], cs) < RTOL_PROFILE, c
        # odeint's own end state: not yet steady to 1e-6 at t=200 s for the 200 um layers (diffusion time 40 s)
        assert relerr(got, go['c_end_%d' % c], cs, floor=1e-9) < (RTOL_PROFILE if n * batch.par[c, 35] < 1e-4 else 5e-6), c
        gsc = np.max(np.abs(go['g_%d' % c]))
        assert np.max(np.abs(out['g'][-1, c, :n].cpu().numpy() - go['g_%d' % c])) < RTOL_PROFILE * gsc
        psc = np.max(np.abs(go['phi_%d' % c]))
        assert np.max(np.abs(out['phi'][-1, c, :n].cpu().numpy() - go['phi_%d' % c])) < RTOL_PROFILE * psc


def be_status(name):
    from catint_b200 import backend as be
    return {v: k for k, v in be.CELL_STATUS.items()}[name]


def test_c4_sweep_4096_cells_properties(bk, resultsdir):
    """size-independent properties on a 64 x 64 C4-type sweep (10 species, block size 11, ragged 101/102
    nodes, several launch waves): every cell is either converged or carries a failure status (thin layers
    whose discrete ODE blows up in finite time, as in the CPU oracle); every converged cell zeroes K1's
    dc/dt and carries exactly the imposed wall fluxes."""
    import torch
    from catint_b200 import backend as be, workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    tp = Transport(resultsdir=resultsdir, **workloads.c4(n_pH=64, n_L=64))
    batch, _ = build_cell_batch(tp)
    assert batch.B == 4096 and batch.b == 11
    db = bk.upload(batch)
    out = bk.solve(db, [200.0], mode=be.MODE_STEADY, max_steps=20000)
    status = out['status'].cpu().numpy()
    ok = status == 0
    assert ok.sum() >= 0.6 * batch.B, int(ok.sum())
    assert set(np.unique(status[~ok])) <= {1, 2, 3, 4, 5, 6}
    c = out['c'][-1].contiguous()
    dcdt, _, _ = bk.rhs(db, c)
    S = batch.S
    D = torch.tensor(batch.par[:, 2 * S:3 * S], device=c.device)
    dx = torch.tensor(batch.par[:, 3 * S + 5], device=c.device)
    scale = (c.abs().amax(dim=1) * D / dx[:, None] ** 2).amax(dim=1)
    ratio = (dcdt.abs().amax(dim=(1, 2)) / scale).cpu().numpy()
    # (measured: median 3e-11, max 1.3e-8 of the size of the cancelling stencil terms)
    assert np.all(ratio[ok] < 1e-7), float(ratio[ok].max())
    assert float(np.median(ratio[ok])) < 1e-9
    flux = out['flux'].cpu().numpy()
    J = batch.par[:, S:2 * S]
    assert np.max(np.abs(flux[ok] - J[ok])) <= 1e-8 * np.max(np.abs(J))
    # the failures are the thinnest layers only
    dxs = batch.par[:, 3 * S + 5]
    assert dxs[~ok].max() < 2.5e-7 and dxs[ok].max() > 1.5e-6


def test_thousand_node_grid_uses_global_state(bk, resultsdir):
    """1001 nodes: the Newton iterate no longer fits in shared memory (workspace path).  Transient
    outputs against the CPU BDF oracle at the same rtol/atol, steady state against its Newton root."""
    from catint_b200 import backend as be, workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    go = load_golden('oracle_n1001.npz')
    tp = Transport(resultsdir=resultsdir, **workloads.co2r_inputs(nx=1000))
    batch, _ = build_cell_batch(tp)
    assert batch.nx_max == 1001 and np.array_equal(batch.par, go['par'])
    b3 = batch.select([0, 0, 0])
    db = bk.upload(b3)
    cs = np.max(np.abs(batch.par[0, :8]))
    tr = bk.solve(db, go['t_out'], mode=be.MODE_TRANSIENT)
    assert tr['status'].tolist() == [0, 0, 0]
    # Conditioning of the transient (tests/golden/make_n1001_golden.py): the golden holds a second CPU run whose
    # initial state was perturbed by 1e-13 relative.  Where the two runs agree to 1e-5 the GPU must match to
    # 1e-5 (t = 1e-3, 1e-2 and 200 s); inside the window where the grid-scale mode of the reference scheme grows
    # and saturates (t = 1 s: the 1e-13 perturbation is amplified to >1e-4, i.e. by >1e9, so integration errors
    # of order rtol = 1.5e-8 are amplified to O(0.1) in ANY integrator, odeint included) only a sanity bound holds.
    sens = go['sensitivity']
    assert sens[0] < 1e-5 and sens[1] < 1e-5 and sens[3] < 1e-9 and sens[2] > 1e-5
    for k in range(len(go['t_out'])):
        tol = 1e-5 if sens[k] < 1e-5 else 0.3
        assert relerr(tr['c'][k, 1].cpu().numpy(), go['bdf_c'][k], cs, floor=1e-9) < tol, k
    st = bk.solve(db, [200.0], mode=be.MODE_STEADY)
    assert st['status'].tolist() == [0, 0, 0]
    assert relerr(st['c'][-1, 2].cpu().numpy(), go['newton_c'], cs) < RTOL_PROFILE
    assert np.max(np.abs(st['phi'][-1, 0].cpu().numpy() - go['phi'])) < RTOL_PROFILE * np.max(np.abs(go['phi']))


@pytest.mark.parametrize('nn', [101, 201, 1001])
def test_stern_boundary_on_graded_mesh(bk, resultsdir, nn):
    """C3-type cells: Stern-layer (Robin) Poisson boundary, phi carried as an unknown (block size S+2),
    geometric mesh with a 0.05 nm first interval; the four phiM x bulk_pH corner cells (extension beyond the
    reference FD code; goldens: tests/golden/make_c3_golden.py).
      nn=101  against the REFERENCE'S integrator (scipy odeint on the restated RHS with the same extension):
              end state and its Newton root; the same file pins the CPU BDF against odeint (1e-8 / 1e-9);
      nn=201, 1001  dense odeint is out of reach; CPU BDF + Newton root (nn=1001 runs through the global-state
              kernel variant: the iterate does not fit in shared memory).
    Cell 2 (phiM=-1.5 V, pH 6) has no bounded solution: odeint (nn=101) and the CPU BDF (all nn) both leave
    every bounded range at t = 0.685..0.689 s; the GPU must report exactly that cell as failed."""
    from catint_b200 import backend as be, workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    go = load_golden('oracle_c3_cells_n%d.npz' % nn)
    tp = Transport(resultsdir=resultsdir, **workloads.c3(n_phi=2, n_pH=2))
    batch, _ = build_cell_batch(tp, poisson_bc='stern', mesh=workloads.geometric_mesh(nn, 5e-11))
    assert batch.b == 10 and batch.nx_max == nn
    assert np.array_equal(batch.par, go['par']) and np.allclose(batch.mesh_xi[0], go['mesh'], rtol=0, atol=0)
    out = bk.solve(bk.upload(batch), [200.0], mode=be.MODE_STEADY, max_steps=50000)
    status = out['status'].cpu().numpy()
    assert [c for c in range(4) if not bool(go['ok_%d' % c])] == [2]
    assert 0.68 < float(go['fail_t_2']) < 0.69
    if nn == 101:
        # the rung that pins the CPU BDF (and the failure) against odeint
        assert [c for c in range(4) if not bool(go['odeint_ok_%d' % c])] == [2]
        assert abs(float(go['odeint_fail_t_2']) - float(go['fail_t_2'])) < 1e-3
        for c in (0, 1, 3):
            assert float(go['bdf_vs_odeint_end_%d' % c]) < 1e-7 and float(go['bdf_vs_odeint_root_%d' % c]) < 1e-8
    for c in range(batch.B):
        if c == 2:
            assert status[c] == be_status('step_underflow'), (c, status)
            continue
        assert status[c] == 0, (c, status)
        cs = np.max(np.abs(batch.par[c, :8]))
        got = out['c'][-1, c].cpu().numpy()
        assert relerr(got, go['newton_c_%d' % c], cs) < RTOL_PROFILE, c
        assert np.max(np.abs(out['phi'][-1, c].cpu().numpy() - go['phi_%d' % c])) < RTOL_PROFILE * np.max(np.abs(go['phi_%d' % c]))
        assert np.max(np.abs(out['g'][-1, c].cpu().numpy() - go['g_%d' % c])) < RTOL_PROFILE * np.max(np.abs(go['g_%d' % c]))
        if nn == 101:
            assert relerr(got, go['odeint_newton_c_%d' % c], cs) < RTOL_PROFILE, c
            assert relerr(got, go['odeint_c_end_%d' % c], cs, floor=1e-9) < RTOL_PROFILE, c


def test_passive_species_elimination_is_exact(bk, resultsdir):
    """Steady solves take passive species (CO, H2: neutral, no homogeneous reaction) out of the block system and
    write their linear steady profile in closed form (include/catint_pnp.h: CATINT_PNP_MODE_KEEP_ALL).  The result
    must agree with the full block system to rounding -- same discrete root -- on cells from the kinetic to the
    transport-limited end of the C2 sweep, and the passive profiles must be exactly linear with slope -J/D."""
    from catint_b200 import backend as be, workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    tp = Transport(resultsdir=resultsdir, **workloads.c2(n_potentials=8))
    batch, _ = build_cell_batch(tp)
    db = bk.upload(batch)
    red = bk.solve(db, [200.0], mode=be.MODE_STEADY)
    red = {k: v.clone() for k, v in red.items()}
    full = bk.solve(db, [200.0], mode=be.MODE_STEADY | be.MODE_KEEP_ALL)
    assert red['status'].tolist() == [0] * 8 and full['status'].tolist() == [0] * 8
    S = batch.S
    cs = np.max(np.abs(batch.par[0, :S]))
    a, b = red['c'][-1].cpu().numpy(), full['c'][-1].cpu().numpy()
    assert relerr(a, b, cs, floor=1e-9) < 1e-8
    assert np.max(np.abs(red['flux'].cpu().numpy() - full['flux'].cpu().numpy())) < 1e-8 * np.max(np.abs(batch.par[:, S:2 * S]))
    assert np.max(np.abs(red['phi'][-1].cpu().numpy() - full['phi'][-1].cpu().numpy())) < 1e-9
    x = np.arange(batch.nx_max) * batch.par[0, 3 * S + 5]
    for name in ('CO', 'H2'):
        k = batch.species.index(name)
        for c in range(8):
            J, D = batch.par[c, S + k], batch.par[c, 2 * S + k]
            want = batch.par[c, k] + J / D * (x[-1] - x)
            assert np.max(np.abs(a[c, :, k] - want)) <= 1e-13 * max(np.max(np.abs(want)), 1e-300)


def test_warm_start_from_results_folder(bk, resultsdir):
    """system['init_folder']: a previous results folder initialises every cell (reference: calculator.py:303-309);
    restarting from the converged sweep must reproduce it with a fraction of the steps."""
    from catint_b200 import workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import Calculator
    kw = workloads.c2(n_potentials=6, phi_min=-0.8, phi_max=-1.1)
    tp = Transport(resultsdir=resultsdir, model_name='first', **kw)
    tp.set_calculator('odeint')
    c1 = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary')
    r1 = c1.run()
    kw2 = workloads.c2(n_potentials=6, phi_min=-0.8, phi_max=-1.1)
    kw2['system']['init_folder'] = tp.outputfoldername
    tp2 = Transport(resultsdir=resultsdir, model_name='second', **kw2)
    tp2.set_calculator('odeint')
    c2 = Calculator(transport=tp2, dt=0.5, tmax=200, ntout=1, mode='stationary')
    r2 = c2.run()
    assert c2.stats['converged'] == 6
    cs = 93.7
    assert relerr(r2['c'][-1], r1['c'][-1], cs) < 1e-8
    assert r2['n_steps'].max() < 0.5 * r1['n_steps'].min()


def test_continuation_reaches_the_same_states_with_fewer_steps(bk, resultsdir):
    """Calculator(continuation=k): every k-th cell of the sweep cold, the others from the converged state of their
    nearest cold neighbour (batch analogue of the reference's 'internal-cont', transport.py:834-842).  All cells
    must end on the steady states of the plain run (same discrete root, 1e-6) and the warm cells must need
    fewer than half the steps."""
    from catint_b200 import workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import Calculator
    kw = workloads.c2(n_potentials=48)
    tp = Transport(resultsdir=resultsdir, model_name='plain', **kw)
    tp.set_calculator('odeint')
    plain = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary')
    r0 = plain.run()
    tp2 = Transport(resultsdir=resultsdir, model_name='cont', **workloads.c2(n_potentials=48))
    tp2.set_calculator('odeint')
    cont = Calculator(transport=tp2, dt=0.5, tmax=200, ntout=1, mode='stationary', continuation=4)
    r1 = cont.run()
    assert plain.stats['converged'] == 48 and cont.stats['converged'] == 48
    cs = 93.7
    assert relerr(r1['c'][-1], r0['c'][-1], cs) < RTOL_PROFILE
    assert np.max(np.abs(r1['flux'] - r0['flux'])) < RTOL_PROFILE * np.max(np.abs(r0['flux']))
    st = cont.continuation_stats
    assert st['cold_cells'] == 13 and st['warm_cells'] == 35
    assert st['warm_steps_mean'] < 0.5 * st['cold_steps_mean']


def test_continuation_solves_failed_warm_cells_again_from_the_bulk_state(bk, resultsdir):
    """a warm-started cell that does not converge (here: every one, the warm wave is capped at one step) is solved
    again cold, so the continuation run can only fail where a plain run fails: all cells converge, the states are
    the plain run's, and the step counts of the re-run cells are those of cold cells"""
    from catint_b200 import workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import Calculator, build_cell_batch
    tp = Transport(resultsdir=resultsdir, model_name='rerun', **workloads.c2(n_potentials=24))
    tp.set_calculator('odeint')
    batch, _ = build_cell_batch(tp)
    plain = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary').solve_batch(batch)
    cont = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary', continuation=4)
    cont.continuation_warm_cap = 1                 # one step never reaches t = 200 s (the first step is at most 0.2 s)
    r = cont.run_continuation(batch)
    st = cont.continuation_stats
    assert st['cold_cells'] == 7 and st['warm_cells'] == 17 and st['rerun_cold_cells'] == 17
    assert np.all(r['status'] == 0) and np.all(plain['status'] == 0)
    assert np.array_equal(r['n_steps'], plain['n_steps'])          # every cell ended up integrated from the bulk state
    assert relerr(r['c'][-1], plain['c'][-1], 93.7) < 1e-10
    assert np.array_equal(r['flux'], plain['flux'])
