"""BASELINE.json configs C3, C4, C5 launched at their stated sizes on one GPU, through the C ABI:
status histogram, size-independent properties of every converged cell, golden cells inside the big batch.
(C2 at size: tests/test_gpu_parity.py::test_steady_state_is_a_root_of_the_rhs_kernel.)

Expected failure sets.  Fixed-flux PNP cells whose imposed current cannot be carried have no bounded
solution; the reference's integrator (scipy odeint) leaves every bounded range in finite time on them
(tests/golden/make_c4_golden.py, make_c3_golden.py 101 hold the odeint evidence).  Which cells those are is
a property of the model, so the tests assert their LOCATION in the sweep (thin layers in C4; strongly
cathodic potentials at low buffer strength in C3) and a count window around the measured count; the window
only allows for cells next to the boundary of that region to flip with rounding-level changes of the
step sequence (measured counts in the asserts' comments).
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def bk():
    import torch
    from catint_b200 import backend as be
    if not torch.cuda.is_available():
        pytest.skip('GPU tests need a B200')
    return be.PnpBackend('cuda:0')


def relerr(got, want, cscale, floor=1e-12):
    return float(np.max(np.abs(got - want) / (np.abs(want) + floor * cscale)))


def histogram(status):
    from catint_b200 import backend as be
    return {be.CELL_STATUS.get(int(k), int(k)): int(v) for k, v in zip(*np.unique(status, return_counts=True))}


def test_c4_full_sweep_65536_cells(bk, resultsdir):
    """C4: 10 species, 256 bulk_pH x 256 boundary thicknesses (10..200 um, ragged 101/102 nodes), block 11."""
    import torch
    from catint_b200 import backend as be, workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    tp = Transport(resultsdir=resultsdir, **workloads.c4())
    batch, _ = build_cell_batch(tp)
    assert batch.B == 65536 and batch.b == 11 and set(np.unique(batch.nx)) == {101, 102}
    db = bk.upload(batch)
    out = bk.solve(db, [200.0], mode=be.MODE_STEADY, max_steps=20000)
    status = out['status'].cpu().numpy()
    ok = status == 0
    hist = histogram(status)
    print('C4 status histogram', hist)
    assert set(hist) <= {'converged', 'step_underflow'}, hist
    # measured: 60327 converged, 5209 step underflow (7.9 %)
    assert 5209 - 300 <= hist.get('step_underflow', 0) <= 5209 + 300, hist
    # the failures are exactly the thin layers: for every bulk_pH row they are a prefix of the thickness axis --
    # separately for the cells with 101 and with 102 nodes (the reference's arange mesh overshoots xmax by one
    # interval for some thicknesses, SURVEY C-6; the grid-scale mode that blows up depends on the parity of the
    # node count: measured, thin layers with 102 nodes stay bounded)
    grid = ok.reshape(256, 256)                       # [pH, L], L ascending
    nxg = batch.nx.reshape(256, 256)
    assert np.all(nxg == nxg[0][None, :])
    for cls in (101, 102):
        cols = np.nonzero(nxg[0] == cls)[0]
        sub = grid[:, cols]
        first_ok = np.argmax(sub, axis=1)
        for i in range(256):
            assert sub[i, first_ok[i]:].all() and not sub[i, :first_ok[i]].any(), (cls, i)
    S = batch.S
    L = batch.par[:, 3 * S + 5].reshape(256, 256) * 100.0
    assert L[~grid].max() < 2.0e-5                   # nothing thicker than 20 um fails (dx < 0.2 um)
    assert L[grid].min() < 1.5e-5                    # and some layers thinner than 15 um do converge
    # the golden (odeint) cells of the 4x4 corner grid sit inside this sweep: same verdict
    go = load_golden('oracle_c4_cells.npz')
    for c in range(16):
        i, j = divmod(c, 4)
        big = (i * 85) * 256 + (j * 85)
        assert np.allclose(batch.par[big], go['par'][c], rtol=1e-12, atol=0), c
        assert bool(go['ok_%d' % c]) == bool(ok[big]), c
        if ok[big]:
            n = int(batch.nx[big])
            cs = np.max(np.abs(batch.par[big, :S]))
            assert relerr(out['c'][-1, big, :n].cpu().numpy(), go['newton_c_%d' % c], cs) < 1e-6, c
    # every converged cell zeroes K1's dc/dt and carries exactly the imposed wall fluxes
    c = out['c'][-1].contiguous()
    dcdt, _, _ = bk.rhs(db, c)
    D = torch.tensor(batch.par[:, 2 * S:3 * S], device=c.device)
    dx = torch.tensor(batch.par[:, 3 * S + 5], device=c.device)
    scale = (c.abs().amax(dim=1) * D / dx[:, None] ** 2).amax(dim=1)
    ratio = (dcdt.abs().amax(dim=(1, 2)) / scale).cpu().numpy()
    assert np.all(ratio[ok] < 1e-7), float(ratio[ok].max())
    assert float(np.median(ratio[ok])) < 1e-9
    flux = out['flux'].cpu().numpy()
    J = batch.par[:, S:2 * S]
    assert np.max(np.abs(flux[ok] - J[ok])) <= 1e-8 * np.max(np.abs(J))


def test_c3_full_sweep_16384_cells(bk, resultsdir):
    """C3: Stern (Robin) boundary, 1001-node graded mesh, 128 phiM x 128 bulk_pH, block 10, state in the
    workspace (global-state kernel variant)."""
    from catint_b200 import backend as be, workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    tp = Transport(resultsdir=resultsdir, **workloads.c3())
    batch, _ = build_cell_batch(tp, poisson_bc='stern', mesh=workloads.geometric_mesh(1001, 5e-11))
    assert batch.B == 16384 and batch.b == 10 and batch.nx_max == 1001
    db = bk.upload(batch)
    out = bk.solve(db, [200.0], mode=be.MODE_STEADY, max_steps=50000)
    status = out['status'].cpu().numpy()
    ok = status == 0
    hist = histogram(status)
    print('C3 status histogram', hist)
    assert set(hist) <= {'converged', 'step_underflow'}, hist
    grid = ok.reshape(128, 128)                       # [phiM (-0.5 -> -1.5), pH (6.0 -> 7.8)]
    # failures: strongly cathodic potential (large current) at low buffer strength (low pH) -- in every pH column
    # they are a suffix of the potential axis, and a column that converges everywhere stays so for higher pH
    n_fail_col = (~grid).sum(axis=0)
    for j in range(128):
        k = 128 - n_fail_col[j]
        assert grid[:k, j].all() and not grid[k:, j].any(), j
    assert np.all(np.diff(n_fail_col) <= 0)
    assert grid[:64].all()                            # phiM >= -1.0 V always has a bounded solution
    # measured on the rank-0 shard (every 8th cell): 288 of 2048 -> ~2300 of 16384
    assert 1900 <= (~ok).sum() <= 2700, hist
    # corner cells of the sweep = the golden cells (CPU BDF pinned against odeint one rung down, make_c3_golden.py)
    go = load_golden('oracle_c3_cells_n1001.npz')
    S = batch.S
    for c, big in enumerate([0, 127, 127 * 128, 127 * 128 + 127]):
        assert np.allclose(batch.par[big], go['par'][c], rtol=1e-12, atol=0), c
        assert bool(go['ok_%d' % c]) == bool(ok[big]), c
        if ok[big]:
            cs = np.max(np.abs(batch.par[big, :S]))
            assert relerr(out['c'][-1, big].cpu().numpy(), go['newton_c_%d' % c], cs) < 1e-6, c
    # converged cells carry exactly the imposed wall fluxes, K+ is depleted at the cathode, CO2 is consumed
    flux = out['flux'].cpu().numpy()
    J = batch.par[:, S:2 * S]
    assert np.max(np.abs(flux[ok] - J[ok])) <= 1e-8 * np.max(np.abs(J))
    names = batch.species
    cw = out['c'][-1, :, 0, :].cpu().numpy()
    assert np.all(cw[ok, names.index('CO2')] < batch.par[ok, names.index('CO2')])


def test_c5_transient_4096_cells(bk, resultsdir):
    """C5: 4096 cells x 5001-node graded mesh, time-dependent from the bulk state to the steady state with
    per-cell adaptive dt (first step ~1e-12 s), outputs at 1e-6, 1e-3, 1 and 200 s."""
    from catint_b200 import backend as be, workloads
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    tp = Transport(resultsdir=resultsdir, **workloads.c5())
    batch, _ = build_cell_batch(tp, mesh=workloads.geometric_mesh(5001, workloads.C5_FIRST_SPACING))
    assert batch.B == 4096 and batch.b == 9 and batch.nx_max == 5001
    db = bk.upload(batch)
    out = bk.solve(db, workloads.C5_T_OUT, mode=be.MODE_TRANSIENT, max_steps=100000)
    status = out['status'].cpu().numpy()
    hist = histogram(status)
    print('C5 status histogram', hist, 'steps mean/max', float(out['n_steps'].double().mean()), int(out['n_steps'].max()))
    assert hist == {'converged': 4096}, hist
    go = load_golden('oracle_c5_cells_n5001.npz')
    S = batch.S
    t_out = go['t_out']
    assert np.allclose(t_out, workloads.C5_T_OUT)
    for k, c in enumerate(int(x) for x in go['cells']):
        assert np.allclose(batch.par[c], go['par'][k], rtol=1e-12, atol=0), c
        cs = np.max(np.abs(batch.par[c, :S]))
        keep = go['nodes_kept_%d' % c]
        sens = go['sensitivity_f6_%d' % c]
        for j in range(len(t_out)):
            # tolerance = measured conditioning of this output time (1e-13 perturbation of the initial state,
            # make_c5_golden.py) with a safety factor, never tighter than 1e-5 (two integrators at rtol 1.5e-8)
            tol = max(1e-5, 30.0 * float(sens[j]))
            got = out['c'][j, c].cpu().numpy()[keep]
            assert relerr(got, go['bdf_c_%d' % c][j], cs, floor=1e-6) < tol, (c, j, tol)
        # the end of the transient is the steady state: Newton root of the discrete residual
        assert relerr(out['c'][-1, c].cpu().numpy(), go['newton_c_%d' % c], cs, floor=1e-9) < 1e-6, c
        gsc = np.max(np.abs(go['g_%d' % c]))
        assert np.max(np.abs(out['g'][-1, c].cpu().numpy() - go['g_%d' % c])) < 1e-6 * gsc
    flux = out['flux'].cpu().numpy()
    J = batch.par[:, S:2 * S]
    assert np.max(np.abs(flux - J)) <= 1e-6 * np.max(np.abs(J))
    # the product CO (zero in the bulk, constant wall flux) accumulates monotonically at the wall; CO2 ends depleted
    names = batch.species
    co = out['c'][:, :, 0, names.index('CO')].cpu().numpy()
    assert np.all(np.diff(co, axis=0) >= -1e-9 * cs) and np.all(co[-1] > 0)
    co2 = out['c'][:, :, 0, names.index('CO2')].cpu().numpy()
    assert np.all(co2[-1] < batch.par[:, names.index('CO2')])
