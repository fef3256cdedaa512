"""Flux equations: wall kinetics given as expressions of the surface state.

Reference semantics: /root/reference/docs/source/topics/flux_definition.rst:100-156 (species[sp]['flux-equation'],
[[name]] = surface concentration), /root/reference/catint/comsol_model.py:986-1000 (the string handed to COMSOL is
RF*flux_factor*(expression)), stoichiometric propagation /root/reference/catint/transport.py:1057-1087.

CPU part: the product's postfix compiler against Python's own evaluation of the same strings (oracle/flux_expr.py),
the coefficient table against the fixed-flux Transport, error behaviour.
GPU part (-m gpu): K1 / K2 / K3 with the device interpreter against the odeint goldens
(tests/golden/make_fluxeq_golden.py).
"""
import numpy as np
import pytest

from conftest import load_golden

F_CONST = 96485.33289


def kinetic_transport(resultsdir, phis, stern=False):
    from catint_b200 import workloads
    from catint_b200.transport import Transport
    kw, ca = workloads.c2_kinetic(stern=stern)
    kw['descriptors'] = {'phiM': list(phis)}
    tp = Transport(resultsdir=resultsdir, comsol_args=ca, **kw)
    exprs = [kw['species'][o]['flux-equation'] for o in tp.flux_eq.owners]
    return tp, exprs


def kinetic_batch(resultsdir, phis, stern=False):
    from catint_b200 import workloads
    from catint_b200.calculator import build_cell_batch
    tp, exprs = kinetic_transport(resultsdir, phis, stern)
    if stern:
        batch, _ = build_cell_batch(tp, poisson_bc='stern', mesh=workloads.geometric_mesh(101, 5e-11))
    else:
        batch, _ = build_cell_batch(tp)
    return tp, batch, exprs


# ---------------------------------------------------------------- host ------------------------------------
def test_compiler_matches_python_evaluation():
    """postfix programs == Python's evaluation of the same strings, incl. units, variables, ^ and unary minus"""
    from catint_b200.fluxeq import FluxEquations
    from oracle.flux_expr import WallKinetics
    names = ['K+', 'CO2', 'OH-', 'CO']
    variables = {'eta': 'phiM-phi-phiEq', 'theta': 'Kads*[[CO2]]/(1+Kads*[[CO2]])'}
    exprs = ['k0*theta*exp(-alpha*F_const*eta/RT)',
             '-k1*[[OH-]]^2*sqrt([[K+]])/(1.5[mol/m^3]+[[CO]]) + 2.5e-3*tanh(eta) - log10([[CO2]])*1e-6',
             'k0*(-[[CO2]])**2*log(2+[[OH-]])']
    params = {'k0': 3e-7, 'k1': 2e-4, 'alpha': 0.45, 'F_const': F_CONST, 'RT': 2477.7, 'phiM': -0.9,
              'phiEq': -0.11, 'Kads': 0.02}
    fe = FluxEquations(names)
    for k, e in enumerate(exprs):
        fe.add(names[k], 'RF*flux_factor*(' + e + ')', variables)
    allp = dict(params, RF=1.7, flux_factor=1.0)
    par = [allp[n] for n in fe.par_names]
    wk = WallKinetics(names, exprs, np.eye(len(names))[:, :3], allp, variables)
    rng = np.random.default_rng(3)
    for _ in range(20):
        c = rng.uniform(0.01, 50.0, size=len(names))
        phi = rng.uniform(-1.2, 0.2)
        want = wk.equations(c, phi)
        got = np.array([fe.evaluate(e, c, phi, par) for e in range(3)])
        assert np.allclose(got, want, rtol=1e-13, atol=0)
    # complex-step derivatives of the oracle against central differences of the compiled program
    c = np.array([12.0, 30.0, 0.5, 0.02]); phi = -0.75
    dc, dphi = wk.jacobian(c, phi)
    for j in range(len(names)):
        h = 1e-6 * c[j]
        cp, cm = c.copy(), c.copy()
        cp[j] += h; cm[j] -= h
        fd = np.array([(fe.evaluate(e, cp, phi, par) - fe.evaluate(e, cm, phi, par)) / (2 * h) for e in range(3)])
        assert np.allclose(dc[:3, j], fd, rtol=1e-6, atol=1e-12 * np.max(np.abs(dc)))
    fd = np.array([(fe.evaluate(e, c, phi + 1e-6, par) - fe.evaluate(e, c, phi - 1e-6, par)) / 2e-6 for e in range(3)])
    assert np.allclose(dphi[:3], fd, rtol=1e-6, atol=1e-12 * np.max(np.abs(dphi)))


def test_program_limits_and_errors():
    from catint_b200.fluxeq import FluxEquations, FluxEqError, MAX_EQ
    fe = FluxEquations(['A', 'B'])
    with pytest.raises(FluxEqError, match='not a transported species'):
        fe.add('A', '[[C]]*2', {})
    with pytest.raises(FluxEqError, match='unbalanced'):
        fe.add('A', '([[A]]*2', {})
    with pytest.raises(FluxEqError, match='several arguments'):
        fe.add('A', 'exp([[A]], 2)', {})
    with pytest.raises(FluxEqError, match='too long'):
        fe.add('A', '+'.join(['[[A]]*%d.5' % k for k in range(40)]), {})
    with pytest.raises(FluxEqError, match='nested too deeply'):
        fe.add('A', 'u', {'u': 'v+1', 'v': 'u+1'})
    for _ in range(MAX_EQ):
        fe.add('A', '[[A]]', {})
    with pytest.raises(FluxEqError, match='at most'):
        fe.add('A', '[[A]]', {})


def test_flux_equation_transport_matches_fixed_flux_transport(resultsdir):
    """J_fixed + coef @ E(bulk state) must equal the flux_bound of a Transport that was given the same numbers as
    plain fluxes (the stoichiometric propagation of transport.py:1057-1087 is linear)."""
    from catint_b200 import workloads
    from catint_b200.transport import Transport
    from oracle.flux_expr import WallKinetics
    phis = [-0.8, -1.05]
    tp, exprs = kinetic_transport(resultsdir, phis)
    fe = tp.flux_eq
    assert fe.owners == ['CO', 'H2'] and fe.n_eq == 2
    names = list(tp.species)
    for p, phiM in enumerate(phis):
        m = tp.derive_for(phiM=phiM)
        params = dict(zip(fe.par_names, m.fpar))
        wk = WallKinetics(names, exprs, fe.coef, params)
        cb = np.array([tp.species[s]['bulk_concentration'] for s in names], dtype=float)
        E = wk.equations(cb, phiM)
        kw = workloads.co2r_inputs()
        kw['species']['CO'] = {'bulk_concentration': 0.0, 'flux': float(E[0])}
        kw['species']['H2'] = {'bulk_concentration': 0.0, 'flux': float(E[1])}
        ref = Transport(resultsdir=resultsdir, model_name='fixed%d' % p, **kw)
        J = m.flux_bound[:, 0] + fe.coef @ E
        assert np.allclose(J, ref.flux_bound[:, 0], rtol=1e-12, atol=1e-18)


def test_unknown_identifier_is_fatal(resultsdir):
    from catint_b200 import workloads
    from catint_b200.transport import Transport
    kw, ca = workloads.c2_kinetic()
    kw['descriptors'] = {'phiM': [-0.9]}
    del ca['parameter']['alpha']
    with pytest.raises(SystemExit):
        Transport(resultsdir=resultsdir, comsol_args=ca, **kw)


def test_batch_carries_programs_and_per_cell_parameters(resultsdir):
    from catint_b200 import backend as be
    phis = [-0.7, -0.9, -1.0, -1.1]
    tp, batch, exprs = kinetic_batch(resultsdir, phis)
    go = load_golden('oracle_fluxeq.npz')
    assert np.array_equal(batch.par, go['dirichlet_par']) and np.array_equal(batch.fpar, go['dirichlet_fpar'])
    assert np.array_equal(batch.flux_eq.coef, go['dirichlet_coef'])
    assert list(batch.flux_eq.par_names) == [str(s) for s in go['dirichlet_par_names']]
    k = batch.flux_eq.par_names.index('phiM')
    assert np.allclose(batch.fpar[:, k], phis)
    sh = batch.shared_struct()
    fq = sh.flux_eq.contents
    assert fq.n_eq == 2 and fq.n_par == len(batch.flux_eq.par_names)
    assert list(fq.code[0])[:fq.n_code[0]] == batch.flux_eq.programs[0].code
    sub = batch.select([1, 3])
    assert np.array_equal(sub.fpar, batch.fpar[[1, 3]]) and sub.flux_eq is batch.flux_eq
    assert be.MAX_FLUX_EQ == 4


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


def relerr(got, want, cscale, floor=1e-12):
    return float(np.max(np.abs(got - want) / (np.abs(want) + floor * cscale)))


@pytest.mark.gpu
def test_rhs_and_wall_jacobian_with_flux_equations(bk, resultsdir):
    """K1 (dc/dt) and K2 (wall block of the Jacobian, dF_0/dy_0 incl. dJ/dc(0)) on perturbed states against the
    oracle, whose expressions are evaluated by Python and differentiated by the complex-step method."""
    go = load_golden('oracle_fluxeq.npz')
    tp, batch, exprs = kinetic_batch(resultsdir, go['phis'])
    db = bk.upload(batch)
    B, S = batch.B, batch.S
    c = np.stack([go['dirichlet_state_%d' % k] for k in range(B)])
    dcdt, _, _ = bk.rhs(db, to_dev(c))
    dcdt = dcdt.cpu().numpy()
    y = np.stack([go['dirichlet_y_state_%d' % k] for k in range(B)])
    Fg, Lg, Dg, Ug = [a.cpu().numpy() for a in bk.jacobian(db, to_dev(y))]
    for k in range(B):
        want = go['dirichlet_rhs_%d' % k]
        assert np.max(np.abs(dcdt[k] - want)) <= 1e-11 * np.max(np.abs(want)), k
        F0, D0 = go['dirichlet_F0_%d' % k], go['dirichlet_D0_%d' % k]
        assert np.max(np.abs(Fg[k, 0] - F0)) <= 1e-12 * np.max(np.abs(F0)), k
        assert np.max(np.abs(Dg[k, 0] - D0)) <= 1e-12 * np.max(np.abs(D0)), k
        # the flux equations make the wall block dense: CO2 column of the CO row is the kinetic derivative
        i_co2, i_co = batch.species.index('CO2'), batch.species.index('CO')
        assert abs(D0[i_co, i_co2]) > 0.0 and abs(D0[i_co2, i_co2]) > abs(D0[i_co, i_co2])


@pytest.mark.gpu
@pytest.mark.parametrize('stern', [False, True])
def test_steady_state_with_flux_equations_matches_odeint(bk, resultsdir, stern):
    """K3 with the wall kinetics evaluated on the surface state (value in every residual, derivative in the wall
    block of the Newton matrix) against the reference's integrator (scipy odeint) on the restated RHS with
    the same expressions: end state, its Newton root, potential, and the wall flux == J(c(0), phi(0))."""
    from catint_b200 import backend as be
    go = load_golden('oracle_fluxeq.npz')
    tag = 'stern' if stern else 'dirichlet'
    tp, batch, exprs = kinetic_batch(resultsdir, go['phis'], stern)
    assert np.array_equal(batch.par, go['%s_par' % tag]) and np.array_equal(batch.fpar, go['%s_fpar' % tag])
    out = bk.solve(bk.upload(batch), [200.0], mode=be.MODE_STEADY, max_steps=50000)
    assert out['status'].tolist() == [0] * batch.B
    S = batch.S
    i_co2 = batch.species.index('CO2')
    depletion = []
    for k in range(batch.B):
        cs = np.max(np.abs(batch.par[k, :S]))
        got = out['c'][-1, k].cpu().numpy()
        assert relerr(got, go['%s_newton_c_%d' % (tag, k)], cs) < 1e-6, k
        assert relerr(got, go['%s_c_end_%d' % (tag, k)], cs, floor=1e-9) < 1e-6, k
        phi = go['%s_phi_%d' % (tag, k)]
        assert np.max(np.abs(out['phi'][-1, k].cpu().numpy() - phi)) < 1e-6 * np.max(np.abs(phi)), k
        J = go['%s_J_%d' % (tag, k)]
        assert np.max(np.abs(out['flux'][k].cpu().numpy() - J)) < 1e-6 * np.max(np.abs(J)), k
        depletion.append(got[0, i_co2] / batch.par[k, i_co2])
    if not stern:
        # the sweep runs from the kinetic into the mass-transport limited regime: the state dependence matters
        assert depletion[0] > 0.99 and depletion[-1] < 0.5
