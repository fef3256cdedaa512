"""Goldens for the flux-equation wall kinetics: C1 chemistry with CO2 reduction first order in the
SURFACE CO2 concentration and Tafel-type potential dependence (workloads.c2_kinetic), four potentials from the
kinetic to the mass-transport limited regime.  Integrator: the reference's scipy odeint (dense FD Jacobian,
default tolerances) on the restated RHS whose wall flux is J_fixed + coef @ E(c(0), phi(0)), the expressions
being evaluated by Python itself (oracle/flux_expr.py -- independent of the product's compiler and of the
device interpreter).  Also stored: the oracle RHS on perturbed states and the wall block of the oracle
Jacobian, for the K1 / K2 parity tests.

    OMP_NUM_THREADS=1 python tests/golden/make_fluxeq_golden.py
"""
import multiprocessing as mp
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
os.environ.setdefault('CATINT_QUIET', '1')

PHIS = [-0.7, -0.9, -1.0, -1.1]


def make(stern=False):
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    from catint_b200 import workloads
    kw, ca = workloads.c2_kinetic(stern=stern)
    kw['descriptors'] = {'phiM': list(PHIS)}
    tp = Transport(resultsdir=tempfile.mkdtemp(), comsol_args=ca, **kw)
    if stern:
        batch, _ = build_cell_batch(tp, poisson_bc='stern', mesh=workloads.geometric_mesh(101, 5e-11))
    else:
        batch, _ = build_cell_batch(tp)
    exprs = [kw['species'][o]['flux-equation'] for o in tp.flux_eq.owners]
    return tp, batch, exprs


def oracle_system(batch, c, exprs):
    from oracle.fixtures import system_from_batch
    from oracle.flux_expr import WallKinetics
    s = system_from_batch(batch, c)
    params = dict(zip(batch.flux_eq.par_names, batch.fpar[c]))
    s.wall_kinetics = WallKinetics(batch.species, exprs, batch.flux_eq.coef, params)
    return s


def work(job):
    stern, c = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle.pnp_oracle import steady_tmesh
    from oracle.pnp_local import LocalForm
    tp, batch, exprs = make(stern)
    s = oracle_system(batch, c, exprs)
    sol, info = s.integrate_odeint(steady_tmesh(), full_output=True)
    assert info['message'] == 'Integration successful.', info['message']
    C = sol[-1].reshape(s.S, s.n)
    lf = LocalForm(s)
    y, inf2 = lf.solve_steady(y0=lf.y_from_c(C), pure_newton=True)
    assert inf2['converged']
    Cn, vn, gn = lf.unpack(y)
    phi0 = vn[0]
    J = s.J + s.wall_kinetics(Cn[:, 0], phi0)
    print('stern %s phiM %.2f: nfe %d, CO2(0) %.4f of %.4f, i_CO %.2f i_H2 %.2f A/m^2' % (
        stern, PHIS[c], int(info['nfe'][-1]), Cn[batch.species.index('CO2'), 0], s.c_bulk[batch.species.index('CO2')],
        -2 * 96485.33289 * J[batch.species.index('CO')], -2 * 96485.33289 * J[batch.species.index('H2')]), flush=True)
    # K1 / K2 material: oracle RHS and wall Jacobian block on a perturbed state
    rng = np.random.default_rng(c)
    Cp = Cn * (1.0 + 0.05 * rng.standard_normal(Cn.shape))
    rhs_p = s.rhs(Cp.reshape(-1)).reshape(s.S, s.n)
    yp = lf.y_from_c(Cp)
    F, L, Dg, U, E0 = lf.residual(yp, blocks=True)
    return dict(c=c, c_end=C.T.copy(), newton_c=Cn.T.copy(), phi=vn, g=gn, J=J, state=Cp.T.copy(), rhs=rhs_p.T.copy(),
                y_state=yp, F0=F[0], D0=Dg[0])


def main():
    out = {}
    with mp.Pool(8) as pool:
        results = pool.map(work, [(st, c) for st in (False, True) for c in range(len(PHIS))], chunksize=1)
    for st in (False, True):
        tp, batch, exprs = make(st)
        tag = 'stern' if st else 'dirichlet'
        out['%s_par' % tag] = batch.par
        out['%s_fpar' % tag] = batch.fpar
        out['%s_coef' % tag] = batch.flux_eq.coef
        out['%s_par_names' % tag] = np.array(batch.flux_eq.par_names)
        for r in results[(4 if st else 0):(8 if st else 4)]:
            for k, v in r.items():
                if k != 'c':
                    out['%s_%s_%d' % (tag, k, r['c'])] = np.array(v)
    out['phis'] = np.array(PHIS)
    np.savez_compressed(os.path.join(HERE, 'oracle_fluxeq.npz'), **out)


if __name__ == '__main__':
    main()
