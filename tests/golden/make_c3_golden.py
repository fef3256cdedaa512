"""Goldens for C3-type cells: CO2R/KHCO3 with the Stern-layer (Robin) Poisson boundary on a geometric mesh
(first interval 0.05 nm), phiM x bulk_pH corner cells.

    OMP_NUM_THREADS=1 python tests/golden/make_c3_golden.py 101|201|1001

  101 nodes : the REFERENCE'S integrator (scipy odeint, dense FD Jacobian, default rtol/atol) on the restated
              RHS with the Stern/graded extension -- end state, Newton root, and next to it the CPU BDF
              (oracle/bdf_local.py) on the local form of the same cells: the rung that pins bdf_local for the
              Stern/graded extension against odeint (keys odeint_*, bdf_vs_odeint_*).
  201, 1001 : dense odeint is impractical (N = 1608 / 8008 with a very stiff mesh); CPU BDF + Newton root.
A cell whose integrator stops (step underflow: the discrete ODE leaves every bounded range in finite time)
is stored with ok_<c> = False and the time it stopped at.  Model arrays come from catint_b200's Transport
(pinned against the reference Transport by ref_*.npz); no /root/reference needed.
"""
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
os.environ.setdefault('CATINT_QUIET', '1')

BLOWUP = 1e3


class BlowUp(Exception):
    pass


def make_batch(nn):
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    from catint_b200 import workloads
    tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c3(n_phi=2, n_pH=2))
    batch, _ = build_cell_batch(tp, poisson_bc='stern', mesh=workloads.geometric_mesh(nn, 5e-11))
    return batch


def work(job):
    nn, c, with_odeint = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle.fixtures import system_from_batch
    from oracle.pnp_local import LocalForm
    from oracle.bdf_local import BdfIntegrator
    from oracle.pnp_oracle import steady_tmesh
    batch = make_batch(nn)
    s = system_from_batch(batch, c)
    lf = LocalForm(s)
    res = {'cell': c}
    t0 = time.time()
    integ = BdfIntegrator(lf, fresh_jacobian=False)
    try:
        y_end = integ.integrate([200.0], max_steps=400000)[-1]
        ok = bool(np.all(np.isfinite(y_end)))
        res['bdf_msg'] = 'reached t=200 s in %d steps' % integ.stats.nst
    except RuntimeError as e:
        ok = False
        res['bdf_msg'] = str(e)
        res['fail_t'] = float(integ.t)
    res['ok'] = ok
    res['bdf_wall'] = time.time() - t0
    if ok:
        y, info = lf.solve_steady(y0=y_end, pure_newton=True)
        assert info['converged'], c
        C, v, g = lf.unpack(y)
        res.update(newton_c=C.T.copy(), phi=v, g=g, bdf_c=y_end[:, :s.S].copy())
    if with_odeint:
        from scipy.integrate import odeint
        cmax_bulk = float(np.max(np.abs(s.c_bulk)))
        st = {'nfe': 0, 't': 0.0, 'cmax': 1.0}

        def f(cvec, t):
            st['nfe'] += 1
            cm = float(np.max(np.abs(cvec))) / cmax_bulk
            st['t'], st['cmax'] = float(t), cm
            if not np.isfinite(cm) or cm > BLOWUP:
                raise BlowUp('max|c| = %.3g x bulk at t = %.6e' % (cm, t))
            return s.rhs(cvec)
        t0 = time.time()
        try:
            sol, info = odeint(f, s.c0_flat(), steady_tmesh(), full_output=True, mxstep=5000000)
            res['odeint_msg'] = info['message']
            res['odeint_ok'] = bool(info['message'] == 'Integration successful.' and np.all(np.isfinite(sol[-1])))
        except BlowUp as e:
            res['odeint_msg'] = 'aborted by the blow-up guard: %s' % e
            res['odeint_ok'] = False
            res['odeint_fail_t'] = st['t']
        res['odeint_nfe'] = st['nfe']
        res['odeint_wall'] = time.time() - t0
        if res['odeint_ok']:
            Ce = sol[-1].reshape(s.S, s.n)
            yo, info = lf.solve_steady(y0=lf.y_from_c(Ce), pure_newton=True)
            assert info['converged'], c
            Co, vo, go_ = lf.unpack(yo)
            res.update(odeint_c_end=Ce.T.copy(), odeint_newton_c=Co.T.copy(), odeint_phi=vo, odeint_g=go_)
            if ok:
                cs = cmax_bulk
                res['bdf_vs_odeint_end'] = float(np.max(np.abs(res['bdf_c'] - Ce.T) / (np.abs(Ce.T) + 1e-9 * cs)))
                res['bdf_vs_odeint_root'] = float(np.max(np.abs(res['newton_c'] - Co.T) / (np.abs(Co.T) + 1e-12 * cs)))
    print('n %d cell %d: bdf ok %s (%s, %.0fs)%s' % (
        nn, c, ok, res['bdf_msg'], res['bdf_wall'],
        '' if not with_odeint else ' | odeint ok %s (%s, nfe %d, %.0fs) bdf-vs-odeint end %.2e root %.2e' % (
            res['odeint_ok'], res['odeint_msg'], res['odeint_nfe'], res['odeint_wall'],
            res.get('bdf_vs_odeint_end', np.nan), res.get('bdf_vs_odeint_root', np.nan))), flush=True)
    return res


def main():
    nn = int(sys.argv[1]) if len(sys.argv) > 1 else 201
    batch = make_batch(nn)
    with mp.Pool(min(batch.B, os.cpu_count())) as pool:
        results = pool.map(work, [(nn, c, nn <= 101) for c in range(batch.B)], chunksize=1)
    out = dict(par=batch.par, mesh=batch.mesh_xi[0], nodes=np.array(nn))
    for r in results:
        c = r.pop('cell')
        for k, v in r.items():
            out['%s_%d' % (k, c)] = np.array(v)
    np.savez_compressed(os.path.join(HERE, 'oracle_c3_cells_n%d.npz' % nn), **out)


if __name__ == '__main__':
    main()
