"""Golden results for 16 cells of the C4-type sweep (10 species, bulk_pH x boundary thickness, ragged
101/102 nodes) from the REFERENCE'S INTEGRATOR: scipy ``odeint`` (ODEPACK LSODA, dense finite-difference
Jacobian, default rtol/atol -- /root/reference/catint/calculator_old.py:947 without ml/mu) on the
restated RHS (oracle/pnp_oracle.py, pinned against the reference's own ``ode_func``).

For every cell the script stores
  ok_<c>            odeint reached t = 200 s with a finite state
  msg_<c>           ODEPACK's message
  c_end_<c>         odeint end state [n,S]                         (ok cells)
  newton_c_<c>, phi_<c>, g_<c>   Newton root of the same discrete residual started from that end state
  nfe_<c>, wall_<c>
and, for the cells where odeint does NOT get there (thin layers, dx = 0.1 um: the reference's discrete ODE
blows up in finite time),
  fail_t_<c>        the last time the RHS was evaluated at
  fail_cmax_<c>     max |c| / max c_bulk of that state
  growth_<c>        [k,3] samples (t, max|c|/max c_bulk, RHS calls) along the run -- the evidence of the blow-up
so that every expected-fail cell in tests/test_gpu_parity.py is an *odeint*-fail cell.

No /root/reference needed at run time (the model arrays come from catint_b200's Transport, whose parity with
the reference Transport is pinned by ref_*.npz).  ~10-40 min on 8 cores:

    OMP_NUM_THREADS=1 python tests/golden/make_c4_golden.py
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

BLOWUP = 1e3          # |c| > BLOWUP * max c_bulk: the trajectory has left every physical range
WALL_BUDGET = float(os.environ.get('CATINT_GOLDEN_BUDGET_S', '5400'))


class BlowUp(Exception):
    pass


def work(job):
    c, z, reactions, par, nx = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from scipy.integrate import odeint
    from oracle.pnp_oracle import PnpSystem, steady_tmesh
    from oracle.pnp_local import LocalForm
    S = len(z)
    n = int(nx)
    x = np.arange(n) * par[3 * S + 5]
    s = PnpSystem(z=z, D=par[2 * S:3 * S], c_bulk=par[0:S], J=par[S:2 * S], x=x, beta=par[3 * S], eps=par[3 * S + 1],
                  reactions=reactions, rate_mode='summed', use_migration=True, phi_wall=par[3 * S + 2],
                  g_bulk=par[3 * S + 3], uniform=True)
    cmax_bulk = float(np.max(np.abs(s.c_bulk)))
    st = {'nfe': 0, 't': 0.0, 'cmax': 1.0, 'growth': [], 't0': time.time()}

    def f(cvec, t):
        st['nfe'] += 1
        cm = float(np.max(np.abs(cvec))) / cmax_bulk
        st['t'], st['cmax'] = float(t), cm
        if st['nfe'] % 2000 == 0 or cm > 10.0 * (st['growth'][-1][1] if st['growth'] else 1.0):
            st['growth'].append((float(t), cm, st['nfe']))
        if not np.isfinite(cm) or cm > BLOWUP:
            raise BlowUp('max|c| = %.3g x bulk at t = %.6e' % (cm, t))
        if time.time() - st['t0'] > WALL_BUDGET:
            raise BlowUp('wall budget exhausted at t = %.6e, max|c| = %.3g x bulk' % (t, cm))
        return s.rhs(cvec)

    res = {'cell': c, 'n': n}
    tm = steady_tmesh()
    try:
        sol, info = odeint(f, s.c0_flat(), tm, full_output=True, mxstep=5000000)
        msg = info['message']
        ok = bool(msg == 'Integration successful.' and np.all(np.isfinite(sol[-1])))
    except BlowUp as e:
        ok, msg, sol = False, 'aborted by the blow-up guard: %s' % e, None
    res.update(ok=ok, msg=msg, nfe=st['nfe'], wall=time.time() - st['t0'],
               growth=np.array(st['growth'] + [(st['t'], st['cmax'], st['nfe'])]))
    if ok:
        C = sol[-1].reshape(S, n)
        lf = LocalForm(s)
        y, inf2 = lf.solve_steady(y0=lf.y_from_c(C), pure_newton=True)
        Cn, vn, gn = lf.unpack(y)
        res.update(c_end=C.T.copy(), newton_c=Cn.T.copy(), phi=vn, g=gn, newton_ok=bool(inf2['converged']))
    else:
        res.update(fail_t=st['t'], fail_cmax=st['cmax'])
    print('cell %2d n %d ok %s nfe %d wall %.0fs  %s' % (c, n, ok, st['nfe'], res['wall'], msg), flush=True)
    return res


def main():
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    from catint_b200 import workloads
    tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c4(n_pH=4, n_L=4))
    batch, _ = build_cell_batch(tp)
    cells = list(range(batch.B))
    jobs = [(c, batch.z, batch.reactions, batch.par[c], batch.nx[c]) for c in cells]
    with mp.Pool(min(len(jobs), os.cpu_count())) as pool:
        results = pool.map(work, jobs, chunksize=1)
    out = dict(cells=np.array(cells), par=batch.par[cells], nx=batch.nx[cells],
               integrator=np.array('scipy.integrate.odeint (LSODA), dense FD Jacobian, default rtol/atol, t_end=200 s'))
    for r in results:
        c = r['cell']
        out['ok_%d' % c] = np.array(r['ok'])
        out['msg_%d' % c] = np.array(r['msg'])
        out['nfe_%d' % c] = np.array(r['nfe'])
        out['wall_%d' % c] = np.array(r['wall'])
        out['growth_%d' % c] = r['growth']
        if r['ok']:
            assert r['newton_ok'], c
            for k in ('c_end', 'newton_c', 'phi', 'g'):
                out['%s_%d' % (k, c)] = r[k]
        else:
            out['fail_t_%d' % c] = np.array(r['fail_t'])
            out['fail_cmax_%d' % c] = np.array(r['fail_cmax'])
    np.savez_compressed(os.path.join(HERE, 'oracle_c4_cells.npz'), **out)


if __name__ == '__main__':
    main()
