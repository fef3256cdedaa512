"""Golden odeint results for a few cells of the C2 potential sweep (oracle only; no
/root/reference needed at run time -- the model arrays come from catint_b200's Transport,
whose parity with the reference Transport is pinned by ref_*.npz).

    OMP_NUM_THREADS=1 python tests/golden/make_sweep_golden.py            # 8 cells  -> oracle_c2_sweep.npz
    OMP_NUM_THREADS=1 python tests/golden/make_sweep_golden.py dense      # 40 more  -> oracle_c2_sweep_dense.npz
                                                                          # (every 32nd cell + the band where the
                                                                          # Tafel current saturates, cells 560-650)
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

CELLS = [0, 146, 292, 438, 585, 731, 877, 1023]
DENSE = sorted(set(list(range(16, 1024, 32)) + list(range(560, 650, 11))) - set(CELLS))


def oracle_system(batch, c, rate_mode='summed'):
    from oracle.pnp_oracle import PnpSystem
    S = batch.S
    p = batch.par[c]
    n = int(batch.nx[c])
    x = np.arange(n) * p[3 * S + 5]
    return PnpSystem(z=batch.z, D=p[2 * S:3 * S], c_bulk=p[0:S], J=p[S:2 * S], x=x, beta=p[3 * S], eps=p[3 * S + 1],
                     reactions=batch.reactions, rate_mode=rate_mode, use_migration=batch.use_migration,
                     phi_wall=p[3 * S + 2], g_bulk=p[3 * S + 3], uniform=True)


def work(args):
    par_row, c = args
    from catint_b200 import backend as be
    from oracle.pnp_oracle import steady_tmesh
    from oracle.pnp_local import LocalForm
    batch = be.CellBatch(par_row['z'], par_row['reactions'], par_row['nu'], par_row['par'][None], par_row['nx'][None])
    s = oracle_system(batch, 0)
    t0 = time.time()
    tm = steady_tmesh()
    sol, info = s.integrate_odeint(tm, full_output=True)
    C = sol[-1].reshape(s.S, s.n)
    lf = LocalForm(s)
    y, inf2 = lf.solve_steady(y0=lf.y_from_c(C), pure_newton=True)
    Cn, vn, gn = lf.unpack(y)
    return c, sol[-1], Cn.reshape(-1), vn, gn, int(info['nfe'][-1]), time.time() - t0, bool(inf2['converged'])


def main():
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    from catint_b200 import workloads
    tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c2())
    batch, _ = build_cell_batch(tp)
    dense = len(sys.argv) > 1 and sys.argv[1] == 'dense'
    CELLS = DENSE if dense else globals()['CELLS']
    jobs = [(dict(z=batch.z, reactions=batch.reactions, nu=batch.nu, par=batch.par[c], nx=batch.nx[c]), c) for c in CELLS]
    with mp.Pool(min(len(jobs), os.cpu_count())) as pool:
        res = pool.map(work, jobs, chunksize=1)
    out = dict(cells=np.array(CELLS), par=batch.par[CELLS])
    for c, cend, cn, vn, gn, nfe, wall, conv in res:
        out['c_end_%d' % c] = cend
        out['newton_c_%d' % c] = cn
        out['newton_potential_%d' % c] = vn
        out['newton_grad_%d' % c] = gn
        print('cell %4d nfe %d wall %.1fs newton %s K+(0)=%.6g' % (c, nfe, wall, conv, cn[0]))
    np.savez_compressed(os.path.join(HERE, 'oracle_c2_sweep_dense.npz' if dense else 'oracle_c2_sweep.npz'), **out)


if __name__ == '__main__':
    main()
