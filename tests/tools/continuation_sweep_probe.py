"""CPU probe (oracle BDF = prototype of the CUDA integrator) of a whole C2 sweep with Calculator(continuation=k): every warm
cell from the nearest cold state and from the interpolated one, steps and distance of the end state from its cold run.
    OMP_NUM_THREADS=1 python tests/tools/continuation_sweep_probe.py 48 4      (the configuration of the GPU test)
Test infrastructure (uses oracle/)."""
import os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np, multiprocessing as mp
from catint_b200 import workloads
from catint_b200.transport import Transport
from catint_b200.calculator import build_cell_batch, continuation_plan, continuation_brackets
from continuation_axis_probe import system_of
from oracle.pnp_local import LocalForm
from oracle.bdf_local import BdfIntegrator
N, K = int(sys.argv[1]), int(sys.argv[2])
tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c2(n_potentials=N)); tp.set_calculator('odeint')
batch, _ = build_cell_batch(tp)
def cold_run(c):
    lf = LocalForm(system_of(batch, c)); b = BdfIntegrator(lf)
    return lf.unpack(b.integrate([200.0])[-1])[0], b.stats.nst
def warm_run(args):
    c, C0 = args
    lf = LocalForm(system_of(batch, c)); b = BdfIntegrator(lf)
    return lf.unpack(b.integrate([200.0], y0=lf.y_from_c(C0))[-1])[0], b.stats.nst
if __name__ == '__main__':
    with mp.Pool(8) as pool:
        cold_all = pool.map(cold_run, range(N))
        cold, warm, near = continuation_plan(N, K)
        left, right, w = continuation_brackets(cold, warm)
        starts = [(int(c), cold_all[cold[l]][0] + (cold_all[cold[r]][0] - cold_all[cold[l]][0]) * ww) for c, l, r, ww in zip(warm, left, right, w)]
        res = pool.map(warm_run, starts)
        starts_n = [(int(c), cold_all[cold[p]][0]) for c, p in zip(warm, near)]
        res_n = pool.map(warm_run, starts_n)
    worst = 0.0
    for (c, _), (C, nst), (Cn, nstn) in zip(starts, res, res_n):
        ref = cold_all[c][0]
        d = np.max(np.abs(C - ref)) / np.max(np.abs(ref)); dn = np.max(np.abs(Cn - ref)) / np.max(np.abs(ref))
        worst = max(worst, d)
        print('cell %3d: cold %4d steps, nearest %4d (dev %.1e), interp %4d (dev %.1e)' % (c, cold_all[c][1], nstn, dn, nst, d))
    print('N=%d k=%d: worst deviation of the interpolated starts from the cold end states %.2e; mean steps cold %.0f nearest %.0f interp %.0f'
          % (N, K, worst, np.mean([cold_all[c][1] for c in warm]), np.mean([r[1] for r in res_n]), np.mean([r[1] for r in res])))
