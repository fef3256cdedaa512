"""C2 potential sweep, plain run against Calculator(continuation=k) (the batch analogue of the reference's
'internal-cont', /root/reference/catint/transport.py:834-842), both through the host-buffer path
(distributed.solve_sharded: H2D of the parameters, solve, D2H of the results):
    python scripts/continuation_bench.py [k] [n_cells ...]        C2 sweeps of the given sizes
    python scripts/continuation_bench.py [k] c3|c4                BASELINE configs C3 / C4 at their full size
(CONT_NO_PLAIN=1 skips the plain run, CONT_PLAIN_ONCE=1 runs it without a warm-up repeat.)
Prints for every sweep size the wall time of both modes, the step counts, and how many cells of the continuation
run deviate from the plain run's steady state by more than 1e-6 (same discrete root or not)."""
import os, sys, time, tempfile
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np
import torch
from catint_b200 import workloads, distributed as D
from catint_b200.transport import Transport
from catint_b200.calculator import Calculator, build_cell_batch


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    sizes = [a if a in ('c3', 'c4') else int(a) for a in sys.argv[2:]] or [1024, 8192]
    for n in sizes:
        ckw, bkw = {}, {}
        if n == 'c3':
            kw = workloads.c3()
            ckw = bkw = dict(poisson_bc='stern', mesh=workloads.geometric_mesh(1001, 5e-11))
        elif n == 'c4':
            kw = workloads.c4()
        else:
            kw = workloads.c2(n_potentials=n)
        tp = Transport(resultsdir=tempfile.mkdtemp(prefix='catint_cont_'), model_name='cont', **kw)
        tp.set_calculator('odeint')
        batch, _ = build_cell_batch(tp, **bkw)
        label, n = n, batch.B
        plain = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary', device='cuda:0', **ckw)
        cont = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary', device='cuda:0', continuation=k, **ckw)
        print('== %s: %d cells, %d nodes, %d species' % (label, n, batch.nx_max, batch.S), flush=True)
        res = {}
        modes = (('plain', lambda: D.solve_sharded(plain, batch)), ('cont', lambda: cont.run_continuation(batch)))
        if os.environ.get('CONT_NO_PLAIN'):
            modes = modes[1:]
        for name, fn in modes:
            if not (name == 'plain' and os.environ.get('CONT_PLAIN_ONCE')):
                r = fn()                                     # warm-up (allocations, first launch)
                del r
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            res[name] = r
            nc = int(np.sum(r['status'] == 0))
            hist = {int(a): int(b) for a, b in zip(*np.unique(r['status'], return_counts=True))}
            print('%5d cells %-5s: %.4f s = %8.1f converged cells/s, converged %d, status %s, steps mean %.0f max %d'
                  % (n, name, dt, nc / dt, nc, hist, r['n_steps'].mean(), r['n_steps'].max()), flush=True)
        if getattr(cont, 'continuation_timing', None):
            print('       stages [s]:', {a: round(b, 4) for a, b in cont.continuation_timing.items()})
        st = getattr(cont, 'continuation_stats', None)
        if st:
            print('      ', {kk: (round(v, 1) if isinstance(v, float) else v) for kk, v in st.items()})
        if 'plain' not in res:
            continue
        a, b = res['plain']['c'][-1], res['cont']['c'][-1]
        both = (res['plain']['status'] == 0) & (res['cont']['status'] == 0)
        scale = np.max(np.abs(a), axis=(1, 2), keepdims=True)
        dev = np.max(np.abs(a - b) / scale, axis=(1, 2))[both]
        print('%5d cells: continuation vs plain steady states over the %d cells converged in both: worst relative '
              'deviation %.2e, cells beyond 1e-6: %d; converged only in plain %d, only with continuation %d'
              % (n, int(both.sum()), dev.max(), int(np.sum(dev > 1e-6)),
                 int(np.sum((res['plain']['status'] == 0) & (res['cont']['status'] != 0))),
                 int(np.sum((res['plain']['status'] != 0) & (res['cont']['status'] == 0)))), flush=True)


if __name__ == '__main__':
    main()
