"""CPU probe (oracle BDF, the algorithmic prototype of the CUDA integrator: same step/order control): how many BDF
steps does a C4 cell need from the bulk state, from the converged state of a neighbour along the bulk-pH axis, and
from one along the boundary-layer axis?  Decides the continuation axis of 2D sweeps (Calculator.run_continuation).
    OMP_NUM_THREADS=1 python tests/tools/continuation_axis_probe.py [i_pH] [i_L] [offsets ...]
Test infrastructure (lives under tests/ because it uses oracle/); not part of the product path."""
import os, sys, tempfile, time
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np
from catint_b200 import workloads
from catint_b200.transport import Transport
from catint_b200.calculator import build_cell_batch
from oracle.pnp_oracle import PnpSystem
from oracle.pnp_local import LocalForm
from oracle.bdf_local import BdfIntegrator


def system_of(batch, c):
    S = batch.S; p = batch.par[c]; n = int(batch.nx[c])
    x = np.arange(n) * p[3 * S + 5]
    return PnpSystem(z=batch.z, D=p[2 * S:3 * S], c_bulk=p[0:S], J=p[S:2 * S], x=x, beta=p[3 * S], eps=p[3 * S + 1],
                     reactions=batch.reactions, rate_mode='summed', use_migration=batch.use_migration,
                     phi_wall=p[3 * S + 2], g_bulk=p[3 * S + 3], uniform=True)


def solve(batch, c, C0=None):
    lf = LocalForm(system_of(batch, c))
    bdf = BdfIntegrator(lf)
    y0 = None if C0 is None else lf.y_from_c(C0)
    t0 = time.time()
    try:
        y = bdf.integrate([200.0], y0=y0, max_steps=20000)[-1]
    except Exception as e:
        return None, bdf.stats.nst, 'failed: %s' % e, time.time() - t0
    C = lf.unpack(y)[0]
    return C, bdf.stats.nst, 'ok', time.time() - t0


def main():
    ip = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    il = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    offs = [int(a) for a in sys.argv[3:]] or [1, 8, 32]
    kw = workloads.c4()
    tp = Transport(resultsdir=tempfile.mkdtemp(prefix='catint_probe_'), **kw)
    tp.set_calculator('odeint')
    keys = list(tp.descriptors)
    v1, v2 = tp.descriptors[keys[0]], tp.descriptors[keys[1]]
    pts = [(ip, il)] + [(ip + o, il) for o in offs] + [(ip, il + o) for o in offs]
    batch, _ = build_cell_batch(tp, points=[{keys[0]: float(v1[a]), keys[1]: float(v2[b])} for a, b in pts])
    print('descriptors: %s (outer, %d) x %s (inner, %d); base cell (%d, %d): %s = %.4g, %s = %.4g'
          % (keys[0], len(v1), keys[1], len(v2), ip, il, keys[0], v1[ip], keys[1], v2[il]), flush=True)
    C_base, nst, msg, dt = solve(batch, 0)
    print('base cell from the bulk state: %d steps (%s, %.0f s)' % (nst, msg, dt), flush=True)
    if C_base is None:
        return
    for j, (a, b) in enumerate(pts[1:], start=1):
        axis = keys[0] if b == il else keys[1]
        n = int(batch.nx[j])
        if n != C_base.shape[1]:
            print('cell (%d, %d): ragged node count, skipped' % (a, b)); continue
        Cw, nw, mw, dw = solve(batch, j, C0=C_base)
        Cc, nc, mc, dc = solve(batch, j)
        dev = float('nan')
        if Cw is not None and Cc is not None:
            dev = float(np.max(np.abs(Cw - Cc)) / np.max(np.abs(Cc)))
        print('neighbour +%d along %-20s: warm %5d steps (%s), cold %5d steps (%s), end states differ by %.1e'
              % (max(a - ip, b - il), axis, nw, mw, nc, mc, dev), flush=True)


if __name__ == '__main__':
    main()
