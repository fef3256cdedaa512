"""CPU probe (oracle BDF): C4 cells along the boundary-layer axis, stride 4 -- warm start from the nearest cold cell vs the
interpolation of the two bracketing cold cells.  Test infrastructure (uses oracle/)."""
import os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np
from catint_b200 import workloads
from catint_b200.transport import Transport
from catint_b200.calculator import build_cell_batch
from continuation_axis_probe import system_of
from oracle.pnp_local import LocalForm
from oracle.bdf_local import BdfIntegrator
tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c4()); tp.set_calculator('odeint')
keys = list(tp.descriptors); v1, v2 = tp.descriptors[keys[0]], tp.descriptors[keys[1]]
ip = 128
for il, k in ((128, 4), (60, 4), (200, 4), (128, 16)):
    idx = [il, il + k, il + k // 2]
    batch, _ = build_cell_batch(tp, points=[{keys[0]: float(v1[ip]), keys[1]: float(v2[i])} for i in idx])
    if len(set(int(x) for x in batch.nx)) > 1:
        print('il %d k %d: ragged node counts %s, skipped' % (il, k, list(batch.nx))); continue
    ends = []
    try:
        for j in (0, 1, 2):
            lf = LocalForm(system_of(batch, j)); b = BdfIntegrator(lf)
            ends.append(lf.unpack(b.integrate([200.0])[-1])[0]); cold = b.stats.nst
        out = []
        for name, C0 in (('nearest', ends[0]), ('interp', 0.5 * ends[0] + 0.5 * ends[1])):
            lf = LocalForm(system_of(batch, 2)); b = BdfIntegrator(lf)
            C = lf.unpack(b.integrate([200.0], y0=lf.y_from_c(C0))[-1])[0]
            out.append('%s %d steps, vs cold %.1e' % (name, b.stats.nst, np.max(np.abs(C - ends[2])) / np.max(np.abs(ends[2]))))
        print('C4 pH idx %d, thickness idx %d (L=%.3g m) k=%d: cold %d steps; %s' % (ip, il + k // 2, v2[il + k // 2], k, cold, '; '.join(out)), flush=True)
    except Exception as e:
        print('il %d k %d: %s' % (il, k, e), flush=True)
