"""CPU probe (oracle BDF = prototype of the CUDA integrator): BDF steps of a warm cell of the C2 sweep started from the
nearest cold neighbour state against the linear interpolation of the two bracketing cold cells, and the distance of
both end states from the cold run.   NCELLS=16384 K=16 OMP_NUM_THREADS=1 python tests/tools/continuation_interp_probe.py 500 4000 8000
Test infrastructure (uses oracle/)."""
import os, sys, tempfile, time
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np
from catint_b200 import workloads
from catint_b200.transport import Transport
from catint_b200.calculator import build_cell_batch
from continuation_axis_probe import system_of
from oracle.pnp_local import LocalForm
from oracle.bdf_local import BdfIntegrator
n = int(os.environ.get('NCELLS', '1024')); k = int(os.environ.get('K', '16'))
tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c2(n_potentials=n)); tp.set_calculator('odeint')
keys = list(tp.descriptors); v1, v2 = tp.descriptors[keys[0]], tp.descriptors[keys[1]]
for base in [int(a) for a in sys.argv[1:]]:
    idx = [base, base + k, base + k // 2]
    batch, _ = build_cell_batch(tp, points=[{keys[0]: float(v1[i]), keys[1]: float(v2[0])} for i in idx])
    ends = []; 
    for j in (0, 1, 2):
        lf = LocalForm(system_of(batch, j)); b = BdfIntegrator(lf)
        ends.append(lf.unpack(b.integrate([200.0])[-1])[0]); cold = b.stats.nst
    out = []
    for name, C0 in (('nearest', ends[0]), ('interp', 0.5 * ends[0] + 0.5 * ends[1])):
        lf = LocalForm(system_of(batch, 2)); b = BdfIntegrator(lf)
        C = lf.unpack(b.integrate([200.0], y0=lf.y_from_c(C0))[-1])[0]
        out.append('%s %d steps, vs cold %.1e' % (name, b.stats.nst, np.max(np.abs(C - ends[2])) / np.max(np.abs(ends[2]))))
    print('n=%d k=%d base %d (phiM %.3f): cold %d steps; %s' % (n, k, base, v1[base + k // 2], cold, '; '.join(out)), flush=True)
