import os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np, torch
from catint_b200 import backend as be, workloads
from catint_b200.transport import Transport
from catint_b200.calculator import build_cell_batch
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 201
go = dict(np.load(os.path.join(os.path.dirname(HERE), 'golden', 'oracle_c3_cells_n%d.npz' % nn)))
tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c3(n_phi=2, n_pH=2))
batch, _ = build_cell_batch(tp, poisson_bc='stern', mesh=workloads.geometric_mesh(nn, 5e-11))
bk = be.PnpBackend('cuda:0'); db = bk.upload(batch)
for prtol, iters in ((1e-10, 8), (1e-10, 30), (1e-8, 8)):
    out = bk.solve(db, [200.0], mode=be.MODE_STEADY, max_steps=50000, polish_rtol=prtol, polish_max_iter=iters)
    st = out['status'].tolist()
    errs = []
    for c in range(batch.B):
        if not bool(go['ok_%d' % c]): errs.append(None); continue
        cs = np.max(np.abs(batch.par[c, :8])); got = out['c'][-1, c].cpu().numpy(); ref = go['newton_c_%d' % c]
        errs.append(float(np.max(np.abs(got - ref) / (np.abs(ref) + 1e-12 * cs))))
    print('polish_rtol', prtol, 'iters', iters, 'status', st, 'steps', out['n_steps'].tolist(), 'relerr', errs)
