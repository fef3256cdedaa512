"""Statistics of a C4-type sweep (10 species, ragged 101/102 nodes): statuses, K1 residual of the converged
cells, wall-flux error, throughput.   python scripts/c4_props.py [n_pH] [n_L]"""
import os, sys, time, tempfile
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np, torch
from catint_b200 import backend as be, workloads
from catint_b200.transport import Transport
from catint_b200.calculator import build_cell_batch
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 64
t0 = time.time()
tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c4(n_pH=n1, n_L=n2))
batch, _ = build_cell_batch(tp)
print('host setup %.1f s, %d cells, b=%d' % (time.time() - t0, batch.B, batch.b))
bk = be.PnpBackend('cuda:0'); db = bk.upload(batch)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    out = bk.solve(db, [200.0], mode=be.MODE_STEADY, max_steps=20000)
    torch.cuda.synchronize(); dt = time.time() - t0
    print('solve %.3f s -> %.0f cells/s' % (dt, batch.B / dt))
status = out['status'].cpu().numpy(); ok = status == 0
print('status histogram', {int(k): int((status == k).sum()) for k in np.unique(status)})
S = batch.S
c = out['c'][-1].contiguous()
dcdt, _, _ = bk.rhs(db, c)
D = torch.tensor(batch.par[:, 2 * S:3 * S], device=c.device); dx = torch.tensor(batch.par[:, 3 * S + 5], device=c.device)
scale = (c.abs().amax(dim=1) * D / dx[:, None] ** 2).amax(dim=1)
ratio = (dcdt.abs().amax(dim=(1, 2)) / scale).cpu().numpy()
print('K1 residual ratio of converged cells: max %.3g median %.3g' % (ratio[ok].max(), np.median(ratio[ok])))
flux = out['flux'].cpu().numpy(); J = batch.par[:, S:2 * S]
print('flux err', np.max(np.abs(flux[ok] - J[ok])) / np.max(np.abs(J)))
cmin = c.amin(dim=(1, 2)).cpu().numpy(); print('min c over converged', cmin[ok].min())
dxs = batch.par[:, 3 * S + 5]
print('dx of failed cells: min %.3g max %.3g; of converged: min %.3g' % (dxs[~ok].min() if (~ok).any() else 0, dxs[~ok].max() if (~ok).any() else 0, dxs[ok].min()))
print('steps mean/max', float(out['n_steps'].double().mean()), int(out['n_steps'].max()))
