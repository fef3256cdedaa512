"""where in time does an outlier cell of the weak-scaling sweep spend its extra steps?"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np, torch, bench
from catint_b200 import backend as be, distributed as D
tp, gb = bench.c2_batch(n_cells=8192)
bk = be.PnpBackend('cuda:0')
b = gb.select(D.shard_indices(gb.B, 2, 8))
out = bk.solve(bk.upload(b), [200.0], mode=be.MODE_STEADY)
st = out['n_steps'].cpu().numpy(); nn = out['n_newton'].cpu().numpy(); ns = out['n_setups'].cpu().numpy()
worst = np.argsort(-st)[:6]
print('worst cells of shard 2:', [(int(j), int(st[j]), int(nn[j]), int(ns[j])) for j in worst])
j = int(worst[0])
S = b.S
print('flux CO of cells j-2..j+2 (A/m2):', [float(b.par[k, S + 3] * 2 * 96485.33) for k in range(j - 2, j + 3)], 'steps', [int(st[k]) for k in range(j - 2, j + 3)])
sub = b.select([j - 1, j, j + 1])
db = bk.upload(sub)
for te in (1e-8, 1e-6, 1e-4, 1e-3, 1e-2, 1e-1, 1.0, 10.0, 50.0, 200.0):
    o = bk.solve(db, [te], mode=be.MODE_TRANSIENT)
    print('t_end %8.1e steps %s newton %s setups %s' % (te, o['n_steps'].tolist(), o['n_newton'].tolist(), o['n_setups'].tolist()))
