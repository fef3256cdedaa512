"""debug: cell (pH 6.0, L = 10 um) of the C4 sweep alone / in the 4x4 batch / in the big batch, with different
workspace pre-fills (is any result batch- or garbage-dependent?)"""
import os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np, torch
from catint_b200 import backend as be, workloads
from catint_b200.transport import Transport
from catint_b200.calculator import build_cell_batch

def batch_of(npH, nL):
    tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c4(n_pH=npH, n_L=nL))
    return build_cell_batch(tp)[0]

bk = be.PnpBackend('cuda:0')
small = batch_of(4, 4)
big = batch_of(256, 256)
print('par bits equal:', np.array_equal(small.par[0], big.par[0]), 'nx', small.nx[0], big.nx[0], 'nx_max', small.nx_max, big.nx_max)
def run(batch, fill, label):
    db = bk.upload(batch)
    ws, need = bk.workspace(db)
    if fill == 'zero': ws.zero_()
    elif fill == 'nan': ws.fill_(255)
    out = bk.solve(db, [200.0], mode=be.MODE_STEADY, max_steps=20000)
    torch.cuda.synchronize()
    st = out['status'].cpu().numpy(); ns = out['n_steps'].cpu().numpy()
    print('%-28s cell0 status %d steps %d | fails %d of %d' % (label, st[0], ns[0], int((st != 0).sum()), len(st)))
    return st
run(small, 'zero', '4x4 zero ws')
run(small, 'nan', '4x4 nan ws')
run(small.select([0]), 'nan', 'cell 0 alone nan ws')
sub = big.select(np.arange(0, 4096))
a = run(sub, 'zero', 'first 16 pH rows zero ws')
b = run(sub, 'nan', 'first 16 pH rows nan ws')
print('same statuses zero/nan:', np.array_equal(a, b))
one = big.select([0, 1, 2, 3])
run(one, 'nan', 'big cells 0..3 nan ws')
g = (a == 0).reshape(16, 256)
for i in range(0, 16, 3): print('row', i, ''.join('1' if v else '0' for v in g[i, :40]))
