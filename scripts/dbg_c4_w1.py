import os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np, torch
from catint_b200 import backend as be, workloads
from catint_b200.transport import Transport
from catint_b200.calculator import build_cell_batch
tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c4(n_pH=4, n_L=4))
batch, _ = build_cell_batch(tp)
bk = be.PnpBackend('cuda:0')
for ms, use_order, keep, sel in ((5, True, False, None), (5, False, False, None), (200, True, True, None), (200, True, False, [1,2,3,5]), (200, True, False, [0,4,8]), (20000, True, False, None)):
    b = batch if sel is None else batch.select(sel)
    db = bk.upload(b)
    if not use_order:
        db.cells.order = None
    try:
        out = bk.solve(db, [200.0], mode=be.MODE_STEADY | (be.MODE_KEEP_ALL if keep else 0), max_steps=ms)
        torch.cuda.synchronize()
        print('ok   max_steps', ms, 'order', use_order, 'keep', keep, 'sel', sel, out['status'].tolist(), out['n_steps'].tolist(), flush=True)
    except Exception as e:
        print('FAIL max_steps', ms, 'order', use_order, 'keep', keep, 'sel', sel, str(e)[:100], flush=True)
        break
