"""bounded run of a Stern / graded-mesh batch for compute-sanitizer:  python scripts/memcheck_case.py <nodes> <max_steps>"""
import os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np, torch
from catint_b200 import backend as be, workloads
from catint_b200.transport import Transport
from catint_b200.calculator import build_cell_batch
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
ms = int(sys.argv[2]) if len(sys.argv) > 2 else 30
tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c3(n_phi=2, n_pH=2))
batch, _ = build_cell_batch(tp, poisson_bc='stern', mesh=workloads.geometric_mesh(nn, 5e-11))
bk = be.PnpBackend('cuda:0'); db = bk.upload(batch)
out = bk.solve(db, [200.0], mode=be.MODE_STEADY, max_steps=ms)
torch.cuda.synchronize()
print('status', out['status'].tolist(), 'steps', out['n_steps'].tolist())
