"""time the shards of an N-rank weak-scaling sweep one after the other on ONE GPU (which shard is the slow one, and why)
    python scripts/shard_check.py <world>"""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np, torch, bench
from catint_b200 import backend as be, distributed as D
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tp, gb = bench.c2_batch(n_cells=bench.CELLS_PER_GPU * world)
bk = be.PnpBackend('cuda:0')
for rank in range(world):
    b = gb.select(D.shard_indices(gb.B, rank, world))
    db = bk.upload(b); out = bk.alloc_outputs(db, 1)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        bk.solve(db, [bench.T_END], mode=be.MODE_STEADY, out=out); torch.cuda.synchronize(); dt = time.time() - t0
    st = out['n_steps'].cpu().numpy(); nn = out['n_newton'].cpu().numpy()
    print('rank %d: %.4f s  steps mean %.1f max %d  newton max %d  conv %d' % (rank, dt, st.mean(), st.max(), nn.max(), int((out['status'] == 0).sum())), flush=True)
