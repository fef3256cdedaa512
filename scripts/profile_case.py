"""Short, profiler-friendly invocation of the hot path: the 1024-cell C2 batch integrated for a
bounded number of BDF steps (enough to be representative, short enough for ncu's replays).
    python scripts/profile_case.py [max_steps] [n_cells]
"""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np
import torch
import bench
from catint_b200 import backend as be

def main():
    max_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    tp, batch = bench.c2_batch(n_cells=n_cells)
    bk = be.PnpBackend('cuda:0')
    db = bk.upload(batch)
    out = bk.alloc_outputs(db, 1)
    prof = torch.zeros((n_cells, 8), dtype=torch.int64, device='cuda:0')
    if os.environ.get('CATINT_PHASES'):
        bk.lib.catint_pnp_debug_profile_buffer(prof.data_ptr())
    for rep in range(2):
        prof.zero_()
        torch.cuda.synchronize(); t0 = time.time()
        bk.solve(db, [bench.T_END], mode=be.MODE_STEADY | (be.MODE_KEEP_ALL if os.environ.get('CATINT_KEEP_ALL') else 0), max_steps=max_steps, out=out)
        torch.cuda.synchronize(); dt = time.time() - t0
        nn = float(out['n_newton'].double().sum())
        ns = float(out['n_setups'].double().sum()); st = float(out['n_steps'].double().sum())
        print('   steps/cell %.1f newton/cell %.1f setups/cell %.1f status0 %d' % (st / n_cells, nn / n_cells, ns / n_cells, int((out['status'] == 0).sum())))
        print('rep %d: %.4f s, %d cells, newton total %.0f -> %.1f us per newton iteration per cell-warp, %.2f us amortised'
              % (rep, dt, n_cells, nn, dt / (nn / n_cells) * 1e6, dt / nn * 1e6))

    stp = out['n_steps'].cpu().numpy(); nw = out['n_newton'].cpu().numpy(); sts = out['n_setups'].cpu().numpy()
    print('steps  min/mean/max', stp.min(), stp.mean(), stp.max(), ' newton', nw.min(), nw.mean(), nw.max(), ' setups', sts.min(), sts.mean(), sts.max())
    print('steps by cell decile:', [int(stp[i]) for i in range(0, n_cells, max(1, n_cells // 10))])
    if os.environ.get('CATINT_PHASES'):
        tot = prof[:, 7].double().cpu().numpy(); print('total cycles min/mean/max %.3g %.3g %.3g' % (tot.min(), tot.mean(), tot.max()))
        j = int(tot.argmax()); pj = prof[j].cpu().numpy()
        print('slowest cell %d: steps %d newton %d setups %d cycles %s' % (j, stp[j], nw[j], sts[j], [int(v) for v in pj]))
        pr = prof.double().mean(dim=0).cpu().numpy()
        names = ['factor', 'residual', 'forward', 'backward', 'assembly(in factor)', 'history', 'correction', 'total']
        print('mean cycles per cell:', {k: '%.3g (%.1f%%)' % (v, 100 * v / pr[7]) for k, v in zip(names, pr)})
        print('per call: factor %.0f cyc (assembly %.0f), residual %.0f, forward %.0f, backward %.0f (per newton), history %.0f, correction %.0f (per step)' % (pr[0] / (ns / n_cells), pr[4] / (ns / n_cells), pr[1] / (nn / n_cells), pr[2] / (nn / n_cells), pr[3] / (nn / n_cells), pr[5] / (st / n_cells), pr[6] / (st / n_cells)))

if __name__ == '__main__':
    main()
