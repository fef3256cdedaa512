"""BASELINE.json configs C3 / C4 / C5 at (a share of) their stated sizes on one GPU: launch time, converged
cells/s, status histogram, step statistics, per-launch algorithmic bytes / flops of SURVEY 8(d).
    python scripts/size_run.py c3|c4|c5|c5u [n_cells_or_stride] [max_steps]
The second argument takes every k-th cell of the full grid (k = stride, e.g. 8 = the share of one rank of an
8-GPU run, which is exactly the round-robin shard of catint_b200/distributed.py).
"""
import json, os, sys, tempfile, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np
import torch
from catint_b200 import backend as be, workloads
from catint_b200.transport import Transport
from catint_b200.calculator import build_cell_batch


def make(cfg):
    d = tempfile.mkdtemp(prefix='catint_size_')
    if cfg == 'c3':
        tp = Transport(resultsdir=d, **workloads.c3())
        batch, _ = build_cell_batch(tp, poisson_bc='stern', mesh=workloads.geometric_mesh(1001, 5e-11))
        return batch, [200.0], be.MODE_STEADY
    if cfg == 'c4':
        tp = Transport(resultsdir=d, **workloads.c4())
        batch, _ = build_cell_batch(tp)
        return batch, [200.0], be.MODE_STEADY
    if cfg == 'c5':
        tp = Transport(resultsdir=d, **workloads.c5())
        batch, _ = build_cell_batch(tp, mesh=workloads.geometric_mesh(5001, workloads.C5_FIRST_SPACING))
        return batch, workloads.C5_T_OUT, be.MODE_TRANSIENT
    if cfg == 'c5u':
        tp = Transport(resultsdir=d, **workloads.c5())
        batch, _ = build_cell_batch(tp)
        return batch, workloads.C5_T_OUT, be.MODE_TRANSIENT
    raise SystemExit('unknown config ' + cfg)


def main():
    cfg = sys.argv[1]
    stride = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    max_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
    t0 = time.time()
    batch, t_out, mode = make(cfg)
    if stride > 1:
        batch = batch.select(np.arange(0, batch.B, stride))
    t_host = time.time() - t0
    bk = be.PnpBackend('cuda:0')
    db = bk.upload(batch)
    out = bk.alloc_outputs(db, len(t_out))
    ws, need = bk.workspace(db)
    prof = None
    if os.environ.get('CATINT_PHASES'):
        prof = torch.zeros((batch.B, 8), dtype=torch.int64, device='cuda:0')
        bk.lib.catint_pnp_debug_profile_buffer(prof.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bk.solve(db, t_out, mode=mode, max_steps=max_steps, out=out)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    status = out['status'].cpu().numpy()
    hist = {be.CELL_STATUS.get(int(k), int(k)): int(v) for k, v in zip(*np.unique(status, return_counts=True))}
    st = out['n_steps'].cpu().numpy().astype(float); nn = out['n_newton'].cpu().numpy().astype(float)
    ns = out['n_setups'].cpu().numpy().astype(float)
    S, n, b = batch.S, int(batch.nx_max), batch.b
    bytes_newton = 16.0 * S * n + (16.0 * b * b * n if 8.0 * b * b * n > 227e3 else 0.0)
    flops_newton = n * (14.0 / 3.0 * b ** 3 + 4.0 * b ** 2)
    line = {'config': cfg, 'cells': int(batch.B), 'stride': stride, 'nodes': n, 'species': S, 'block': b,
            'seconds': sec, 'converged': int((status == 0).sum()), 'cells_per_s': float((status == 0).sum() / sec),
            'status': hist, 'steps_mean': float(st.mean()), 'steps_max': float(st.max()),
            'newton_mean': float(nn.mean()), 'setups_mean': float(ns.mean()), 'newton_total': float(nn.sum()),
            'setups_total': float(ns.sum()),
            'workspace_GB': need / 1e9, 'host_build_s': t_host,
            'hbm_GBs_8d': float(nn.sum() * bytes_newton / sec / 1e9),
            'fp64_TFs_8d': float((ns.sum() * n * (14.0 / 3.0 * b ** 3) + nn.sum() * n * 4.0 * b * b) / sec / 1e12),
            'fp64_TFs_8d_dense_per_newton': float(nn.sum() * flops_newton / sec / 1e12)}
    print(json.dumps(line), flush=True)
    if prof is not None:
        pr = prof.double().mean(dim=0).cpu().numpy()
        names = ['factor', 'residual', 'forward', 'backward', 'assembly(in factor)', 'history', 'correction', 'total']
        print('mean cycles per cell:', {k: '%.3g (%.1f%%)' % (v, 100 * v / pr[7]) for k, v in zip(names, pr)})
        print('per call: factor %.0f (assembly %.0f), residual %.0f, forward %.0f, backward %.0f (per newton), history %.0f, '
              'correction %.0f (per step); per node and sweep: fwd %.0f bwd %.0f cycles' % (
                  pr[0] / ns.mean(), pr[4] / ns.mean(), pr[1] / nn.mean(), pr[2] / nn.mean(), pr[3] / nn.mean(),
                  pr[5] / st.mean(), pr[6] / st.mean(), pr[2] / nn.mean() / (n / 2), pr[3] / nn.mean() / (n / 2)))
    if cfg.startswith('c5') or cfg == 'c3':
        bad = np.nonzero(status != 0)[0][:20]
        print('first failing cells', bad.tolist(), [int(s) for s in status[bad]])


if __name__ == '__main__':
    main()
