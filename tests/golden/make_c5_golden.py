"""Goldens for C5 (BASELINE.json configs[4]): time-dependent transient from the bulk state to the steady
state with adaptive dt on the 5001-node graded mesh (geometric, first interval 0.05 nm; SURVEY 8d allows
"uniform or graded"), CO2R/KHCO3 chemistry, default Poisson boundary (block size 9), 64 phiM x 64 flux
scalings (descriptor 'RF').  Six cells of the 4096-cell grid: the four corners and two interior points.

    OMP_NUM_THREADS=1 python tests/golden/make_c5_golden.py [n_nodes]

Dense odeint is out of reach at N = 40 008; the checker is the CPU BDF of oracle/bdf_local.py (pinned against
scipy odeint at 101 nodes: tests/test_oracle.py, make_c3_golden.py 101) at the same rtol/atol, plus the Newton
root of the discrete residual for the end state.  Every cell is integrated twice, the second time from an
initial state perturbed by 1e-13 relative, which measures the conditioning of every output time (see
make_n1001_golden.py); the test takes its tolerance from that measurement.

`uniform_probe_*` records what the same CPU BDF does on the UNIFORM 5001-node mesh (dx = 10 nm) for two corner
cells (round 1 had dropped C5 on the unverified claim that this case has no bounded solution: it has one; the
cells that do blow up are the thin LAYERS of C4, L = 10 um, proven with odeint in make_c4_golden.py).
"""
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
os.environ.setdefault('CATINT_QUIET', '1')

CELLS = [0, 63, 4032, 4095, 2015, 3023]          # (phi index, scale index) = divmod(cell, 64)


def make_batch(nn, uniform=False):
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    from catint_b200 import workloads
    tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.c5(nx=nn - 1))
    keys = list(tp.descriptors)
    pts = []
    for c in CELLS:
        i1, i2 = divmod(c, 64)
        pts.append({keys[0]: float(tp.descriptors[keys[0]][i1]), keys[1]: float(tp.descriptors[keys[1]][i2])})
    mesh = None if uniform else workloads.geometric_mesh(nn, workloads.C5_FIRST_SPACING)
    batch, _ = build_cell_batch(tp, points=pts, mesh=mesh)
    return batch


def work(job):
    nn, k, uniform = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from catint_b200 import workloads
    from oracle.fixtures import system_from_batch
    from oracle.pnp_local import LocalForm
    from oracle.bdf_local import BdfIntegrator
    batch = make_batch(nn, uniform)
    s = system_from_batch(batch, k)
    lf = LocalForm(s)
    t_out = list(workloads.C5_T_OUT)
    res = {'k': k}
    rng = np.random.default_rng(k)
    runs = []
    for pert in ((0.0,) if uniform else (0.0, 1e-13)):
        y0 = lf.y_bulk()
        if pert:
            C = y0[:, :s.S].T * (1.0 + pert * rng.standard_normal((s.S, s.n)))
            C[:, -1] = s.c_bulk
            y0 = lf.y_from_c(C)
        t0 = time.time()
        integ = BdfIntegrator(lf, fresh_jacobian=False)
        try:
            outs = integ.integrate(t_out, y0=y0, max_steps=400000)
            msg = 'reached t=%g s in %d steps' % (t_out[-1], integ.stats.nst)
            ok = bool(all(np.all(np.isfinite(o)) for o in outs))
        except RuntimeError as e:
            ok, msg, outs = False, str(e), None
            res['fail_t'] = float(integ.t)
        print('%s n %d cell %d pert %.0e: %s (%.0f s)' % ('uniform' if uniform else 'graded', nn, CELLS[k], pert, msg,
                                                          time.time() - t0), flush=True)
        if pert == 0.0:
            res.update(ok=ok, msg=msg, steps=integ.stats.nst)
            if not ok:
                return res
            y_end = outs[-1]
        if ok:
            runs.append(np.stack([o[:, :s.S] for o in outs]))
    if uniform:
        return res
    cs = np.max(np.abs(s.c_bulk))
    # transient states are stored on a node subset (all of the first 200 wall nodes, then every 25th): small fixture
    keep = np.unique(np.concatenate([np.arange(0, min(200, s.n)), np.arange(0, s.n, 25), [s.n - 1]]))
    res['nodes_kept'] = keep
    res['bdf_c'] = runs[0][:, keep, :]
    if len(runs) == 2:
        # relative difference of the two runs; floor 1e-9*c_bulk (like the steady-state tests) and 1e-6*c_bulk (what
        # the transient test uses: species that start from zero are compared absolutely while they are still tiny)
        for name, floor in (('sensitivity', 1e-9), ('sensitivity_f6', 1e-6)):
            res[name] = np.array([float(np.max(np.abs(runs[0][j] - runs[1][j]) / (np.abs(runs[0][j]) + floor * cs)))
                                  for j in range(len(t_out))])
    y, info = lf.solve_steady(y0=y_end, pure_newton=True)
    res['newton_ok'] = bool(info['converged'])
    C, v, g = lf.unpack(y)
    res.update(newton_c=C.T.copy(), phi=v, g=g)
    res['end_vs_root'] = float(np.max(np.abs(runs[0][-1] - C.T) / (np.abs(C.T) + 1e-9 * cs)))
    return res


def main():
    nn = int(sys.argv[1]) if len(sys.argv) > 1 else 5001
    from catint_b200 import workloads
    batch = make_batch(nn)
    jobs = [(nn, k, False) for k in range(len(CELLS))] + [(nn, 0, True), (nn, 3, True)]
    with mp.Pool(min(len(jobs), os.cpu_count())) as pool:
        results = pool.map(work, jobs, chunksize=1)
    out = dict(cells=np.array(CELLS), par=batch.par, mesh=batch.mesh_xi[0], nodes=np.array(nn),
               t_out=np.array(workloads.C5_T_OUT))
    for (nn_, k, uniform), r in zip(jobs, results):
        if uniform:
            out['uniform_probe_ok_%d' % CELLS[k]] = np.array(r['ok'])
            out['uniform_probe_msg_%d' % CELLS[k]] = np.array(r['msg'])
            continue
        c = CELLS[k]
        for key in ('ok', 'msg', 'steps', 'fail_t', 'bdf_c', 'nodes_kept', 'sensitivity', 'sensitivity_f6', 'newton_c', 'phi', 'g', 'newton_ok', 'end_vs_root'):
            if key in r and r[key] is not None:
                out['%s_%d' % (key, c)] = np.array(r[key])
        if r.get('ok'):
            print('cell %d: steps %d, sensitivity %s (floor 1e-6: %s), end state vs Newton root %.2e' % (
                c, r['steps'], np.array2string(r['sensitivity'], precision=2),
                np.array2string(r['sensitivity_f6'], precision=2), r['end_vs_root']))
    np.savez_compressed(os.path.join(HERE, 'oracle_c5_cells_n%d.npz' % nn), **out)


if __name__ == '__main__':
    main()
