"""torchrun check of the sharded Calculator.run(): every rank must end with the full, ordered result,
identical to a single-rank solve.   python -m torch.distributed.run --nproc-per-node 2 scripts/dist_check.py"""
import os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np
import torch
import torch.distributed as dist
from catint_b200 import workloads, distributed as D
from catint_b200.transport import Transport
from catint_b200.calculator import Calculator, build_cell_batch

def main():
    rank = int(os.environ['RANK']); local = int(os.environ['LOCAL_RANK']); world = int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda:%d' % local))
    kw = workloads.c2(n_potentials=21, phi_min=-0.7, phi_max=-1.2)
    tp = Transport(resultsdir=tempfile.mkdtemp(prefix='dist_r%d_' % rank), **kw)     # rank 0's folder is broadcast
    tp.set_calculator('odeint')
    calc = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary', device='cuda:%d' % local)
    res = calc.run()
    batch, _ = build_cell_batch(tp)
    ref = calc.solve_batch(batch)                      # all cells on this rank alone
    ok = all(np.array_equal(res[k], ref[k]) for k in ('c', 'phi', 'g', 'flux', 'status', 'n_steps', 'n_newton', 'n_setups'))
    ok = ok and res['c'].dtype == ref['c'].dtype and res['status'].dtype == ref['status'].dtype
    print('rank %d/%d: sharded == single-rank: %s, converged %d/%d, CO2(0) first/last %.6f %.6f' % (
        rank, world, ok, int((res['status'] == 0).sum()), batch.B,
        tp.alldata[0]['species']['CO2']['surface_concentration'], tp.alldata[-1]['species']['CO2']['surface_concentration']), flush=True)
    # continuation run (two sharded waves, wave 1 gathered and kept on the devices): the plain run's steady states
    cont = Calculator(transport=tp, dt=0.5, tmax=200, ntout=1, mode='stationary', device='cuda:%d' % local, continuation=4)
    rc = cont.run_continuation(batch)
    scale = np.max(np.abs(ref['c'][-1]), axis=(1, 2), keepdims=True)
    devn = float(np.max(np.abs(rc['c'][-1] - ref['c'][-1]) / scale))
    okc = devn < 1e-6 and bool(np.all(rc['status'] == 0)) and rc['status'].dtype == ref['status'].dtype \
        and float(np.max(np.abs(rc['flux'] - ref['flux']))) < 1e-6 * float(np.max(np.abs(ref['flux'])))
    st = cont.continuation_stats
    print('rank %d/%d: continuation == plain: %s (worst deviation %.1e, cold %d cells %.0f steps, warm %d cells %.0f steps)'
          % (rank, world, okc, devn, st['cold_cells'], st['cold_steps_mean'], st['warm_cells'], st['warm_steps_mean']), flush=True)
    ok = ok and okc
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)

if __name__ == '__main__':
    main()
