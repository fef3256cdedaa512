#!/usr/bin/env python
"""bench.py -- converged 1D PNP cells/sec on the C2 workload (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one pass of the hot path over one batch: every rank integrates its
1024-cell CO2R/KHCO3 potential sweep (101 nodes, 8 species, buffer reactions,
migration) from the bulk state to t=200 s at scipy-odeint tolerances
(rtol=atol=1.49e-8) and Newton-polishes the steady state.  Weak scaling: the
per-GPU batch is fixed, rank r solves the same sweep shifted by r/(N*1024) V.

Printed JSON line (rank 0): value = converged cells / device time with the cell
parameters resident in HBM (CUDA events on the launching stream, L2 flushed
between steps, max over ranks); e2e = the same through Calculator.solve_batch
with host buffers (H2D of the parameters, solve, D2H of all results per step);
roofline of the dominant kernel (pnp_bdf_kernel); cpu_baseline = the oracle
(scipy odeint on the restated RHS) on the host cores.

--impl reference: the reference's CPU path (oracle port; the reference's own FD
integrator is orphaned Python 2 and cannot run) as a
multiprocessing sweep over the host cores, same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault('CATINT_QUIET', '1')

METRIC = 'converged 1D PNP cells/sec (batched sweep)'
CELLS_PER_GPU = 1024
T_END = 200.0
RTOL = ATOL = 1.49012e-8
WORKLOAD = ('C2: CO2R at Au in CO2-saturated KHCO3 (pH 6.8, 8 species, 5 buffer reactions, migration), '
            '50 um boundary layer, 101 nodes, %d-point potential sweep phiM=-0.5..-1.5 V with Tafel currents '
            '0.005..150 A/m^2, bulk state -> t=200 s at rtol=atol=1.49e-8 + Newton polish of the steady state'
            % CELLS_PER_GPU)


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum of one launch of the same workload), or None."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.isfile(p):
        d = json.load(open(p)).get(kernel)
        if d:
            return float(d['dram_bytes_read']) + float(d['dram_bytes_write'])
    return None


def c2_batch(rank=0, world=1, n_cells=CELLS_PER_GPU):
    """the rank's sweep as a host CellBatch (model tables from the fixture-verified Transport)."""
    import tempfile
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    from catint_b200 import workloads
    kw = workloads.c2(n_potentials=n_cells)
    shift = (rank / float(max(world, 1))) * (1.0 / n_cells) if world > 1 else 0.0
    kw['descriptors'] = {'phiM': [p - shift for p in kw['descriptors']['phiM']]}
    tp = Transport(resultsdir=tempfile.mkdtemp(prefix='catint_bench_'), model_name='bench', **kw)
    tp.set_calculator('odeint')
    batch, _ = build_cell_batch(tp)
    return tp, batch


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(smax)), 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------
# CPU side: the oracle as a multiprocessing sweep
# ---------------------------------------------------------------------------
def _cpu_cell(args):
    par, nx, z, reactions, nu_mode, t_end = args
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle.pnp_oracle import PnpSystem, steady_tmesh
    S = len(z)
    x = np.arange(int(nx)) * par[3 * S + 5]
    s = PnpSystem(z=z, D=par[2 * S:3 * S], c_bulk=par[0:S], J=par[S:2 * S], x=x, beta=par[3 * S], eps=par[3 * S + 1],
                  reactions=reactions, rate_mode=nu_mode, use_migration=True,
                  phi_wall=par[3 * S + 2], g_bulk=par[3 * S + 3], uniform=True)
    t0 = time.time()
    sol, info = s.integrate_odeint(steady_tmesh(t_end), full_output=True)
    ok = bool(np.all(np.isfinite(sol[-1])) and info['message'] == 'Integration successful.')
    return ok, time.time() - t0, int(info['nfe'][-1])


def cpu_sweep(batch, n_cells, cores, seed=0):
    """odeint oracle on a seed-0 random subsample of the batch, one process per core."""
    import multiprocessing as mp
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(batch.B, size=n_cells, replace=False))
    jobs = [(batch.par[c], batch.nx[c], batch.z, batch.reactions, 'summed', T_END) for c in pick]
    ctx = mp.get_context('fork')
    t0 = time.time()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_cell, jobs, chunksize=1)
    wall = time.time() - t0
    n_ok = sum(1 for r in res if r[0])
    return n_ok / wall, wall, n_ok, pick


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    tp, batch = c2_batch()
    n_cells = min(cores, batch.B)
    budget = float(os.environ.get('CATINT_REF_BUDGET_S', '240'))
    t_start = time.time()
    times, values = [], []
    warm = 0
    # one warm-up step tells the step time; the rest of W/K is clamped to the wall-clock budget
    if args.warmup > 0:
        v, wall, n_ok, _ = cpu_sweep(batch, n_cells, cores)
        warm = 1
        step_t = wall
    else:
        step_t = 70.0
    k_eff = max(1, min(args.steps, int((budget - (time.time() - t_start)) // max(step_t, 1.0))))
    for k in range(k_eff):
        v, wall, n_ok, _ = cpu_sweep(batch, n_cells, cores, seed=k)
        times.append(wall); values.append(v)
    value = float(np.sum([v * t for v, t in zip(values, times)]) / np.sum(times))
    sample = ('%d seed-k random cells of the 1024-cell sweep per step (one per core), scipy odeint on the '
              'numpy-vectorised restated RHS, dense FD Jacobian, t_end=200 s, default rtol/atol; '
              'requested steps/warmup %d/%d clamped to a %d s budget' % (n_cells, args.steps, args.warmup, int(budget)))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'cells/s', 'n_gpus': args.gpus,
        'steps': k_eff, 'warmup': warm, 'ms_per_step': 1e3 * float(np.mean(times)), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'cells_per_step': n_cells},
        'cpu_baseline': {'value': value, 'unit': 'cells/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'cells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def rhs_roofline(bk, batch, dev, n_cells=131072, reps=10):
    """K1 (pnp_rhs_kernel) alone on a batch larger than L2: achieved HBM bandwidth of the streaming
    RHS stage, algorithmic bytes 16*S*n per cell (read c, write dc/dt)."""
    import torch
    from catint_b200 import workloads
    big = workloads.replicate_batch(batch, n_cells)
    db = bk.upload(big)
    S, n = big.S, big.nx_max
    c = torch.empty((n_cells, n, S), dtype=torch.float64, device=dev)
    c[:] = torch.tensor(big.par[0, :S], device=dev)[None, None, :]
    c *= 1.0 + 0.01 * torch.rand_like(c)
    lib = bk.lib
    import ctypes
    dcdt = torch.empty_like(c)
    def launch():
        rc = lib.catint_pnp_rhs_batch(ctypes.byref(db.shared), ctypes.byref(db.cells), n_cells, c.data_ptr(),
                                      dcdt.data_ptr(), None, None,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = 16.0 * S * n * n_cells
    peak, which = measured_peaks()
    ach = nbytes / (ms * 1e-3) / 1e9
    return {'bound': 'hbm', 'kernel': 'pnp_rhs_kernel', 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
            'frac': ach / peak, 'traffic': ncu_traffic('pnp_rhs_kernel'), 'peak_source': which, 'cells': n_cells,
            'algorithmic_bytes_per_launch': nbytes, 'ms_per_launch': ms,
            'note': 'input+output 2x%.0f MB per launch (> 126 MB L2), back-to-back launches' % (nbytes / 2e6)}


# ---------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from catint_b200 import backend as be
    from catint_b200.calculator import Calculator

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))

    tp, batch = c2_batch(rank, world)
    bk = be.PnpBackend(dev)
    db = bk.upload(batch)
    out = bk.alloc_outputs(db, 1)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def one_step():
        flush.fill_(1)                                   # evict L2 (outside the timed events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bk.solve(db, [T_END], mode=be.MODE_STEADY, rtol=RTOL, atol=ATOL, out=out)
        e1.record()
        return e0, e1

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = bk.launches
    evs = [one_step() for _ in range(args.steps)]
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    launches = bk.launches - launches0
    dev_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    n_conv = int((out['status'] == 0).sum().item())
    n_newton = out['n_newton'].to(torch.float64)
    newton_total = float(n_newton.sum().item())
    steps_mean = float(out['n_steps'].to(torch.float64).mean().item())

    # ---- e2e: host buffers in, host buffers out, every step -----------------------
    calc = Calculator(transport=tp, dt=0.5, tmax=T_END, ntout=1, mode='stationary', rtol=RTOL, atol=ATOL, device=dev)
    pinned = {'par': torch.from_numpy(batch.par).pin_memory(), 'nx': torch.from_numpy(batch.nx).pin_memory()}
    e2e_steps = max(1, min(args.steps, 3))
    res = calc.solve_batch(batch, backend=bk, pinned=pinned)          # warm
    sync_all()
    t0 = time.perf_counter()
    e2e_conv = 0
    for _ in range(e2e_steps):
        res = calc.solve_batch(batch, backend=bk, pinned=pinned)
        e2e_conv += int(np.sum(res['status'] == 0))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d, d2h = int(res['h2d_bytes']), int(res['d2h_bytes'])

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
        c = torch.tensor([n_conv, e2e_conv, newton_total], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        n_conv_all, e2e_conv_all, newton_all = float(c[0]), float(c[1]), float(c[2])
    else:
        n_conv_all, e2e_conv_all, newton_all = float(n_conv), float(e2e_conv), newton_total

    if rank == 0:
        value = n_conv_all * args.steps / (dev_ms * 1e-3)
        e2e_value = e2e_conv_all / e2e_s
        # roofline of pnp_bdf_kernel (one launch per step and rank).  Algorithmic bytes:
        #   per Newton iteration  16*S*n (state in, update out) + 2*8*n*REC (the stored block factors are
        #                         streamed once by the forward and once by the backward solve sweep)
        #   per factorisation     8*n*REC written (REC = b*bp + 4*b doubles per node, bp = b rounded up to even)
        #   per BDF step          26*8*n*b (Nordsieck history: predict pass reads/writes 6 vectors, correction
        #                         pass reads 7 and writes 7)
        S, n, b = batch.S, int(batch.nx_max), batch.b
        rec = b * (b + (b & 1)) + 4 * b
        bytes_newton = 16.0 * S * n + 16.0 * n * rec
        bytes_factor = 8.0 * n * rec
        bytes_step = 26.0 * 8.0 * n * b
        setups_total = float(out['n_setups'].to(torch.float64).sum().item())
        steps_total = float(out['n_steps'].to(torch.float64).sum().item())
        algo_bytes = newton_total * bytes_newton + setups_total * bytes_factor + steps_total * bytes_step
        flops = (setups_total * n * (8.0 / 3.0 * b ** 3 + 4.0 * b ** 2)        # block inverse + W column + Schur
                 + newton_total * n * (4.0 * b ** 2 + 40.0 * S))             # two mat-vecs + residual
        launch_s = dev_ms * 1e-3 / args.steps
        peak, which = measured_peaks()
        achieved = algo_bytes / launch_s / 1e9                                  # this rank's launch
        roofline = {'bound': 'hbm', 'kernel': 'pnp_bdf_kernel<9,false,true,false>', 'achieved': achieved, 'peak': peak,
                    'unit': 'GB/s', 'frac': achieved / peak, 'traffic': ncu_traffic('pnp_bdf_kernel'), 'peak_source': which,
                    'algorithmic_bytes_per_launch': algo_bytes,
                    'newton_iterations_per_launch': newton_total, 'factorisations_per_launch': setups_total,
                    'bdf_steps_per_launch': steps_total,
                    'bytes_per_newton_iteration': bytes_newton, 'bytes_per_factorisation': bytes_factor,
                    'bytes_per_bdf_step': bytes_step,
                    'fp64_tflops_algorithmic': flops / launch_s / 1e12,
                    'note': 'one warp per cell, 1024 cells = 7 warps/SM: the kernel is bound by the dependency '
                            'latency of the sequential block sweeps, not by HBM; traffic = DRAM bytes of one launch '
                            '(ncu, L2 absorbs the rest of the algorithmic bytes), see DESIGN.md 6'}
        line = {
            'metric': METRIC, 'value': value, 'unit': 'cells/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'cells_per_gpu': CELLS_PER_GPU,
                       'l2': 'flushed between timed steps (256 MiB device write)',
                       'mean_bdf_steps_per_cell': steps_mean, 'mean_newton_per_cell': newton_total / batch.B,
                       'mean_factorisations_per_cell': setups_total / batch.B},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'cells/s', 'h2d_bytes_per_step': h2d * args.gpus,
                    'd2h_bytes_per_step': d2h * args.gpus, 'steps': e2e_steps},
            'gpu_launches': launches * args.gpus,
            'converged_cells_per_step': n_conv_all,
            'roofline': roofline,
        }
        if args.gpus == 1:
            line['roofline_rhs'] = rhs_roofline(bk, batch, dev)
        if args.gpus == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_cells = min(cores, batch.B)
            v, wall, n_ok, pick = cpu_sweep(batch, n_cells, cores)
            line['cpu_baseline'] = {
                'value': v, 'unit': 'cells/s', 'cores': cores, 'kind': 'port',
                'sample': '%d seed-0 random cells of the same 1024-cell sweep, one per core, scipy odeint on the '
                          'numpy-vectorised restated RHS (dense FD Jacobian, t_end=200 s, default rtol/atol), '
                          '%.1f s wall, %d converged' % (n_cells, wall, n_ok)}
        else:
            line['cpu_baseline'] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    import __graft_entry__ as g
    if not os.path.isfile(g.LIB):
        g.build()
    return run_gpu(args)


if __name__ == '__main__':
    sys.exit(main())
