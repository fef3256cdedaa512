import os, sys, ctypes
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CATINT_QUIET', '1')
import numpy as np, torch, bench
from catint_b200 import backend as be, workloads
tp, batch = bench.c2_batch(n_cells=64)
bk = be.PnpBackend('cuda:0')
for n_cells in (8192, 131072):
    big = workloads.replicate_batch(batch, n_cells); db = bk.upload(big)
    S, n = big.S, big.nx_max
    c = torch.empty((n_cells, n, S), dtype=torch.float64, device='cuda:0'); c[:] = torch.tensor(big.par[0, :S], device='cuda:0')[None, None, :]
    dcdt = torch.empty_like(c)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def launch():
        assert bk.lib.catint_pnp_rhs_batch(ctypes.byref(db.shared), ctypes.byref(db.cells), n_cells, c.data_ptr(), dcdt.data_ptr(), None, None, st) == 0
    for _ in range(3): launch()
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    nb = 16.0 * S * n * n_cells
    print(n_cells, 'ms', [round(t, 3) for t in ts], 'best GB/s %.0f' % (nb / (min(ts) * 1e-3) / 1e9))
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); dcdt.copy_(c); t1.record(); torch.cuda.synchronize(); print('  torch copy same size: %.3f ms' % t0.elapsed_time(t1))
