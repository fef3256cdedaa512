"""Sharding of a cell batch over the ranks of torch.distributed.

The reference's only parallel strategy is data parallelism over the descriptor
grid with a static round-robin map ``itask % mpi_size == mpi_rank``
(/root/reference/catint/calculator.py:209-212) and a gather of the result
dictionaries (/root/reference/catint/catint_io.py:154-178; disabled upstream).
Here: one process per GPU, cell j -> rank j % world_size, no traffic while the
cells integrate, and ONE collective at the end: every rank packs its results
(concentrations, potential, gradient, wall fluxes, status and counters) into one
device buffer [cells_per_rank, width], a single ``all_gather_into_tensor`` over
NCCL (NVLink/NVSwitch) assembles them on every GPU, and one device->host copy
brings them to the host, where the round-robin order is undone.  (gloo with CPU
tensors in the tests of the host logic: same code path.)
"""
import os

import numpy as np

# result arrays and the axis that indexes cells; packed in this order
_CELL_AXIS = {'c': 1, 'phi': 1, 'g': 1, 'flux': 0, 'status': 0, 'n_steps': 0, 'n_newton': 0, 'n_setups': 0}
_INT_FIELDS = ('status', 'n_steps', 'n_newton', 'n_setups')


def world():
    """(rank, world_size) of the initialised process group, else of the launcher's environment
    (torchrun sets RANK / WORLD_SIZE before init_process_group), else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    try:
        return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    except ValueError:
        return 0, 1


def group_ready():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized()
    except ImportError:
        return False


def shard_indices(n_cells, rank, world_size):
    return np.arange(rank, n_cells, world_size)


def result_shapes(n_out, nx_max, S):
    """per-cell shapes of the result arrays (cell axis removed), in packing order"""
    return {'c': (n_out, nx_max, S), 'phi': (n_out, nx_max), 'g': (n_out, nx_max), 'flux': (S,),
            'status': (), 'n_steps': (), 'n_newton': (), 'n_setups': ()}


def empty_results(n_out, nx_max, S):
    """result dict of a rank whose shard is empty (fewer cells than ranks)"""
    out = {}
    for name, shp in result_shapes(n_out, nx_max, S).items():
        axis = _CELL_AXIS[name]
        full = list(shp)
        full.insert(axis, 0)
        out[name] = np.zeros(full, dtype=np.int32 if name in _INT_FIELDS else np.float64)
    return out


def pack_results(local, per, n_out, nx_max, S, device=None):
    """result dict (torch tensors on any device, or numpy arrays) -> one float64 tensor [per, width],
    cell-major, rows beyond the shard's length zero.  int32 fields are exact in float64."""
    import torch
    cols = []
    n_loc = None
    for name, shp in result_shapes(n_out, nx_max, S).items():
        a = local[name]
        t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        if device is not None:
            t = t.to(device)
        t = t.movedim(_CELL_AXIS[name], 0).to(torch.float64)
        n_loc = t.shape[0]
        cols.append(t.reshape(n_loc, int(np.prod(shp)) if shp else 1))
    packed = torch.cat(cols, dim=1)
    if n_loc < per:
        pad = torch.zeros((per - n_loc, packed.shape[1]), dtype=torch.float64, device=packed.device)
        packed = torch.cat([packed, pad], dim=0)
    return packed.contiguous()


def unpack_results(gathered, n_cells, world_size, n_out, nx_max, S):
    """host array [world_size*per, width] -> ordered result dict (numpy), undoing cell j -> rank j % N"""
    per = gathered.shape[0] // world_size
    g = gathered.reshape(world_size, per, -1)
    order = np.zeros((n_cells, g.shape[2]))
    for r in range(world_size):
        idx = shard_indices(n_cells, r, world_size)
        order[idx] = g[r, :len(idx)]
    full = {}
    off = 0
    for name, shp in result_shapes(n_out, nx_max, S).items():
        w = int(np.prod(shp)) if shp else 1
        a = order[:, off:off + w].reshape((n_cells,) + shp)
        off += w
        if name in _INT_FIELDS:
            a = np.rint(a).astype(np.int32)
        full[name] = np.ascontiguousarray(np.moveaxis(a, 0, _CELL_AXIS[name]))
    return full


# Result arrays are numpy views of page-locked blocks (see _to_host_pinned).  A caller that keeps MANY results alive
# holds that much page-locked memory; CATINT_PINNED_RESULTS=0 hands out pageable copies instead (one more host copy).
PINNED_RESULTS = os.environ.get('CATINT_PINNED_RESULTS', '1') != '0'


def _as_result(h):
    a = h.numpy()
    return a if PINNED_RESULTS else np.array(a, copy=True)


def _to_host_pinned(t, key=None):
    """device tensor -> a FRESH pinned host tensor (asynchronous copy on the current stream; the caller synchronises
    and hands out its numpy view).  Pageable copies of the gathered results cost several times the kernel at 8
    GPUs, and copying a reused staging buffer into a new numpy array costs as much again (first-touch page faults:
    ~3 GB/s).  torch's caching host allocator recycles the pinned block once the caller has dropped the previous
    result, so a loop of solves pays the page-locking once.  CPU tensors pass through."""
    import torch
    if t.device.type != 'cuda':
        return t
    buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    return buf


def gather_results(local, n_cells, world_size, n_out, nx_max, S, device=None, root_only=False, to_host=True):
    """ONE all_gather of the packed per-rank results; the round-robin order (cell j -> rank j % N) is undone and
    the fields are split ON THE DEVICE (a transpose and slices), then one copy per field into pinned host memory.
    root_only: only rank 0 copies the gathered results to its host (the others return None) -- what a driver
    that saves on rank 0 needs; the default hands every rank the complete result like the reference's
    reduce_dict_mpi + sync_mpi (catint_io.py:154-178).  to_host=False: every rank keeps the ordered fields as
    tensors on the gather's device (the next wave of a continuation run reads its initial states from them; no host
    round trip)."""
    import torch
    import torch.distributed as dist
    per = (n_cells + world_size - 1) // world_size
    backend = dist.get_backend()
    dev = None
    if backend == 'nccl':
        dev = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
    packed = pack_results(local, per, n_out, nx_max, S, device=dev)
    out = torch.empty((world_size * per, packed.shape[1]), dtype=torch.float64, device=packed.device)
    dist.all_gather_into_tensor(out, packed)
    if root_only and to_host and dist.get_rank() != 0:
        return None
    width = out.shape[1]
    # row r*per + i holds cell i*N + r: [N, per, w] -> [per, N, w] -> cells in order
    ordered = out.reshape(world_size, per, width).transpose(0, 1).reshape(world_size * per, width)[:n_cells]
    full, staged = {}, {}
    off = 0
    for name, shp in result_shapes(n_out, nx_max, S).items():
        w = int(np.prod(shp)) if shp else 1
        a = ordered[:, off:off + w].reshape((n_cells,) + shp)
        off += w
        if name in _INT_FIELDS:
            a = torch.round(a).to(torch.int32)
        a = a.movedim(0, _CELL_AXIS[name]).contiguous()
        if not to_host:
            full[name] = a
            continue
        staged[name] = _to_host_pinned(a, name)
    if not to_host:
        full['gather_bytes'] = int(out.numel() * 8)
        return full
    if out.device.type == 'cuda':
        torch.cuda.current_stream(out.device).synchronize()
    for name, h in staged.items():
        full[name] = _as_result(h)                    # view of this call's own pinned block (kept alive by the array)
    full['gather_bytes'] = int(out.numel() * 8)
    return full


def solve_sharded(calc, batch, solve_fn=None, device=None, n_out=None, root_only=False, to_host=True):
    """Solve ``batch`` with the cells split over all ranks; every rank returns
    the complete, ordered result dict (numpy).  ``solve_fn(sub_batch) -> dict`` of
    torch tensors (device) or numpy arrays defaults to the CUDA path
    ``calc.solve_batch_device`` (tests of the host logic inject their own
    function; there is no CPU solver in the product).  A rank whose shard is
    empty (fewer cells than ranks) skips the solve and still joins the gather,
    like the reference's ``itask % size != rank: continue``.  ``to_host=False``: the ordered result stays on
    the device of every rank (dict of torch tensors), see ``results_to_host``."""
    rank, ws = world()
    if ws > 1 and not group_ready():
        raise RuntimeError('catint_b200: WORLD_SIZE=%d but torch.distributed is not initialised; call '
                           'torch.distributed.init_process_group first (one process per GPU)' % ws)
    if ws == 1:
        if not to_host:
            return (solve_fn or calc.solve_batch_device)(batch)
        if solve_fn is None:
            return calc.solve_batch(batch)
        return _to_host(solve_fn(batch))
    if solve_fn is None:
        solve_fn = calc.solve_batch_device
    if n_out is None:
        n_out = len(calc.output_times())
    idx = shard_indices(batch.B, rank, ws)
    extra = {}
    if len(idx) == 0:
        local = empty_results(n_out, batch.nx_max, batch.S)
    else:
        local = solve_fn(batch.select(idx))
        extra = {k: v for k, v in local.items() if k not in _CELL_AXIS}
    full = gather_results(local, batch.B, ws, n_out, batch.nx_max, batch.S, device=device, root_only=root_only,
                          to_host=to_host)
    if full is not None:
        full.update(extra)
    return full


def results_to_host(res, root_only=False):
    """ordered result dict of device tensors (solve_sharded(..., to_host=False)) -> numpy, through the pinned
    staging buffers of the gather; root_only: ranks other than 0 return None"""
    import torch
    if root_only and world()[0] != 0:
        return None
    staged, full = {}, {}
    dev = None
    for k, v in res.items():
        if torch.is_tensor(v):
            staged[k] = _to_host_pinned(v.contiguous(), k)
            if v.device.type == 'cuda':
                dev = v.device
        else:
            full[k] = v
    if dev is not None:
        torch.cuda.current_stream(dev).synchronize()
    for k, h in staged.items():
        full[k] = _as_result(h)
    return full


def _to_host(res):
    out = {}
    for k, v in res.items():
        out[k] = v.cpu().numpy() if hasattr(v, 'cpu') else v
    return out
