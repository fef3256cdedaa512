"""Sharding of a cell batch over the ranks of torch.distributed.

The reference's only parallel strategy is data parallelism over the descriptor
grid with a static round-robin map ``itask % mpi_size == mpi_rank``
(/root/reference/catint/calculator.py:209-212) and a gather of the result
dictionaries (/root/reference/catint/catint_io.py:154-178; disabled upstream).
Here: one process per GPU, cell j -> rank j % world_size, no traffic while the
cells integrate, ONE final all_gather of the result arrays (NCCL over
NVLink/NVSwitch on a B200 box; gloo in the CPU tests of the host logic).
"""
import numpy as np

# result arrays and the axis that indexes cells
_CELL_AXIS = {'c': 1, 'phi': 1, 'g': 1, 'flux': 0, 'status': 0, 'n_steps': 0, 'n_newton': 0, 'n_setups': 0}


def world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_indices(n_cells, rank, world_size):
    return np.arange(rank, n_cells, world_size)


def gather_results(local, n_cells, rank, world_size, device=None):
    """all_gather the per-rank result dicts (numpy) and restore the original
    cell order.  Shards are padded to the same length for the collective."""
    import torch
    import torch.distributed as dist
    per = (n_cells + world_size - 1) // world_size
    backend = dist.get_backend()
    dev = torch.device(device if (backend == 'nccl' and device is not None) else
                       ('cuda' if backend == 'nccl' else 'cpu'))
    full = {}
    for name, axis in _CELL_AXIS.items():
        a = np.asarray(local[name])
        a = np.moveaxis(a, axis, 0)
        pad = np.zeros((per,) + a.shape[1:], dtype=a.dtype)
        pad[:a.shape[0]] = a
        t = torch.from_numpy(np.ascontiguousarray(pad)).to(dev)
        out = torch.empty((world_size * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t)
        out = out.cpu().numpy().reshape((world_size, per) + tuple(t.shape[1:]))
        merged = np.zeros((n_cells,) + a.shape[1:], dtype=a.dtype)
        for r in range(world_size):
            idx = shard_indices(n_cells, r, world_size)
            merged[idx] = out[r, :len(idx)]
        full[name] = np.moveaxis(merged, 0, axis)
    return full


def solve_sharded(calc, batch, solve_fn=None):
    """Solve ``batch`` with the cells split over all ranks; every rank returns
    the complete, ordered result dict.  ``solve_fn(sub_batch) -> dict`` defaults
    to the CUDA path ``calc.solve_batch`` (tests of the host logic inject their
    own function; there is no CPU solver in the product)."""
    rank, ws = world()
    if solve_fn is None:
        solve_fn = calc.solve_batch
    if ws == 1:
        return solve_fn(batch)
    idx = shard_indices(batch.B, rank, ws)
    sub = batch.select(idx)
    local = solve_fn(sub)
    dev = None
    try:
        import torch
        if torch.cuda.is_available():
            dev = 'cuda:%d' % torch.cuda.current_device()
    except ImportError:
        pass
    full = gather_results(local, batch.B, rank, ws, device=dev)
    for k in local:
        if k not in full:
            full[k] = local[k]
    return full
