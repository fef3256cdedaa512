"""Calculator: the sweep / time-stepping driver (host side).

Drop-in for the reference's ``catint.calculator.Calculator``
(/root/reference/catint/calculator.py:53-241) restricted to the
finite-difference PNP path whose integrators live in
/root/reference/catint/calculator_old.py:210-1140 (``integrate_pnp``).

Same constructor keywords (calculator.py:55-56), same calc-string parsing
(:81-97), same time mesh and output indices (:105-138), same descriptor walk
order and labels (:204-219).  Difference by design: instead of solving one
descriptor point after the other, ``run()`` gathers the whole descriptor grid
into one cell batch, integrates all cells concurrently on the GPU(s) (one warp
per cell, cells sharded round-robin over the ranks of torch.distributed exactly
like ``itask % mpi_size == mpi_rank`` at calculator.py:209-212) and then fills
the result containers the reference's tools read (tp.cout, tp.potential,
tp.efield, tp.total_charge, tp.alldata[i][...]; conventions of
/root/reference/catint/comsol_reader.py:186-300).
"""
import os
import sys
import time

import numpy as np

from .units import unit_F
from . import backend as _be

# finite-difference integrators of the reference that map onto the BDF/Newton GPU integrator
FD_ODE_CALCS = ['odeint', 'lsoda', 'vode']
# the reference's fixed-step steppers (calculator_old.py:976-1029, :457-564), K4 of the backend
FD_STEP_CALCS = ['FTCS', 'Crank-Nicolson']


def poisson_bc_code(pb_bound):
    """pb_bound -> (backend code, wall potential, bulk potential, wall gradient, bulk gradient); the pairs the
    reference's get_potential_and_gradient accepts (calculator_old.py:703-705, :776-803)"""
    pw, pbk = pb_bound['potential']['wall'], pb_bound['potential']['bulk']
    gw, gb = pb_bound['gradient']['wall'], pb_bound['gradient']['bulk']
    if gw is not None and gb is not None:
        return None
    code = None
    if pw is not None and pbk is not None and gw is None and gb is None:
        code = _be.BC_DIRICHLET_BOTH
    elif pw is not None and pbk is None and gb is not None:
        code = _be.BC_DIRICHLET_WALL_NEUMANN_BULK
    elif pw is not None and pbk is None and gw is not None:
        code = _be.BC_DIRICHLET_WALL_NEUMANN_WALL
    elif pbk is not None and pw is None and gw is not None:
        code = _be.BC_DIRICHLET_BULK_NEUMANN_WALL
    elif pbk is not None and pw is None and gb is not None:
        code = _be.BC_DIRICHLET_BULK_NEUMANN_BULK
    if code is None:
        return None
    f = lambda x: 0.0 if x is None else float(x)
    return code, f(pw), f(pbk), f(gw), f(gb)


def build_cell_batch(tp, rate_mode='summed', points=None, poisson_bc='dirichlet', mesh=None):
    """Derive the per-cell parameter records for every point of the descriptor
    grid of ``tp`` (or for the explicit list ``points`` of override dicts).

    poisson_bc  'dirichlet' = the reference FD default (wall potential, bulk gradient,
                transport.py:207-210); 'stern' = Robin wall condition of the reference's COMSOL path
                (surface charge C_S*((phiM-phiPZC)-phi(0)), comsol_model.py:613,982) with phi(L)=0
    mesh        optional normalised node positions xi[n] in [0,1] (non-uniform mesh x = L*xi,
                same for all cells); default: the reference's uniform mesh

    Returns (CellBatch, models) where models[i] is the DerivedModel of cell i
    (light objects; used to label results)."""
    keys = list(tp.descriptors)
    if points is None:
        points = []
        for v1 in tp.descriptors[keys[0]]:
            for v2 in tp.descriptors[keys[1]]:
                points.append({keys[0]: float(v1), keys[1]: float(v2)})
    models = []
    cache = {}
    for ov in points:
        key = tuple(sorted(ov.items()))
        if key not in cache:
            cache[key] = tp.derive_for(**ov)
        models.append(cache[key])
    m0 = models[0]
    names = list(m0.species)
    S = len(names)
    for m in models:
        if list(m.species) != names:
            tp.logger.error('| CI | -- | species set changes along the descriptor grid; cannot batch')
            sys.exit()
        if m.flux_bound is None:
            tp.logger.error('| CI | -- | catmap fluxes are not available in the FD-PNP backend')
            sys.exit()
        if (m.flux_eq is None) != (m0.flux_eq is None) or (m.flux_eq is not None and (
                [p.code for p in m.flux_eq.programs] != [p.code for p in m0.flux_eq.programs]
                or [p.consts for p in m.flux_eq.programs] != [p.consts for p in m0.flux_eq.programs]
                or not np.array_equal(m.flux_eq.coef, m0.flux_eq.coef))):
            tp.logger.error('| CI | -- | flux equations change along the descriptor grid; cannot batch')
            sys.exit()
    # reaction table (species not transported -- H2O, e- -- drop out, calculator_old.py:169-171)
    reactions = []
    if m0.use_electrolyte_reactions and m0.electrolyte_reactions is not None:
        for r in m0.electrolyte_reactions:
            entry = m0.electrolyte_reactions[r]
            if 'rates' not in entry:
                continue
            ed = [names.index(s) for s in entry['reaction'][0] if s in names]
            pr = [names.index(s) for s in entry['reaction'][1] if s in names]
            reactions.append((ed, pr, entry['rates'][0], entry['rates'][1]))
    nu = _be.stoichiometry(S, reactions, rate_mode)
    B = len(models)
    par = np.zeros((B, _be.npar(S)))
    nx = np.zeros(B, dtype=np.int32)
    bc_code = None
    z = np.array([m0.species[s]['charge'] for s in names], dtype=np.int32)
    for c, m in enumerate(models):
        par[c, 0:S] = [m.species[s]['bulk_concentration'] for s in names]
        par[c, S:2 * S] = m.flux_bound[:, 0]          # inward flux = +species[sp]['flux']
        par[c, 2 * S:3 * S] = m.D
        par[c, 3 * S + 0] = m.beta
        par[c, 3 * S + 1] = m.eps
        if poisson_bc == 'stern':
            par[c, 3 * S + 2] = m.system['phiM'] - m.system['phiPZC']
            par[c, 3 * S + 3] = 0.0
        else:
            bc = poisson_bc_code(m.pb_bound)
            if bc is None:
                tp.logger.error('| CI | -- | pb_bound must give one potential and one gradient, or both potentials '
                                '(calculator_old.py:703-705, :776-803)')
                sys.exit()
            if bc_code is None:
                bc_code = bc[0]
            elif bc_code != bc[0]:
                tp.logger.error('| CI | -- | the kind of Poisson boundary changes along the descriptor grid; cannot batch')
                sys.exit()
            par[c, 3 * S + 2], par[c, 3 * S + 6], par[c, 3 * S + 7], par[c, 3 * S + 3] = bc[1:]
        par[c, 3 * S + 4] = m.system['Stern capacitance'] * 1e-2     # micro F/cm^2 -> F/m^2
        if mesh is None:
            par[c, 3 * S + 5] = m.dx
            nx[c] = m.nx
        else:
            par[c, 3 * S + 5] = m.xmax                                  # x = L*xi
            nx[c] = len(mesh)
    if poisson_bc not in ('dirichlet', 'stern'):
        tp.logger.error('| CI | -- | unknown poisson_bc "{}"'.format(poisson_bc))
        sys.exit()
    kw = {}
    if mesh is not None:
        kw = dict(mesh_id=np.zeros(B, dtype=np.int32), mesh_xi=np.asarray(mesh, dtype=np.float64)[None, :])
    if m0.flux_eq is not None:
        # wall kinetics as expressions of the surface state: programs shared, parameter values per cell
        kw['flux_eq'] = m0.flux_eq
        kw['fpar'] = np.array([m.fpar for m in models], dtype=float).reshape(B, len(m0.flux_eq.par_names))
    batch = _be.CellBatch(z, reactions, nu, par, nx, use_migration=m0.use_migration, species=names,
                          poisson_bc=_be.BC_STERN_ROBIN if poisson_bc == 'stern' else bc_code,
                          **kw)
    return batch, models


def continuation_plan(n_cells, k, row_length=None):
    """cold / warm split of a sweep for Calculator(continuation=k) and, for every warm cell, the position (in the
    cold list) of its nearest cold neighbour.  The cells of a 2D descriptor grid are numbered row by row (inner
    descriptor fastest, calculator.py:196-212): with ``row_length`` the plan is made per row -- every k-th cell of
    a row and its last cell are cold, and a warm cell only looks for neighbours in its own row, never across the
    row end, where the flattened neighbour is the other end of the inner descriptor's range."""
    L = int(row_length) if row_length else int(n_cells)
    if L <= 0 or n_cells % L != 0:
        L = int(n_cells)
    in_row = np.arange(0, L, k)
    if in_row[-1] != L - 1:
        in_row = np.append(in_row, L - 1)
    rows = n_cells // L
    cold = (np.arange(rows)[:, None] * L + in_row[None, :]).ravel()
    warm = np.setdiff1d(np.arange(n_cells), cold)
    if len(warm) == 0:
        return cold, warm, np.zeros(0, dtype=np.int64)
    wr, wi = warm // L, warm % L
    j = np.searchsorted(in_row, wi)                                  # in_row[j-1] < wi < in_row[j]
    j = np.where(wi - in_row[j - 1] <= in_row[j] - wi, j - 1, j)     # ties: the lower neighbour
    return cold, warm, wr * len(in_row) + j


def continuation_brackets(cold, warm, row_length=None):
    """for every warm cell the positions (in the cold list) of the two cold cells of its row that bracket it and
    its relative position between them, (warm - left) / (right - left): the first and the last cell of a row are
    cold, so every warm cell has both.  Used for the interpolated starting state of Calculator.run_continuation."""
    cold, warm = np.asarray(cold), np.asarray(warm)
    right = np.searchsorted(cold, warm)                               # cold[right-1] < warm < cold[right]
    left = right - 1
    w = (warm - cold[left]) / (cold[right] - cold[left]).astype(float)
    if row_length:
        assert np.all(cold[left] // row_length == warm // row_length) and np.all(cold[right] // row_length == warm // row_length)
    return left, right, w


COLD_WAVE_CELLS = 1024        # cells of one resident wave per GPU (1184 warp slots on a B200): a cold wave of up to this
#                               size lasts as long as its slowest cell, whatever the number of cells


def auto_continuation_k(n_cells, world_size=1, row_length=None):
    """continuation='auto': the stride k that makes the cold wave ONE resident wave per GPU, or None where a plain
    run is the better choice -- sweeps of fewer than 4 resident waves (measured: 1024 cells 7.1 k cells/s with
    continuation against 7.4 k plain) and 2D descriptor grids (measured on C4: neighbours along the boundary-layer
    axis are poor starting states)."""
    if row_length is not None or n_cells < 4 * COLD_WAVE_CELLS * max(1, int(world_size)):
        return None
    return max(2, int(n_cells) // (COLD_WAVE_CELLS * max(1, int(world_size))))


def shard_indices(n_cells, rank, world_size):
    """round-robin cell -> rank map (calculator.py:209-212: itask % size == rank)"""
    return np.arange(rank, n_cells, world_size)


class Calculator():

    def __init__(self, transport=None, dt=None, tmax=None, ntout=1, calc=None,
                 scale_pb_grid=None, tau_jacobi=1e-7, tau_scf=5e-5, mix_scf=0.5, mode='time-dependent',
                 rtol=1.49012e-8, atol=1.49012e-8, rate_mode='summed', device=None, max_steps=100000,
                 poisson_bc='dirichlet', mesh=None, continuation=None):
        """Reference keywords (calculator.py:55-56) plus backend options:
        rtol/atol  error tolerances of the BDF integrator (scipy odeint defaults)
        rate_mode  'summed' (default) or 'legacy_overwrite' 
        poisson_bc 'dirichlet' (reference FD default) or 'stern' (Robin wall of the COMSOL path)
        mesh       optional normalised non-uniform node positions (see build_cell_batch)
        mode       'time-dependent': state at the output times of the time mesh
                   'stationary': integrate to tmax, then Newton-polish the steady residual
        continuation  None (every cell starts from the bulk state, the reference's 'internal-reinit',
                   transport.py:834-842) or an integer k >= 2 ('internal-cont' for a batch): every k-th cell of the
                   sweep is solved cold, the cells in between start from the converged state of the nearest cold
                   cell and are integrated over the same time span.  Stationary mode only; the warm cells reach
                   the same steady state with a fraction of the steps (see Calculator.run_continuation).
                   'auto': k = cells / (1024 x GPUs) for a one-descriptor sweep of at least 4096 cells per GPU,
                   a plain run otherwise (auto_continuation_k).
        """
        self.mode = mode
        self.tau_scf = tau_scf
        self.mix_scf = mix_scf
        if transport is None:
            print('No transport object provided for calculator. Stopping here.')
            sys.exit()
        self.tp = transport
        if calc is None:
            calc = self.tp.calc
        if calc is None:
            self.tp.logger.error('No calculator found with this name. Aborting.')
            sys.exit()
        self.tp.ntout = ntout

        self.use_lax_friedrich = False
        parts = calc.split('--')
        self.calc_method = None
        if len(parts) > 1:
            if parts[-1] == 'LF':
                self.use_lax_friedrich = True
            else:
                self.calc_method = parts[-1]
            self.calc = parts[0]
        else:
            self.calc = calc
        self.calc_list = ['FTCS', 'Crank-Nicolson', 'odeint', 'vode', 'lsoda', 'dopri5', 'dop853', 'odeint',
                          'odespy', 'comsol']
        if self.calc not in self.calc_list:
            self.tp.logger.error('No calculator found with this name. Aborting.')
            sys.exit()
        if self.calc not in FD_ODE_CALCS + FD_STEP_CALCS:
            self.tp.logger.error('| CI | -- | calculator "{}" is outside the scope of the B200 FD-PNP backend '
                                 '(available: {})'.format(self.calc, FD_ODE_CALCS + FD_STEP_CALCS))
            sys.exit()
        if self.use_lax_friedrich and self.calc not in FD_STEP_CALCS:
            self.tp.logger.error('| CI | -- | Lax-Friedrichs terms belong to the fixed-step steppers (FTCS--LF, '
                                 'Crank-Nicolson--LF)')
            sys.exit()
        if scale_pb_grid is not None:
            # calculator_old.py:684-712, :805-813: 'linear' divides dx by the option string (TypeError upstream), 'log'
            # integrates a logarithmic grid with the uniform dx -- neither is a result to reproduce; never ignore it
            self.tp.logger.error('| CI | -- | scale_pb_grid is not available (the reference option cannot run / integrates '
                                 'the log grid with the uniform dx); pass mesh= for a graded grid')
            sys.exit()
        self.scale_pb_grid = scale_pb_grid
        self.tau_jacobi = tau_jacobi
        self.rtol, self.atol = rtol, atol
        self.rate_mode = rate_mode
        self.poisson_bc = poisson_bc
        self.mesh = mesh
        self.device = device
        self.max_steps = max_steps
        self.continuation = continuation
        if continuation is not None and (mode != 'stationary' or
                                         (continuation != 'auto' and (isinstance(continuation, str) or int(continuation) < 2))):
            self.tp.logger.error('| CI | -- | continuation needs an integer >= 2 or "auto", and mode="stationary"')
            sys.exit()

        # time mesh and output indices (calculator.py:105-138)
        if dt is not None:
            self.tp.dt = dt
        if tmax is not None:
            self.tp.tmax = tmax
        if tmax is not None or dt is not None:
            self.tp.tmesh = np.arange(0, self.tp.tmax + self.tp.dt, self.tp.dt)
        else:
            self.tp.logger.warning('No time mesh given, defaulting to range(0,1,0.1)')
            self.tp.dt, self.tp.tmax = 0.1, 0.9
            self.tp.tmesh = np.arange(0, 1, 0.1)
        self.tp.nt = len(self.tp.tmesh)
        self.tp.tmesh_init = self.tp.tmesh
        self.tp.nt_init = self.tp.nt
        self.tp.dt_init = self.tp.dt
        self.tp.tmax_init = self.tp.tmax
        self.oldtime = np.inf
        self.tp.itout = []
        stride = int(self.tp.nt / float(self.tp.ntout))
        for it, t in enumerate(self.tp.tmesh):
            if it == self.tp.nt - 1:
                self.tp.itout.append(it)
            elif it > 1 and stride > 0 and it % stride == 0:
                self.tp.itout.append(it)
        self.tp.ntout = len(self.tp.itout)
        self.stats = {}

    # ------------------------------------------------------------------
    def get_rates(self, C):
        """homogeneous rates [S,nx] of the concentrations C[S,nx] (host helper with the
        semantics of self.rate_mode; the device evaluates the same table)."""
        tp = self.tp
        names = list(tp.species)
        rates = np.zeros([tp.nspecies, C.shape[1]])
        if not tp.use_electrolyte_reactions or tp.electrolyte_reactions is None:
            return rates
        rx = []
        for r in tp.electrolyte_reactions:
            e = tp.electrolyte_reactions[r]
            if 'rates' in e:
                rx.append(([names.index(s) for s in e['reaction'][0] if s in names],
                           [names.index(s) for s in e['reaction'][1] if s in names], e['rates'][0], e['rates'][1]))
        nu = _be.stoichiometry(tp.nspecies, rx, self.rate_mode)
        for r, (ed, pr, kf, kr) in enumerate(rx):
            net = kf * np.prod(C[ed, :], axis=0) - kr * np.prod(C[pr, :], axis=0)
            rates += nu[:, r:r + 1] * net[None, :]
        return rates

    # ------------------------------------------------------------------
    def output_times(self):
        ts = [float(self.tp.tmesh[i]) for i in self.tp.itout if self.tp.tmesh[i] > 0.0]
        if not ts:
            ts = [float(self.tp.tmax)]
        return ts

    def solve_batch_device(self, batch, backend=None, pinned=None, y0=None, max_steps=None):
        """host CellBatch -> result dict of DEVICE tensors (plus the host->device byte count): upload of the
        cell parameters and the solve.  The sharded driver gathers these on the device."""
        import torch
        if backend is None:
            backend = self._backend()
        db = backend.upload(batch, pinned=pinned)
        if self.calc in FD_STEP_CALCS:
            return self._step_batch_device(backend, db, y0)
        if batch.poisson_bc not in (_be.BC_DIRICHLET_WALL_NEUMANN_BULK, _be.BC_STERN_ROBIN,
                                    _be.BC_DIRICHLET_BULK_NEUMANN_BULK):
            self.tp.logger.error('| CI | -- | the implicit integrator takes the pb_bound pairs with a bulk gradient (wall '
                                 'or bulk potential) or the Stern boundary; the other pairs run with calc=FTCS / '
                                 'Crank-Nicolson')
            sys.exit()
        mode = _be.MODE_STEADY if self.mode == 'stationary' else _be.MODE_TRANSIENT
        y0_dev = None
        if torch.is_tensor(y0):
            y0_dev = y0.to(device=backend.device, dtype=torch.float64).contiguous()
        elif y0 is not None:
            y0_dev = torch.as_tensor(np.ascontiguousarray(y0, dtype=np.float64)).to(backend.device)
        out = backend.solve(db, self.output_times(), mode=mode, rtol=self.rtol, atol=self.atol,
                            max_steps=self.max_steps if max_steps is None else int(max_steps), y0=y0_dev)
        out = dict(out)
        out['h2d_bytes'] = db.h2d_bytes + (0 if y0 is None else int(y0_dev.numel() * 8))
        return out

    def _step_batch_device(self, backend, db, y0):
        """calc = FTCS / Crank-Nicolson (calculator_old.py:1121-1140): the fixed time mesh of the reference, outputs
        after the steps in tp.itout (FTCS counts from 0, Crank-Nicolson from 1, as the reference's loops do)"""
        import torch
        tp = self.tp
        b = db.batch
        stepper = _be.STEPPER_FTCS if self.calc == 'FTCS' else _be.STEPPER_CRANK_NICOLSON
        first = 0 if stepper == _be.STEPPER_FTCS else 1
        itout = [i for i in tp.itout if first <= i < tp.nt]
        if not itout:
            itout = [tp.nt - 1]
        c0 = None
        if y0 is not None:
            c0 = torch.as_tensor(np.ascontiguousarray(y0, dtype=np.float64)).to(backend.device)
        out = backend.step(db, stepper, tp.dt, tp.nt, itout, c0=c0, lax_friedrich=self.use_lax_friedrich)
        B = b.B
        dev = backend.device
        finite = torch.isfinite(out['c'][-1]).reshape(B, -1).all(dim=1)
        out = dict(out)
        out['flux'] = torch.zeros((B, b.S), dtype=torch.float64, device=dev)
        out['status'] = torch.where(finite, 0, 4).to(torch.int32)        # CATINT_PNP_CELL_NOT_FINITE
        out['n_steps'] = torch.full((B,), tp.nt - first, dtype=torch.int32, device=dev)
        out['n_newton'] = torch.zeros((B,), dtype=torch.int32, device=dev)
        out['n_setups'] = torch.zeros((B,), dtype=torch.int32, device=dev)
        out['h2d_bytes'] = db.h2d_bytes
        return out

    def solve_batch(self, batch, backend=None, pinned=None, y0=None):
        """host CellBatch -> host result dict (numpy).  Host->device copies, the
        solve and the device->host copies all happen here.  y0: optional initial
        concentrations [B,nx_max,S] (default: bulk everywhere, the reference's c0)."""
        out = self.solve_batch_device(batch, backend=backend, pinned=pinned, y0=y0)
        host = {k: (v.cpu().numpy() if hasattr(v, 'cpu') else v) for k, v in out.items()}
        host['d2h_bytes'] = sum(int(v.numel() * v.element_size()) for v in out.values() if hasattr(v, 'numel'))
        return host

    def _backend(self):
        if getattr(self, '_bk', None) is None:
            self._bk = _be.PnpBackend(self.device)
        return self._bk

    def initial_state_from_folder(self, batch):
        """system['init_folder'] (reference: calculator.py:303-309, a previous results folder) -> initial
        concentrations of every cell from that run's alldata (warm start / continuation of a sweep).
        The restart folder must use the same descriptor grid and mesh (reference: calculator.py:242-251)."""
        tp = self.tp
        folder = tp.system.get('init_folder', None)
        if folder is None:
            return None
        from .catint_io import load_obj
        alldata = load_obj('alldata', folder)
        names = list(tp.species)
        if len(alldata) != batch.B:
            tp.logger.error('| CI | -- | init_folder {} holds {} cells, the current descriptor grid {}'.format(
                folder, len(alldata), batch.B))
            sys.exit()
        y0 = np.zeros((batch.B, batch.nx_max, batch.S))
        for c in range(batch.B):
            n = int(batch.nx[c])
            for k, sp in enumerate(names):
                conc = np.asarray(alldata[c]['species'][sp]['concentration'], dtype=float)
                if len(conc) != n:
                    tp.logger.error('| CI | -- | init_folder cell {} has {} nodes, expected {}'.format(c, len(conc), n))
                    sys.exit()
                y0[c, :n, k] = conc
        tp.logger.info('| CI | -- | Initialised {} cells from {}'.format(batch.B, folder))
        return y0

    def run(self):
        tp = self.tp
        t0 = time.time()
        keys = list(tp.descriptors)
        batch, models = build_cell_batch(tp, rate_mode=self.rate_mode, poisson_bc=self.poisson_bc, mesh=self.mesh)
        for i1, v1 in enumerate(tp.descriptors[keys[0]]):
            for i2, v2 in enumerate(tp.descriptors[keys[1]]):
                tp.logger.debug('| CI | -- | cell {} : {} = {} and {} = {}'.format(
                    str(i1 + 1).zfill(4) + '_' + str(i2 + 1).zfill(4), keys[0], v1, keys[1], v2))
        tp.logger.info('| CI | -- | Starting batched calculation of {} cells ({} x {})'.format(
            batch.B, len(tp.descriptors[keys[0]]), len(tp.descriptors[keys[1]])))
        from . import distributed as _dist
        y0 = self.initial_state_from_folder(batch)
        k_cont = None
        if y0 is None and self.continuation is not None:
            n1, n2 = len(tp.descriptors[keys[0]]), len(tp.descriptors[keys[1]])
            row_length = n2 if (n1 > 1 and n2 > 1) else None
            k_cont = (auto_continuation_k(batch.B, _dist.world()[1], row_length) if self.continuation == 'auto'
                      else int(self.continuation))
        if k_cont is not None:
            res = self.run_continuation(batch, row_length=row_length, k=k_cont)
        elif y0 is None:
            res = _dist.solve_sharded(self, batch)
        else:
            res = _dist.solve_sharded(self, batch,
                                      solve_fn=lambda sub: self.solve_batch_device(sub, y0=y0[sub.origin]))
        t1 = time.time()
        if res is not None:
            self.scatter_results(batch, models, res)
            nconv = int(np.sum(res['status'] == 0))
            self.stats = {'cells': batch.B, 'converged': nconv, 'seconds': t1 - t0,
                          'n_steps': res['n_steps'], 'n_newton': res['n_newton'], 'status': res['status']}
            tp.logger.info('| CI | -- | {} of {} cells converged in {:.3f} s'.format(nconv, batch.B, t1 - t0))
            for c in np.nonzero(res['status'] != 0)[0]:
                tp.logger.warning('| CI | -- | cell {} ({}) did not converge: {}'.format(
                    c, tp.alldata_names[c], _be.CELL_STATUS.get(int(res['status'][c]), res['status'][c])))
            tp.save()
        return res

    # ------------------------------------------------------------------
    def run_continuation(self, batch, root_only=False, row_length=None, k=None):
        """two waves over a sweep (the batch analogue of the reference's COMSOL option 'internal-cont',
        transport.py:834-842: "the solution of the previous parameter set is used to initialize the next"):
        wave 1 solves every k-th cell from the bulk state, wave 2 the others from a state built out of the
        converged wave-1 states, over the full time span (so a poor starting guess costs steps, not accuracy).
        Starting state of a warm cell: the linear interpolation, along the sweep, of the two cold cells that
        bracket it; if one of them failed, the state of the nearest one; else the cell's own bulk state.  Only
        cold cells with the warm cell's bulk composition and node count are used (from another composition the
        integration may end on another root of the discrete system).  A warm cell may take at most
        3x the steps of the slowest converged cold cell; warm cells that fail (or hit that cap) are solved again
        from the bulk state (wave 3), so a cell that converges in a plain run converges here.
        All waves are sharded like a plain run.  The results of wave 1 stay on the devices: every rank builds the
        initial states of its wave-2 cells there (index_select of the gathered wave-1 states), the waves are
        merged there, and ONE device->host copy brings the whole batch to the host (root_only: of rank 0 only;
        the other ranks return None).  row_length: cells per row of a 2D descriptor grid (the plan is made per
        row); k overrides self.continuation.  Returns the result dict of the whole batch."""
        import torch
        from . import distributed as _dist
        k = int(self.continuation if k is None else k)
        B, S = batch.B, batch.S
        cold, warm, nearest = continuation_plan(B, k, row_length)
        L_row = int(row_length) if (row_length and B % int(row_length) == 0) else None

        def as_tensors(r):
            return {key: (torch.as_tensor(v) if isinstance(v, np.ndarray) else v) for key, v in r.items()}

        timing = os.environ.get('CATINT_CONT_TIMING')         # debug: synchronised wall time of the four stages
        marks = []

        def mark(name):
            if timing:
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))

        mark('start')
        r1 = as_tensors(_dist.solve_sharded(self, batch.select(cold), to_host=False))
        mark('wave1')
        dev = r1['c'].device
        if len(warm) == 0:
            return _dist.results_to_host(r1, root_only=root_only)
        near_t = torch.as_tensor(nearest, device=dev)    # position (in wave 1) of every warm cell's nearest cold neighbour
        c1 = r1['c'][-1]                                            # [n_cold, nx_max, S]
        ok = (r1['status'] == 0).index_select(0, near_t)            # failed neighbour: cold start from the bulk state
        # ... and so does a cell whose neighbour has another bulk composition (a sweep over bulk_pH or a bulk
        # concentration): from such a state the integration may end on another root of the discrete system than the
        # cold run selects (CPU prototype on C4, profiles/r2/continuation_axis_probe.txt: 4 % off, or a blow-up).
        cb_w, cb_n = batch.par[warm, 0:S], batch.par[cold[nearest], 0:S]
        same_bulk = np.all(np.abs(cb_w - cb_n) <= 1e-12 * np.abs(cb_n), axis=1)
        same_bulk &= batch.nx[warm] == batch.nx[cold[nearest]]     # ragged grids: a state only fits a cell of its node count
        ok = ok & torch.as_tensor(same_bulk, device=dev)
        bulk = torch.as_tensor(np.ascontiguousarray(batch.par[warm, 0:S])).to(dev)

        conv1 = r1['status'] == 0
        cold_max = int(torch.where(conv1, r1['n_steps'], torch.zeros_like(r1['n_steps'])).max()) if len(cold) else 0
        warm_cap = int(min(self.max_steps, max(1000, 3 * cold_max)))
        if getattr(self, 'continuation_warm_cap', None):              # test / tuning knob: explicit cap of the warm wave
            warm_cap = int(self.continuation_warm_cap)

        # starting state: linear interpolation (in the cell index = along the swept descriptor) between the converged
        # states of the two cold cells that bracket the warm cell -- second-order close to its own steady state where
        # the nearest neighbour's state is first-order close (CPU prototype on the 16384-cell C2 sweep: 244 -> 103,
        # 138 -> 52, 70 -> 20 steps; same end states) -- when both have converged and share the cell's bulk
        # composition; else the nearest neighbour's state under the same conditions; else the bulk state.
        interp = getattr(self, 'continuation_interpolate', True)
        if interp:
            left, right, wgt = continuation_brackets(cold, warm, L_row)
            left_t, right_t = torch.as_tensor(left, device=dev), torch.as_tensor(right, device=dev)
            w_t = torch.as_tensor(wgt, device=dev, dtype=c1.dtype)
            same2 = np.ones(len(warm), dtype=bool)
            for side in (left, right):
                cb_s = batch.par[cold[side], 0:S]
                same2 &= np.all(np.abs(cb_w - cb_s) <= 1e-12 * np.abs(cb_s), axis=1)
                same2 &= batch.nx[warm] == batch.nx[cold[side]]
            ok2 = conv1.index_select(0, left_t) & conv1.index_select(0, right_t) & torch.as_tensor(same2, device=dev)
        else:
            ok2 = torch.zeros_like(ok)
        ok = ok | ok2                                               # warm-started one way or the other

        def warm_start(sb):
            o = torch.as_tensor(np.asarray(sb.origin), device=dev)
            src = c1.index_select(0, near_t.index_select(0, o))
            if interp:
                a = c1.index_select(0, left_t.index_select(0, o))
                b = c1.index_select(0, right_t.index_select(0, o))
                mix = a + (b - a) * w_t.index_select(0, o)[:, None, None]
                src = torch.where(ok2.index_select(0, o)[:, None, None], mix, src)
            y0 = torch.where(ok.index_select(0, o)[:, None, None], src, bulk.index_select(0, o)[:, None, :])
            return self.solve_batch_device(sb, y0=y0.contiguous(), max_steps=warm_cap)

        sub = batch.select(warm)
        sub.origin = np.arange(sub.B)                  # shards of `sub` index the warm list by their position in `sub`
        r2 = as_tensors(_dist.solve_sharded(self, sub, solve_fn=warm_start, to_host=False))
        mark('wave2')
        # wave 3: warm-started cells that did not converge start again from the bulk state (every rank holds the same
        # gathered wave-2 status, so every rank takes the same decision)
        rerun = torch.nonzero((r2['status'] != 0) & ok.to(r2['status'].device)).flatten()
        n_rerun = int(rerun.numel())
        if n_rerun:
            r3 = as_tensors(_dist.solve_sharded(self, batch.select(warm[rerun.cpu().numpy()]), to_host=False))
            for key, v in r2.items():
                if torch.is_tensor(v) and v.ndim > 0 and key in r3:
                    v.index_copy_(1 if key in ('c', 'phi', 'g') else 0, rerun.to(v.device), r3[key].to(v.device))
        mark('wave3')
        cold_t, warm_t = torch.as_tensor(cold, device=dev), torch.as_tensor(warm, device=dev)
        res = {}
        for key, v in r1.items():
            if not torch.is_tensor(v) or v.ndim == 0:
                res[key] = v + r2[key] if key in ('h2d_bytes', 'gather_bytes') and key in r2 else v
                continue
            ax = 1 if key in ('c', 'phi', 'g') else 0
            shape = list(v.shape)
            shape[ax] = B
            full = torch.empty(shape, dtype=v.dtype, device=dev)
            full.index_copy_(ax, cold_t, v)
            full.index_copy_(ax, warm_t, r2[key].to(dev))
            res[key] = full
        self.continuation_stats = {'cold_cells': len(cold), 'warm_cells': len(warm),
                                   'cold_steps_mean': float(r1['n_steps'].double().mean()),
                                   'warm_steps_mean': float(r2['n_steps'].double().mean()),
                                   'warm_newton_mean': float(r2['n_newton'].double().mean()),
                                   'warm_setups_mean': float(r2['n_setups'].double().mean()),
                                   'warm_step_cap': warm_cap, 'rerun_cold_cells': n_rerun,
                                   'warm_started_cells': int(ok.sum()), 'interpolated_starts': int(ok2.sum())}
        mark('merge')
        host = _dist.results_to_host(res, root_only=root_only)
        mark('to_host')
        if timing:
            self.continuation_timing = {b[0]: b[1] - a[1] for a, b in zip(marks[:-1], marks[1:])}
        return host

    def scatter_results(self, batch, models, res):
        """fill tp.cout/potential/efield/total_charge (last cell, like the serial
        reference loop would leave them) and tp.alldata[i] for every cell.  Per-cell
        quantities are computed for the whole batch at once (numpy); the per-cell loop only
        hands out views (64k-cell sweeps: seconds, not minutes)."""
        tp = self.tp
        names = list(tp.species)
        S = len(names)
        c_all = res['c']          # [n_out,B,nx_max,S]
        cfin_all = c_all[-1]      # [B,nx_max,S]
        g_all, phi_all = res['g'][-1], res['phi'][-1]
        q = np.array([tp.species[s]['charge'] for s in names]) * unit_F
        rho_all = cfin_all @ q    # [B,nx_max]
        with np.errstate(invalid='ignore', divide='ignore'):
            if 'H+' in names:
                ph_all = -np.log10(cfin_all[:, :, names.index('H+')] / 1000.)
            elif 'OH-' in names:
                ph_all = 14 + np.log10(cfin_all[:, :, names.index('OH-')] / 1000.)
            else:
                ph_all = None
        # current density = flux*nel*F/nprod/10 mA/cm^2 (comsol_reader.py:243-246); one factor per species
        m0 = models[0]
        cd_factor = {}
        if m0.electrode_reactions is not None:
            for sp in names:
                if sp in m0.electrode_reactions:
                    er = m0.electrode_reactions[sp]
                    nprod = len([a for a in er['reaction'][1] if a == sp])
                    cd_factor[sp] = er['nel'] * unit_F / nprod / 10.
        flux = res['flux']
        status = res['status']
        nxs = batch.nx
        for c in range(batch.B):
            n = int(nxs[c])
            ad = tp.alldata[c]
            spd = ad['species']
            cf = cfin_all[c, :n]
            for k, sp in enumerate(names):
                d = spd.setdefault(sp, {})
                d['concentration'] = cf[:, k]
                d['surface_concentration'] = float(cf[0, k])
                fl = float(flux[c, k])
                d['electrode_flux'] = fl
                if sp in cd_factor:
                    d['electrode_current_density'] = fl * cd_factor[sp]
            sysd = ad['system']
            sysd['potential'] = phi_all[c, :n]
            sysd['efield'] = -g_all[c, :n]
            sysd['charge_density'] = rho_all[c, :n]
            sysd['surface_potential'] = float(phi_all[c, 0])
            if ph_all is not None:
                sysd['pH'] = ph_all[c, :n]
                if np.isfinite(ph_all[c, 0]):
                    sysd['surface_pH'] = float(ph_all[c, 0])
                else:
                    tp.logger.warning('| CI | -- | negative surface concentration, surface pH cannot be evaluated '
                                      'for cell {}'.format(c))
            sysd['status'] = _be.CELL_STATUS.get(int(status[c]), int(status[c]))
        # containers of the serial FD path (calculator_old.py:816-818, 966-973): last cell
        last = batch.B - 1
        n = int(batch.nx[last])
        tp.cout = [np.ascontiguousarray(c_all[k, last, :n, :].T).reshape(-1) for k in range(c_all.shape[0])]
        tp.efield = -res['g'][-1, last, :n]
        tp.potential = res['phi'][-1, last, :n]
        tp.total_charge = c_all[-1, last, :n, :] @ q
        for k, sp in enumerate(names):
            tp.species[sp]['surface_concentration'] = float(c_all[-1, last, 0, k])
        last_sys = tp.alldata[last]['system']
        if 'surface_pH' in last_sys:
            tp.system['surface_pH'] = last_sys['surface_pH']
        tp.system['surface_potential'] = last_sys['surface_potential']
        tp.system['potential'] = np.array(last_sys['potential'])
        tp.system['efield'] = np.array(last_sys['efield'])
        tp.system['charge_density'] = np.array(last_sys['charge_density'])
