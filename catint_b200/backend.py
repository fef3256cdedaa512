"""ctypes binding of the C ABI in include/catint_pnp.h and the cell-batch container.

PyTorch is only used for device memory, streams and (elsewhere)
torch.distributed; the solver itself is the hand-written CUDA library
``libcatint_pnp.so`` built from catint_b200/csrc.  There is NO CPU fallback:
every compute entry point raises if the library or a B200 is missing.
"""
import ctypes
import os

import numpy as np

MAX_SPECIES = 14
MAX_REACTIONS = 12
MAX_REACTANTS = 4

BC_DIRICHLET_WALL_NEUMANN_BULK = 0
BC_STERN_ROBIN = 1
BC_DIRICHLET_BOTH = 2                 # the reference's other pb_bound pairs (calculator_old.py:776-803):
BC_DIRICHLET_WALL_NEUMANN_WALL = 3    # available in potential() and in the fixed-step steppers (step())
BC_DIRICHLET_BULK_NEUMANN_WALL = 4
BC_DIRICHLET_BULK_NEUMANN_BULK = 5
STEPPER_FTCS = 0
STEPPER_CRANK_NICOLSON = 1
MODE_TRANSIENT = 0
MODE_STEADY = 1
MODE_KEEP_ALL = 0x100   # flag: no elimination of passive species (see include/catint_pnp.h)

CELL_STATUS = {0: 'converged', 1: 'max_steps', 2: 'corrector_failed', 3: 'error_test_failed',
               4: 'not_finite', 5: 'polish_failed', 6: 'step_underflow', 7: 'bad_input'}
MAX_BLOCK = 13          # unknowns per node handled by catint_pnp_solve_batch / _jacobian_batch
MIN_NODES = 6           # CATINT_PNP_MIN_NODES
MAX_OUTPUT_TIMES = 4096

_LIB_NAME = 'libcatint_pnp.so'
_lib = None


MAX_FLUX_EQ, MAX_FLUX_CODE, MAX_FLUX_CONST, MAX_FLUX_PAR = 4, 96, 32, 24


class CatintPnpFluxEq(ctypes.Structure):
    _fields_ = [
        ('n_eq', ctypes.c_int32), ('n_par', ctypes.c_int32),
        ('n_code', ctypes.c_int32 * MAX_FLUX_EQ),
        ('code', (ctypes.c_int32 * MAX_FLUX_CODE) * MAX_FLUX_EQ),
        ('consts', (ctypes.c_double * MAX_FLUX_CONST) * MAX_FLUX_EQ),
        ('coef', (ctypes.c_double * MAX_FLUX_EQ) * MAX_SPECIES),
    ]


class CatintPnpShared(ctypes.Structure):
    _fields_ = [
        ('S', ctypes.c_int32), ('nx_max', ctypes.c_int32), ('R', ctypes.c_int32),
        ('poisson_bc', ctypes.c_int32), ('use_migration', ctypes.c_int32), ('n_mesh', ctypes.c_int32),
        ('z', ctypes.c_int32 * MAX_SPECIES),
        ('educt', (ctypes.c_int32 * MAX_REACTANTS) * MAX_REACTIONS),
        ('product', (ctypes.c_int32 * MAX_REACTANTS) * MAX_REACTIONS),
        ('kf', ctypes.c_double * MAX_REACTIONS),
        ('kr', ctypes.c_double * MAX_REACTIONS),
        ('nu', (ctypes.c_double * MAX_REACTIONS) * MAX_SPECIES),
        ('flux_eq', ctypes.POINTER(CatintPnpFluxEq)),
    ]


class CatintPnpCells(ctypes.Structure):
    _fields_ = [('par', ctypes.c_void_p), ('nx', ctypes.c_void_p), ('mesh_id', ctypes.c_void_p),
                ('mesh_xi', ctypes.c_void_p), ('fpar', ctypes.c_void_p), ('order', ctypes.c_void_p)]


class CatintPnpControl(ctypes.Structure):
    _fields_ = [('mode', ctypes.c_int32), ('max_steps', ctypes.c_int32), ('n_out', ctypes.c_int32),
                ('polish_max_iter', ctypes.c_int32),
                ('rtol', ctypes.c_double), ('atol', ctypes.c_double), ('h0', ctypes.c_double),
                ('polish_rtol', ctypes.c_double), ('t_out', ctypes.POINTER(ctypes.c_double))]


EXPORTS = ['catint_pnp_version', 'catint_pnp_last_error', 'catint_pnp_device_count',
           'catint_pnp_workspace_bytes', 'catint_pnp_rhs_batch', 'catint_pnp_jacobian_batch',
           'catint_pnp_solve_batch', 'catint_pnp_debug_profile_buffer', 'catint_pnp_potential_batch',
           'catint_pnp_step_batch']


def library_path():
    # CATINT_PNP_LIB: development hook for A/B measurements of variant builds (scripts/build_variant.sh)
    return os.environ.get('CATINT_PNP_LIB') or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load_library():
    """dlopen the in-tree CUDA library (built by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.isfile(path):
        raise RuntimeError('catint_b200: %s not found; run `python -c "import __graft_entry__ as g; g.build()"` '
                           '(there is no CPU fallback)' % path)
    lib = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    lib.catint_pnp_version.restype = ctypes.c_int
    lib.catint_pnp_last_error.restype = ctypes.c_char_p
    lib.catint_pnp_device_count.restype = ctypes.c_int
    lib.catint_pnp_workspace_bytes.restype = ctypes.c_size_t
    lib.catint_pnp_workspace_bytes.argtypes = [ctypes.POINTER(CatintPnpShared), ctypes.c_int64]
    lib.catint_pnp_rhs_batch.restype = ctypes.c_int
    lib.catint_pnp_rhs_batch.argtypes = [ctypes.POINTER(CatintPnpShared), ctypes.POINTER(CatintPnpCells),
                                         ctypes.c_int64, vp, vp, vp, vp, vp]
    lib.catint_pnp_jacobian_batch.restype = ctypes.c_int
    lib.catint_pnp_jacobian_batch.argtypes = [ctypes.POINTER(CatintPnpShared), ctypes.POINTER(CatintPnpCells),
                                              ctypes.c_int64, vp, vp, vp, vp, vp, vp]
    lib.catint_pnp_solve_batch.restype = ctypes.c_int
    lib.catint_pnp_solve_batch.argtypes = [ctypes.POINTER(CatintPnpShared), ctypes.POINTER(CatintPnpCells),
                                           ctypes.c_int64, vp, ctypes.POINTER(CatintPnpControl),
                                           vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.catint_pnp_potential_batch.restype = ctypes.c_int
    lib.catint_pnp_potential_batch.argtypes = [ctypes.POINTER(CatintPnpShared), ctypes.POINTER(CatintPnpCells),
                                               ctypes.c_int64, vp, vp, vp, vp, vp]
    lib.catint_pnp_step_batch.restype = ctypes.c_int
    lib.catint_pnp_step_batch.argtypes = [ctypes.POINTER(CatintPnpShared), ctypes.POINTER(CatintPnpCells),
                                          ctypes.c_int64, vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                                          ctypes.c_int32, vp, ctypes.c_int32, vp, vp, vp, vp]
    lib.catint_pnp_debug_profile_buffer.restype = None
    lib.catint_pnp_debug_profile_buffer.argtypes = [vp]
    _lib = lib
    return lib


def npar(S):
    return 3 * S + 8


class CellBatch(object):
    """Structure-of-arrays description of B independent cells (host, numpy).

    species-level tables are shared by the batch; everything a descriptor can
    change is per cell (``par`` records, see include/catint_pnp.h).
    """

    def __init__(self, z, reactions, nu, par, nx, nx_max=None, use_migration=True,
                 poisson_bc=BC_DIRICHLET_WALL_NEUMANN_BULK, mesh_id=None, mesh_xi=None, species=None,
                 flux_eq=None, fpar=None):
        self.z = np.asarray(z, dtype=np.int32)
        self.S = len(self.z)
        self.reactions = [(list(e), list(p), float(kf), float(kr)) for (e, p, kf, kr) in reactions]
        self.R = len(self.reactions)
        self.nu = np.asarray(nu, dtype=np.float64).reshape(self.S, max(self.R, 0)) if self.R else np.zeros((self.S, 0))
        self.par = np.ascontiguousarray(par, dtype=np.float64)
        self.nx = np.ascontiguousarray(nx, dtype=np.int32)
        self.B = len(self.nx)
        assert self.par.shape == (self.B, npar(self.S)), (self.par.shape, self.B, npar(self.S))
        self.nx_max = int(nx_max if nx_max is not None else self.nx.max())
        self.use_migration = bool(use_migration)
        self.poisson_bc = int(poisson_bc)
        self.mesh_id = None if mesh_id is None else np.ascontiguousarray(mesh_id, dtype=np.int32)
        self.mesh_xi = None if mesh_xi is None else np.ascontiguousarray(mesh_xi, dtype=np.float64)
        self.species = list(species) if species is not None else None
        # wall kinetics as expressions (catint_b200/fluxeq.py: FluxEquations) and their per-cell parameters [B, n_par]
        self.flux_eq = flux_eq
        self.fpar = None
        if flux_eq is not None and flux_eq.n_eq > 0:
            self.fpar = np.ascontiguousarray(fpar, dtype=np.float64).reshape(self.B, len(flux_eq.par_names))
        self.origin = np.arange(self.B)          # cell indices in the batch this one was selected from
        if self.S > MAX_SPECIES or self.R > MAX_REACTIONS:
            raise ValueError('at most %d species and %d reactions' % (MAX_SPECIES, MAX_REACTIONS))
        if self.B and (int(self.nx.min()) < MIN_NODES or int(self.nx.max()) > self.nx_max):
            raise ValueError('every cell needs %d <= nx <= nx_max = %d nodes (got %d..%d)'
                             % (MIN_NODES, self.nx_max, int(self.nx.min()), int(self.nx.max())))
        if self.mesh_id is not None:
            if self.mesh_xi is None or self.mesh_xi.ndim != 2 or self.mesh_xi.shape[1] < self.nx_max:
                raise ValueError('mesh_xi must be [n_mesh, >= nx_max]')
            if self.B and int(self.mesh_id.max()) >= self.mesh_xi.shape[0]:
                raise ValueError('mesh_id out of range')

    def check_solvable(self):
        """limits of the integrator / Jacobian kernels (the RHS kernel takes up to MAX_SPECIES species)"""
        if not 2 <= self.b <= MAX_BLOCK:
            raise ValueError('catint_pnp_solve_batch handles 2..%d unknowns per node (S+1, or S+2 with the Stern '
                             'boundary); got %d' % (MAX_BLOCK, self.b))

    @property
    def b(self):
        return self.S + (2 if self.poisson_bc == BC_STERN_ROBIN else 1)

    def shared_struct(self):
        sh = CatintPnpShared()
        sh.S, sh.nx_max, sh.R = self.S, self.nx_max, self.R
        sh.poisson_bc, sh.use_migration = self.poisson_bc, int(self.use_migration)
        sh.n_mesh = 0 if self.mesh_xi is None else int(self.mesh_xi.shape[0])
        for k in range(self.S):
            sh.z[k] = int(self.z[k])
        for r in range(MAX_REACTIONS):
            for e in range(MAX_REACTANTS):
                sh.educt[r][e] = -1
                sh.product[r][e] = -1
        for r, (ed, pr, kf, kr) in enumerate(self.reactions):
            if len(ed) > MAX_REACTANTS or len(pr) > MAX_REACTANTS:
                raise ValueError('at most %d reactants per reaction side' % MAX_REACTANTS)
            for e, k in enumerate(ed):
                sh.educt[r][e] = int(k)
            for e, k in enumerate(pr):
                sh.product[r][e] = int(k)
            sh.kf[r], sh.kr[r] = kf, kr
            for k in range(self.S):
                sh.nu[k][r] = float(self.nu[k, r])
        self._fq_struct = None
        if self.flux_eq is not None and self.flux_eq.n_eq > 0:
            fq = CatintPnpFluxEq()
            fq.n_eq, fq.n_par = self.flux_eq.n_eq, len(self.flux_eq.par_names)
            for e, prog in enumerate(self.flux_eq.programs):
                fq.n_code[e] = len(prog.code)
                for k, w in enumerate(prog.code):
                    fq.code[e][k] = int(w)
                for k, v in enumerate(prog.consts):
                    fq.consts[e][k] = float(v)
            for k in range(self.S):
                for e in range(fq.n_eq):
                    fq.coef[k][e] = float(self.flux_eq.coef[k][e])
            self._fq_struct = fq                   # keep the host struct alive as long as the shared struct
            sh.flux_eq = ctypes.pointer(fq)
            sh._fq_keepalive = fq
        return sh

    def passive_species(self):
        """indices of the species a steady solve takes out of the block system (same rule as reduce_passive in
        csrc/pnp_capi.cu): no charge or migration off, no part in any homogeneous reaction, not read by a flux
        equation.  Host-side mirror for reporting (bench roofline, tests); the library decides on its own."""
        passive = [(int(self.z[k]) == 0 or not self.use_migration) for k in range(self.S)]
        for (ed, pr, kf, kr) in self.reactions:
            for k in list(ed) + list(pr):
                passive[k] = False
        for k in range(self.S):
            if self.R and np.any(self.nu[k] != 0.0):
                passive[k] = False
        if self.flux_eq is not None:
            for prog in self.flux_eq.programs:
                for w in prog.code:
                    if (w & 0xff) == 2:
                        passive[w >> 8] = False
        idx = [k for k in range(self.S) if passive[k]]
        return idx if 0 < len(idx) < self.S else []

    def launch_order(self):
        """cells sorted by predicted integration cost, expensive first (int32 permutation; longest processing time
        first).  The cost of a cell grows with the gradients its wall fluxes impose: sum_k |J_k| / D_k * L relative
        to the total bulk concentration (C2 sweep: 800 -> 1400 BDF steps along this number; the outliers with 1700
        steps sit where the Tafel current saturates and are not predictable from the inputs).  Only the schedule
        depends on it.  Measured on the 1024-cell C2 launch: -1.5 %; an SM-aware variant (expensive blocks placed on
        the 40 SMs that get no second block -- block j and j+148 do share an SM) gained nothing because the
        slowest cells are exactly those outliers."""
        S = self.S
        L = self.par[:, 3 * S + 5] * np.maximum(self.nx - 1, 1)
        if self.mesh_id is not None and self.mesh_xi is not None:
            L = self.par[:, 3 * S + 5] * self.mesh_xi[np.maximum(self.mesh_id, 0), np.maximum(self.nx - 1, 0)]
        grad = np.sum(np.abs(self.par[:, S:2 * S]) / np.maximum(self.par[:, 2 * S:3 * S], 1e-300), axis=1) * L
        cost = grad / np.maximum(np.sum(np.abs(self.par[:, 0:S]), axis=1), 1e-300) * (1.0 + 1e-3 * self.nx)
        return np.argsort(-cost, kind='stable').astype(np.int32)

    def select(self, idx):
        """sub-batch (used for sharding cells over ranks)."""
        idx = np.asarray(idx)
        sub = self._select(idx)
        sub.origin = self.origin[idx]
        return sub

    def _select(self, idx):
        return CellBatch(self.z, self.reactions, self.nu, self.par[idx], self.nx[idx], nx_max=self.nx_max,
                         use_migration=self.use_migration, poisson_bc=self.poisson_bc,
                         mesh_id=None if self.mesh_id is None else self.mesh_id[idx], mesh_xi=self.mesh_xi,
                         species=self.species, flux_eq=self.flux_eq,
                         fpar=None if self.fpar is None else self.fpar[idx])


def stoichiometry(S, reactions, rate_mode='summed'):
    """nu[S,R] with R_k = sum_r nu[k,r]*net_r.  'summed' = live semantics of the
    reference (catint/comsol_model.py:809-846); 'legacy_overwrite' = literal
    behaviour of catint/calculator.py:145-194 where every reactant first resets
    its rate, so only the last reaction touching a species survives, once."""
    nu = np.zeros((S, len(reactions)))
    for r, (ed, pr, kf, kr) in enumerate(reactions):
        if rate_mode == 'summed':
            for k in ed:
                nu[k, r] -= 1.0
            for k in pr:
                nu[k, r] += 1.0
        elif rate_mode == 'legacy_overwrite':
            for k in ed:
                nu[k, :] = 0.0
                nu[k, r] = -1.0
            for k in pr:
                nu[k, :] = 0.0
                nu[k, r] = +1.0
        else:
            raise ValueError(rate_mode)
    return nu


class DeviceBatch(object):
    """CellBatch uploaded to one GPU (torch tensors keep the memory alive)."""

    def __init__(self, batch, device, pinned=None):
        import torch
        self.batch = batch
        self.device = torch.device(device)
        src = pinned if pinned is not None else {}
        def up(name, arr, dtype):
            if arr is None:
                return None
            t = src.get(name)
            if t is None:
                t = torch.from_numpy(np.ascontiguousarray(arr))
            return t.to(self.device, dtype=dtype, non_blocking=True)
        self.par = up('par', batch.par, torch.float64)
        self.nx = up('nx', batch.nx, torch.int32)
        self.mesh_id = up('mesh_id', batch.mesh_id, torch.int32)
        self.mesh_xi = up('mesh_xi', batch.mesh_xi, torch.float64)
        self.fpar = up('fpar', batch.fpar if (batch.fpar is not None and batch.fpar.size) else None, torch.float64)
        self.shared = batch.shared_struct()
        self.cells = CatintPnpCells()
        self.cells.par = self.par.data_ptr()
        self.cells.nx = self.nx.data_ptr()
        self.cells.mesh_id = self.mesh_id.data_ptr() if self.mesh_id is not None else None
        self.cells.mesh_xi = self.mesh_xi.data_ptr() if self.mesh_xi is not None else None
        self.cells.fpar = self.fpar.data_ptr() if self.fpar is not None else None
        # launch order: expensive cells first (a scheduling hint, see CellBatch.launch_order)
        self.order = up('order', batch.launch_order() if batch.B > 1 else None, torch.int32)
        self.cells.order = self.order.data_ptr() if self.order is not None else None
        self.h2d_bytes = sum(int(t.numel() * t.element_size()) for t in
                             (self.par, self.nx, self.mesh_id, self.mesh_xi, self.fpar, self.order) if t is not None)


class PnpBackend(object):
    """Thin launcher: torch tensors in, torch tensors out, one CUDA stream."""

    def __init__(self, device=None):
        import torch
        self.torch = torch
        self.lib = load_library()
        if not torch.cuda.is_available() or self.lib.catint_pnp_device_count() <= 0:
            raise RuntimeError('catint_b200: no B200 (sm_100) device visible; the solver has no CPU fallback')
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self._ws = None
        self.launches = 0

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError('%s failed (%d): %s' % (what, rc, self.lib.catint_pnp_last_error().decode()))

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def upload(self, batch, pinned=None):
        with self.torch.cuda.device(self.device):
            return DeviceBatch(batch, self.device, pinned=pinned)

    def workspace(self, dbatch):
        need = self.lib.catint_pnp_workspace_bytes(ctypes.byref(dbatch.shared), dbatch.batch.B)
        if self._ws is None or self._ws.numel() < need:
            self._ws = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        return self._ws, need

    # -- K1 ---------------------------------------------------------------
    def rhs(self, dbatch, c):
        """c [B,nx_max,S] (device, float64) -> (dcdt, g, phi)"""
        torch = self.torch
        b = dbatch.batch
        assert c.shape == (b.B, b.nx_max, b.S) and c.dtype == torch.float64 and c.is_contiguous()
        dcdt = torch.zeros_like(c)
        g = torch.zeros((b.B, b.nx_max), dtype=torch.float64, device=self.device)
        phi = torch.zeros_like(g)
        with torch.cuda.device(self.device):
            rc = self.lib.catint_pnp_rhs_batch(ctypes.byref(dbatch.shared), ctypes.byref(dbatch.cells), b.B,
                                               c.data_ptr(), dcdt.data_ptr(), g.data_ptr(), phi.data_ptr(),
                                               self._stream())
        self._check(rc, 'catint_pnp_rhs_batch')
        self.launches += 1
        return dcdt, g, phi

    # -- Poisson routine / K4 ----------------------------------------------
    def potential(self, dbatch, c):
        """get_potential_and_gradient for every pb_bound pair: c [B,nx_max,S] -> (v, grad_v, lapl_v) [B,nx_max]"""
        torch = self.torch
        b = dbatch.batch
        assert c.shape == (b.B, b.nx_max, b.S) and c.dtype == torch.float64 and c.is_contiguous()
        v = torch.zeros((b.B, b.nx_max), dtype=torch.float64, device=self.device)
        g, lp = torch.zeros_like(v), torch.zeros_like(v)
        with torch.cuda.device(self.device):
            rc = self.lib.catint_pnp_potential_batch(ctypes.byref(dbatch.shared), ctypes.byref(dbatch.cells), b.B,
                                                     c.data_ptr(), v.data_ptr(), g.data_ptr(), lp.data_ptr(),
                                                     self._stream())
        self._check(rc, 'catint_pnp_potential_batch')
        self.launches += 1
        return v, g, lp

    def step(self, dbatch, stepper, dt, nt, itout, c0=None, lax_friedrich=False):
        """the reference's fixed-step steppers (FTCS / Crank-Nicolson): nt steps of size dt; returns dict with
        c [n_out,B,nx_max,S], phi, g [n_out,B,nx_max] after the steps listed in itout"""
        torch = self.torch
        b = dbatch.batch
        itout = np.ascontiguousarray(itout, dtype=np.int32)
        n_out = len(itout)
        if c0 is None:
            c0 = torch.as_tensor(np.ascontiguousarray(
                np.broadcast_to(b.par[:, None, 0:b.S], (b.B, b.nx_max, b.S)))).to(self.device)
        assert c0.shape == (b.B, b.nx_max, b.S) and c0.dtype == torch.float64 and c0.is_contiguous()
        it_dev = torch.as_tensor(itout).to(self.device)
        out = {'c': torch.zeros((n_out, b.B, b.nx_max, b.S), dtype=torch.float64, device=self.device),
               'phi': torch.zeros((n_out, b.B, b.nx_max), dtype=torch.float64, device=self.device),
               'g': torch.zeros((n_out, b.B, b.nx_max), dtype=torch.float64, device=self.device)}
        with torch.cuda.device(self.device):
            rc = self.lib.catint_pnp_step_batch(ctypes.byref(dbatch.shared), ctypes.byref(dbatch.cells), b.B,
                                                c0.data_ptr(), int(stepper), int(bool(lax_friedrich)), float(dt),
                                                int(nt), it_dev.data_ptr(), n_out, out['c'].data_ptr(),
                                                out['phi'].data_ptr(), out['g'].data_ptr(), self._stream())
        self._check(rc, 'catint_pnp_step_batch')
        self.launches += 1
        return out

    # -- K2 ---------------------------------------------------------------
    def jacobian(self, dbatch, y):
        """y [B,nx_max,b] -> (F [B,nx,b], Lb, Db, Ub [B,nx,b,b])"""
        torch = self.torch
        b = dbatch.batch
        nb = b.b
        b.check_solvable()
        assert y.shape == (b.B, b.nx_max, nb) and y.dtype == torch.float64 and y.is_contiguous()
        F = torch.zeros_like(y)
        blocks = [torch.zeros((b.B, b.nx_max, nb, nb), dtype=torch.float64, device=self.device) for _ in range(3)]
        with torch.cuda.device(self.device):
            rc = self.lib.catint_pnp_jacobian_batch(ctypes.byref(dbatch.shared), ctypes.byref(dbatch.cells), b.B,
                                                    y.data_ptr(), F.data_ptr(), blocks[0].data_ptr(),
                                                    blocks[1].data_ptr(), blocks[2].data_ptr(), self._stream())
        self._check(rc, 'catint_pnp_jacobian_batch')
        self.launches += 1
        return F, blocks[0], blocks[1], blocks[2]

    # -- K3 ---------------------------------------------------------------
    def alloc_outputs(self, dbatch, n_out):
        torch = self.torch
        b = dbatch.batch
        dev = self.device
        return {
            'c': torch.zeros((n_out, b.B, b.nx_max, b.S), dtype=torch.float64, device=dev),
            'phi': torch.zeros((n_out, b.B, b.nx_max), dtype=torch.float64, device=dev),
            'g': torch.zeros((n_out, b.B, b.nx_max), dtype=torch.float64, device=dev),
            'flux': torch.zeros((b.B, b.S), dtype=torch.float64, device=dev),
            'status': torch.full((b.B,), -1, dtype=torch.int32, device=dev),
            'n_steps': torch.zeros((b.B,), dtype=torch.int32, device=dev),
            'n_newton': torch.zeros((b.B,), dtype=torch.int32, device=dev),
            'n_setups': torch.zeros((b.B,), dtype=torch.int32, device=dev),
        }

    def solve(self, dbatch, t_out, mode=MODE_STEADY, rtol=1.49012e-8, atol=1.49012e-8, y0=None,
              max_steps=100000, h0=0.0, polish_rtol=1e-10, polish_max_iter=8, out=None):
        """returns dict of device tensors: c [n_out,B,nx,S], phi, g [n_out,B,nx],
        flux [B,S], status, n_steps, n_newton [B] (``out``: reuse a dict from alloc_outputs)"""
        torch = self.torch
        b = dbatch.batch
        t_out = np.ascontiguousarray(np.atleast_1d(np.asarray(t_out, dtype=np.float64)))
        n_out = len(t_out)
        b.check_solvable()
        if n_out > MAX_OUTPUT_TIMES:
            raise ValueError('at most %d output times per solve' % MAX_OUTPUT_TIMES)
        dev = self.device
        if out is None:
            out = self.alloc_outputs(dbatch, n_out)
        assert out['c'].shape == (n_out, b.B, b.nx_max, b.S)
        ws, need = self.workspace(dbatch)
        ctl = CatintPnpControl()
        ctl.mode, ctl.max_steps, ctl.n_out, ctl.polish_max_iter = int(mode), int(max_steps), n_out, int(polish_max_iter)
        ctl.rtol, ctl.atol, ctl.h0, ctl.polish_rtol = float(rtol), float(atol), float(h0), float(polish_rtol)
        ctl.t_out = t_out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        if y0 is not None:
            assert y0.shape == (b.B, b.nx_max, b.S) and y0.dtype == torch.float64 and y0.is_contiguous()
        with torch.cuda.device(dev):
            rc = self.lib.catint_pnp_solve_batch(
                ctypes.byref(dbatch.shared), ctypes.byref(dbatch.cells), b.B,
                y0.data_ptr() if y0 is not None else None, ctypes.byref(ctl),
                out['c'].data_ptr(), out['phi'].data_ptr(), out['g'].data_ptr(), out['flux'].data_ptr(),
                out['status'].data_ptr(), out['n_steps'].data_ptr(), out['n_newton'].data_ptr(),
                out['n_setups'].data_ptr(),
                ws.data_ptr(), ctypes.c_size_t(need), self._stream())
        self._check(rc, 'catint_pnp_solve_batch')
        self.launches += 1
        return out
