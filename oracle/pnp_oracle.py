"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (numpy / scipy) of CatINT's 1D finite-difference
Poisson-Nernst-Planck right-hand side and its scipy ``odeint`` driver.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module, and only as the checker or
as the timed CPU baseline -- never as something the product routes through.

What it restates (all citations into /root/reference):

* RHS ``ode_func``                       catint/calculator_old.py:827-935
* Poisson by two cumulative sums          catint/calculator_old.py:680-819
  (default BCs: potential fixed at the wall, gradient fixed in the bulk,
  ``pb_bound`` default catint/transport.py:207-210)
* homogeneous mass-action rates           catint/calculator_old.py:159-208
  (= catint/calculator.py:145-194), with both semantics:
    - ``legacy_overwrite``: literal behaviour of that code (every reactant
      resets ``rates[k,i]=0.0`` first, so only the last reaction touching a
      species survives)
    - ``summed``: the live semantics of catint/comsol_model.py:809-846 where
      those resets are commented out (default of the product)
* integrator call                         catint/calculator_old.py:946-948
  (``scipy.integrate.odeint``; the reference passes ml=mu=nspecies which only
  works for pure diffusion -- SURVEY 0-4 -- so the oracle uses the dense
  finite-difference Jacobian, i.e. no ml/mu)
* result containers                       catint/calculator_old.py:816-818,966-973

Sign convention: the reference's wall stencil *subtracts* ``flux_bound``
(catint/calculator_old.py:902-909) whereas the current Transport gives
products a positive flux (catint/transport.py:1023-1032).  ``PnpSystem.J`` is
the INWARD flux (positive = species enters the electrolyte at the wall); the
literal reference behaviour is therefore ``J = -flux_bound[:,0]``.  The
product uses ``J = +species[sp]['flux']`` (docs/source/topics/flux_definition.rst:60-65).

Pinning: the reference ships no tests or golden vectors and its FD integrator
is orphaned Python 2, so the odeint *trajectory* is "parity unpinned" by the
reference's own tests.  What IS pinned (tests/golden/make_golden.py, run in
the build container against /root/reference):
  - the RHS: this module's ``rhs`` is checked against the reference's own
    ``ode_func`` (calculator_old.py source loaded and made importable in
    memory, evaluated on the reference Transport's arrays);
  - the setup arrays (c0, flux_bound, D, charges, beta, eps, xmesh, bulk
    equilibria) against the reference ``Transport``.
Third-party arithmetic: scipy.integrate.odeint (ODEPACK LSODA), unpinned in
the reference (no requirements file); here scipy 1.18.1 / numpy 2.3.5.

Extension beyond the reference FD code: non-uniform meshes and a
Stern-layer Robin boundary for the Poisson equation (the latter only exists
in the reference's COMSOL path, catint/comsol_model.py:613,982).  Both reduce
to the reference stencil on a uniform mesh with the default BCs.
"""
import numpy as np

# catint/units.py:4,13,16
UNIT_R = 8.3144598
UNIT_EPS0 = 8.854187817e-12
UNIT_F = 96485.33289


class PnpSystem(object):
    """Plain-array description of one cell (one sweep point).

    z          integer charges [S]            (catint/transport.py:1240-1276)
    D          diffusion coefficients [S]     (catint/transport.py:423-434)
    c_bulk     bulk concentrations [S] mol/m^3
    J          inward wall flux [S] mol/m^2/s
    x          mesh [n] (uniform: np.arange(0,xmax+dx,dx), transport.py:459)
    beta       1/(R T)                        (catint/transport.py:312)
    eps        eps_r*eps_0                    (catint/transport.py:311)
    reactions  list of (educt_idx, product_idx, kf, kr); index lists hold one
               entry per stoichiometric unit, species that are not transported
               (H2O, e-) already dropped     (catint/transport.py:1098-1132)
    rate_mode  'summed' | 'legacy_overwrite'
    poisson_bc 'dirichlet_wall_neumann_bulk' (phi_wall, g_bulk) or
               'stern_robin' (phiM, phiPZC, C_stern in F/m^2, phi_bulk=0)
    """

    def __init__(self, z, D, c_bulk, J, x, beta, eps, reactions=(),
                 rate_mode='summed', use_migration=True,
                 poisson_bc='dirichlet_wall_neumann_bulk',
                 phi_wall=0.0, g_bulk=0.0, phiM=0.0, phiPZC=0.0, C_stern=0.2,
                 uniform=None, wall_kinetics=None):
        self.z = np.asarray(z, dtype=float)
        self.q = self.z * UNIT_F
        self.D = np.asarray(D, dtype=float)
        self.c_bulk = np.asarray(c_bulk, dtype=float)
        self.J = np.asarray(J, dtype=float)
        self.x = np.asarray(x, dtype=float)
        self.S = len(self.z)
        self.n = len(self.x)
        self.beta = float(beta)
        self.eps = float(eps)
        self.reactions = [(list(e), list(p), float(kf), float(kr)) for (e, p, kf, kr) in reactions]
        self.rate_mode = rate_mode
        self.use_migration = bool(use_migration)
        self.poisson_bc = poisson_bc
        self.phi_wall = float(phi_wall)
        self.g_bulk = float(g_bulk)
        self.phiM = float(phiM)
        self.phiPZC = float(phiPZC)
        self.C_stern = float(C_stern)
        self.h = np.diff(self.x)
        if uniform is None:
            uniform = bool(np.allclose(self.h, self.h[0], rtol=1e-12, atol=0.0))
        self.uniform = uniform
        # the reference works with the scalar dx=xmax/nx (transport.py:452)
        self.dx = float(self.h[0])
        self.b = self.S + (2 if poisson_bc == 'stern_robin' else 1)
        # optional flux equations (oracle/flux_expr.py: WallKinetics): J = self.J + wall_kinetics(c(0), phi(0))
        self.wall_kinetics = wall_kinetics

    def wall_flux(self, C, v):
        """inward wall flux of the state C[S,n] (v: potential, only its wall value is used)"""
        if self.wall_kinetics is None:
            return self.J
        phi0 = v[0] if self.use_migration else self.phi_wall
        return self.J + self.wall_kinetics(C[:, 0], phi0)

    # ------------------------------------------------------------------
    def c0_flat(self):
        """species-major initial state c0[k*n+i]=c_bulk[k] (transport.py:1396-1412)"""
        return np.repeat(self.c_bulk, self.n)

    # ------------------------------------------------------------------
    def rates(self, C):
        """homogeneous rates R[S,n] (calculator_old.py:159-208)."""
        S, n = C.shape
        R = np.zeros((S, n))
        if self.rate_mode == 'summed':
            for (ed, pr, kf, kr) in self.reactions:
                net = kf * np.prod(C[ed, :], axis=0) if len(ed) else kf * np.ones(n)
                net = net - (kr * np.prod(C[pr, :], axis=0) if len(pr) else kr * np.ones(n))
                for k in ed:
                    R[k] -= net
                for k in pr:
                    R[k] += net
        elif self.rate_mode == 'legacy_overwrite':
            for (ed, pr, kf, kr) in self.reactions:
                fwd = kf * np.prod(C[ed, :], axis=0) if len(ed) else kf * np.ones(n)
                bwd = kr * np.prod(C[pr, :], axis=0) if len(pr) else kr * np.ones(n)
                for k in ed:
                    R[k] = 0.0
                    R[k] -= fwd
                    R[k] += bwd
                for k in pr:
                    R[k] = 0.0
                    R[k] += fwd
                    R[k] -= bwd
        else:
            raise ValueError(self.rate_mode)
        return R

    # ------------------------------------------------------------------
    def poisson(self, C):
        """(v, grad_v, lapl_v) as in get_potential_and_gradient
        (calculator_old.py:680-819), default BC combination: gradient given in
        the bulk (:793-796), potential given at the wall (:798-800).
        Non-uniform extension: g_i=g_{i+1}-rho_i*h_i, v_i=v_{i-1}+g_i*h_{i-1}."""
        n = self.n
        lapl = np.zeros(n)
        for k in range(self.S):                             # :767-771, same summation order
            lapl -= self.q[k] * C[k, :] / self.eps
        g = np.zeros(n)
        v = np.zeros(n)
        h = self.h
        if self.poisson_bc == 'dirichlet_wall_neumann_bulk':
            g[n - 1] = self.g_bulk                          # :794
            for i in range(n - 2, 0, -1):                   # :753-759
                hi = self.dx if self.uniform else h[i]
                g[i] = g[i + 1] - lapl[i] * hi
            if self.uniform:
                g[0] = g[1] + (g[1] - g[2])                 # :796
            else:
                g[0] = g[1] + (g[1] - g[2]) * h[0] / h[1]
            v[0] = self.phi_wall                            # :777
            for i in range(1, n - 1):                       # :760-761
                hi = self.dx if self.uniform else h[i - 1]
                v[i] = v[i - 1] + g[i] * hi
            if self.uniform:
                v[n - 1] = v[n - 2] + (v[n - 2] - v[n - 3])  # :800
            else:
                v[n - 1] = v[n - 2] + (v[n - 2] - v[n - 3]) * h[n - 2] / h[n - 3]
        else:
            # Stern/Robin wall, phi(L)=0.  g_i=g_{i+1}-lapl_i*h_i for
            # i=n-2..1, g_0 by linear extrapolation, phi_i=phi_{i-1}+g_i*h_{i-1}
            # for i=1..n-1, phi_{n-1}=0, eps*g_0=-C_S*((phiM-phiPZC)-phi_0).
            # Everything is affine in the unknown g_{n-1}: solve for it.
            def chain(gb):
                gg = np.zeros(n)
                gg[n - 1] = gb
                for i in range(n - 2, 0, -1):
                    gg[i] = gg[i + 1] - lapl[i] * h[i]
                gg[0] = gg[1] + (gg[1] - gg[2]) * h[0] / h[1]
                # integrate phi backwards from phi_{n-1}=0
                pp = np.zeros(n)
                for i in range(n - 1, 0, -1):
                    pp[i - 1] = pp[i] - gg[i] * h[i - 1]
                res = self.eps * gg[0] + self.C_stern * ((self.phiM - self.phiPZC) - pp[0])
                return gg, pp, res
            _, _, r0 = chain(0.0)
            _, _, r1 = chain(1.0)
            gb = -r0 / (r1 - r0)
            g, v, _ = chain(gb)
        return v, g, lapl

    # ------------------------------------------------------------------
    def rhs(self, c, with_field=False):
        """dc/dt, species-major flat (calculator_old.py:827-935), vectorised."""
        S, n = self.S, self.n
        C = np.asarray(c, dtype=float).reshape(S, n)
        if self.use_migration:
            v, g, lapl = self.poisson(C)
        else:
            v = np.zeros(n); g = np.zeros(n); lapl = np.zeros(n)   # :893-896
        R = self.rates(C) if len(self.reactions) else np.zeros((S, n))
        dC = np.zeros((S, n))
        D = self.D[:, None]
        bq = (self.beta * self.q)[:, None]
        Cg = C * g[None, :]
        if self.uniform:
            dx = self.dx
            d2 = (C[:, 2:] - 2 * C[:, 1:-1] + C[:, :-2]) / (dx ** 2)            # :890
            dcg = (Cg[:, 2:] - Cg[:, :-2]) / (2. * dx)                           # :892
            dC[:, 1:-1] = D * (d2 + bq * dcg) + R[:, 1:-1]                       # :920-927
            # wall node, one-sided stencil reaching node 2, no reaction term (:902-915)
            dC[:, 0] = (self.D * ((C[:, 2] - C[:, 0]) / (2. * dx)
                                  + self.beta * self.q * C[:, 1] * g[1]) + self.wall_flux(C, v)) / dx
        else:
            h = self.h
            hm = h[:-1][None, :]
            hp = h[1:][None, :]
            d2 = 2.0 * ((C[:, 2:] - C[:, 1:-1]) / hp - (C[:, 1:-1] - C[:, :-2]) / hm) / (hm + hp)
            dcg = (Cg[:, 2:] - Cg[:, :-2]) / (hm + hp)
            dC[:, 1:-1] = D * (d2 + bq * dcg) + R[:, 1:-1]
            dC[:, 0] = (self.D * ((C[:, 2] - C[:, 0]) / (h[0] + h[1])
                                  + self.beta * self.q * C[:, 1] * g[1]) + self.wall_flux(C, v)) / h[0]
        dC[:, n - 1] = 0.0                                                       # :886
        if with_field:
            return dC.reshape(-1), v, g, lapl
        return dC.reshape(-1)

    # ------------------------------------------------------------------
    def rhs_literal(self, c):
        """Loop-literal transcription of ode_func's control flow
        (calculator_old.py:884-927) for the uniform mesh / default BCs; used
        to cross-check ``rhs`` and as the 'what the reference would execute'
        CPU timing.  Pure Python loops: small cases only."""
        assert self.uniform and self.poisson_bc == 'dirichlet_wall_neumann_bulk'
        S, n, dx = self.S, self.n, self.dx
        C = np.zeros([S, n])
        for k in range(S):
            C[k, :] = c[k * n:(k + 1) * n]
        if self.use_migration:
            v, grad_v, lapl_v = self.poisson(C)
        else:
            grad_v = [0.0] * n
        rates = self.rates(C) if len(self.reactions) else np.zeros([S, n])
        DC_DT = np.zeros([S, n])
        for k in range(S):
            DC_DT[k, -1] = 0.0
            for i in range(0, n - 1):
                dc_dx_2 = (C[k, i + 1] - 2 * C[k, i] + C[k, i - 1]) / (dx ** 2)
                if self.use_migration:
                    dcgradv_dx = (C[k, i + 1] * grad_v[i + 1] - C[k, i - 1] * grad_v[i - 1]) / (2. * dx)
                else:
                    dcgradv_dx = 0.0
                if i == 0:
                    DC_DT[k, i] = (self.D[k] * ((C[k, 2] - C[k, 0]) / (2. * dx)
                                                + self.beta * self.q[k] * C[k, 1] * grad_v[1])
                                   + self.J[k]) / dx
                else:
                    DC_DT[k, i] = self.D[k] * (dc_dx_2 + self.beta * self.q[k] * dcgradv_dx) + rates[k, i]
        out = np.zeros([S * n])
        for k in range(S):
            out[k * n:(k + 1) * n] = DC_DT[k, :]
        return out

    # ------------------------------------------------------------------
    def integrate_odeint(self, tmesh, c0=None, rtol=None, atol=None, mxstep=5000000,
                         full_output=False):
        """scipy odeint on the restated RHS (calculator_old.py:946-948) with the
        dense finite-difference Jacobian (no ml/mu).  Returns the
        solution array [len(tmesh), S*n] (row n = state at tmesh[n]; the
        reference keeps rows listed in tp.itout, :968-973)."""
        from scipy.integrate import odeint
        if c0 is None:
            c0 = self.c0_flat()
        kw = {}
        if rtol is not None:
            kw['rtol'] = rtol
        if atol is not None:
            kw['atol'] = atol
        sol, info = odeint(lambda c, t: self.rhs(c), c0, tmesh, full_output=True,
                           mxstep=mxstep, **kw)
        if full_output:
            return sol, info
        return sol

    # ------------------------------------------------------------------
    def wall_flux_discrete(self, C):
        """discrete inward flux that makes dc/dt(node 0)=0, i.e.
        -D_k((c_2-c_0)/(h0+h1)+beta q_k c_1 g_1); equals J at a steady state."""
        _, g, _ = self.poisson(C) if self.use_migration else (None, np.zeros(self.n), None)
        h = self.h
        den = 2. * self.dx if self.uniform else (h[0] + h[1])
        return -self.D * ((C[:, 2] - C[:, 0]) / den + self.beta * self.q * C[:, 1] * g[1])


def steady_tmesh(t_end=200.0, t_first=1e-9, npts=60):
    """geometric output mesh 1e-9...t_end used for steady-state oracle runs
    (SURVEY 8d, C1)."""
    return np.concatenate([[0.0], np.geomspace(t_first, t_end, npts)])
