"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Local (block-tridiagonal) form of the restated CatINT FD-PNP system and a CPU
sparse Newton / pseudo-transient solver for its steady state (oracle ladder
step (ii)).  It is the large-grid / Stern-extension checker; it is
itself validated against ``pnp_oracle.PnpSystem.integrate_odeint`` on small
grids (tests/test_oracle.py).

Derivation: carry g=dphi/dx as an extra per-node unknown so that
the two cumulative sums of the reference's Poisson solve
(/root/reference/catint/calculator_old.py:732-764,793-800) become the local
rows  g_i - g_{i+1} + rho_i*h_i = 0.  Unknowns per node y_i=(c_0..c_{S-1}, g
[, phi]); rows per node = transport rows of calculator_old.py:884-927 plus the
algebraic g (and phi) rows.  The matrix is block tridiagonal plus one extra
block (0,2) (wall stencil :902-909 and the extrapolation g_0=2g_1-g_2 :796).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .pnp_oracle import PnpSystem


def stoich_matrix(sys_):
    """nu[S,R] such that R_k = sum_r nu[k,r]*net_r reproduces PnpSystem.rates
    in either mode (legacy-overwrite is 'just a different stoichiometry
    table')."""
    S = sys_.S
    R = len(sys_.reactions)
    nu = np.zeros((S, R))
    if sys_.rate_mode == 'summed':
        for r, (ed, pr, kf, kr) in enumerate(sys_.reactions):
            for k in ed:
                nu[k, r] -= 1.0
            for k in pr:
                nu[k, r] += 1.0
    else:
        for r, (ed, pr, kf, kr) in enumerate(sys_.reactions):
            for k in ed:
                nu[k, :] = 0.0
                nu[k, r] = -1.0
            for k in pr:
                nu[k, :] = 0.0
                nu[k, r] = +1.0
    return nu


class LocalForm(object):
    def __init__(self, sys_: PnpSystem):
        self.s = sys_
        self.nu = stoich_matrix(sys_)
        self.S = sys_.S
        self.n = sys_.n
        self.b = sys_.b
        self.stern = sys_.poisson_bc == 'stern_robin'
        s = sys_
        n = s.n
        h = s.h
        # per-node stencil coefficients (interior i=1..n-2)
        self.am = np.zeros(n); self.ap = np.zeros(n); self.ac = np.zeros(n)
        if s.uniform:
            dx = s.dx
            self.am[1:-1] = 1.0 / dx ** 2
            self.ap[1:-1] = 1.0 / dx ** 2
            self.ac[1:-1] = 1.0 / (2.0 * dx)
            self.hi = np.full(n, dx)        # h_i     (node i -> i+1)
            self.him = np.full(n, dx)       # h_{i-1} (node i-1 -> i)
            self.w0 = 1.0 / (2.0 * dx)      # wall gradient 1/(h0+h1)
            self.ih0 = 1.0 / dx
            self.ext = 1.0                  # h0/h1
        else:
            hm = h[:-1]; hp = h[1:]
            self.am[1:-1] = 2.0 / (hm * (hm + hp))
            self.ap[1:-1] = 2.0 / (hp * (hm + hp))
            self.ac[1:-1] = 1.0 / (hm + hp)
            self.hi = np.concatenate([h, [h[-1]]])
            self.him = np.concatenate([[h[0]], h])
            self.w0 = 1.0 / (h[0] + h[1])
            self.ih0 = 1.0 / h[0]
            self.ext = h[0] / h[1]

    # ------------------------------------------------------------------
    def y_from_c(self, C):
        """consistent y [n,b] from concentrations C[S,n] (solves the algebraic rows)."""
        s = self.s
        y = np.zeros((self.n, self.b))
        y[:, :self.S] = C.T
        if s.use_migration:
            v, g, _ = s.poisson(C)
            y[:, self.S] = g
            if self.stern:
                y[:, self.S + 1] = v
        return y

    def y_bulk(self):
        return self.y_from_c(np.repeat(self.s.c_bulk[:, None], self.n, axis=1))

    # ------------------------------------------------------------------
    def net_rates(self, c):
        """net_r [R,n] and d net_r / d c_j [R,S,n] for c[n,S]."""
        s = self.s
        n = c.shape[0]
        R = len(s.reactions)
        net = np.zeros((R, n))
        dnet = np.zeros((R, self.S, n))
        for r, (ed, pr, kf, kr) in enumerate(s.reactions):
            f = np.full(n, kf)
            for k in ed:
                f = f * c[:, k]
            bwd = np.full(n, kr)
            for k in pr:
                bwd = bwd * c[:, k]
            net[r] = f - bwd
            for p, k in enumerate(ed):
                t = np.full(n, kf)
                for q_, k2 in enumerate(ed):
                    if q_ != p:
                        t = t * c[:, k2]
                dnet[r, k] += t
            for p, k in enumerate(pr):
                t = np.full(n, kr)
                for q_, k2 in enumerate(pr):
                    if q_ != p:
                        t = t * c[:, k2]
                dnet[r, k] -= t
        return net, dnet

    # ------------------------------------------------------------------
    def residual(self, y, blocks=False):
        """F[n,b]: transport rows = dc/dt, algebraic rows = constraint.
        With blocks=True also returns (L,Dg,U,E0) = dF_i/dy_{i-1}, dy_i,
        dy_{i+1} [n,b,b] and dF_0/dy_2 [b,b]."""
        s = self.s
        S, n, b = self.S, self.n, self.b
        c = y[:, :S]
        g = y[:, S] if s.use_migration else np.zeros(n)
        D = s.D[None, :]
        bq = (s.beta * s.q)[None, :]
        F = np.zeros((n, b))
        iS = np.arange(S)
        have_rx = len(s.reactions) > 0
        if have_rx:
            net, dnet = self.net_rates(c)
            Rk = (self.nu @ net).T                    # [n,S]
        else:
            Rk = np.zeros((n, S))
        am = self.am[1:-1, None]; ap = self.ap[1:-1, None]; ac = self.ac[1:-1, None]
        cg = c * g[:, None]
        F[1:-1, :S] = D * (am * c[:-2] - (am + ap) * c[1:-1] + ap * c[2:]
                           + bq * (cg[2:] - cg[:-2]) * ac) + Rk[1:-1]
        Jw = s.J
        if s.wall_kinetics is not None:
            phi0 = y[0, S + 1] if self.stern else s.phi_wall
            Jw = s.J + s.wall_kinetics(c[0], phi0)
        F[0, :S] = (s.D * ((c[2] - c[0]) * self.w0 + s.beta * s.q * c[1] * g[1]) + Jw) * self.ih0
        F[n - 1, :S] = s.c_bulk - c[n - 1]            # bulk Dirichlet (frozen node, :886)
        gg = y[:, S]
        if s.use_migration:
            rho = (c @ s.q) / s.eps                   # = -lapl_v
            F[1:-1, S] = -(gg[1:-1] - gg[2:] - rho[1:-1] * self.hi[1:-1])
            F[0, S] = -(gg[0] - gg[1] - (gg[1] - gg[2]) * self.ext)
            if not self.stern:
                F[n - 1, S] = s.g_bulk - gg[n - 1]
            else:
                ph = y[:, S + 1]
                F[1:-1, S + 1] = -(ph[1:-1] - ph[:-2] - gg[1:-1] * self.him[1:-1])
                F[n - 1, S] = -(ph[n - 1] - ph[n - 2] - gg[n - 1] * self.him[n - 1])
                F[n - 1, S + 1] = -ph[n - 1]
                F[0, S + 1] = -((s.eps / s.C_stern) * gg[0] + (s.phiM - s.phiPZC) - ph[0])
        else:
            F[:, S] = -gg
            if self.stern:
                F[:, S + 1] = -y[:, S + 1]
        if not blocks:
            return F
        L = np.zeros((n, b, b)); Dg = np.zeros((n, b, b)); U = np.zeros((n, b, b)); E0 = np.zeros((b, b))
        I = slice(1, n - 1)
        # interior transport rows
        L[I, iS, iS] = (D * (am - bq * g[:-2, None] * ac))
        U[I, iS, iS] = (D * (ap + bq * g[2:, None] * ac))
        Dg[I, iS, iS] = (-D * (am + ap))
        if s.use_migration:
            L[I, iS, S] = -(D * bq * c[:-2] * ac)
            U[I, iS, S] = (D * bq * c[2:] * ac)
        if have_rx:
            # dR_k/dc_j = sum_r nu[k,r] dnet[r,j]
            dR = np.einsum('kr,rjn->nkj', self.nu, dnet)
            Dg[I, :S, :S] += dR[I]
        # wall transport rows
        Dg[0, iS, iS] = -s.D * self.w0 * self.ih0
        if s.wall_kinetics is not None:
            dJc, dJphi = s.wall_kinetics.jacobian(c[0], y[0, S + 1] if self.stern else s.phi_wall)
            Dg[0, :S, :S] += dJc * self.ih0
            if self.stern:
                Dg[0, :S, S + 1] += dJphi * self.ih0
        E0[iS, iS] = s.D * self.w0 * self.ih0
        U[0, iS, iS] = s.D * s.beta * s.q * g[1] * self.ih0
        if s.use_migration:
            U[0, iS, S] = s.D * s.beta * s.q * c[1] * self.ih0
        # bulk rows
        Dg[n - 1, iS, iS] = -1.0
        # algebraic rows (note F = -(constraint) so dF = -d(constraint))
        if s.use_migration:
            Dg[I, S, S] = -1.0
            U[I, S, S] = 1.0
            Dg[I, S, :S] = (s.q[None, :] / s.eps) * self.hi[1:-1, None]
            Dg[0, S, S] = -1.0
            U[0, S, S] = 1.0 + self.ext
            E0[S, S] = -self.ext
            if not self.stern:
                Dg[n - 1, S, S] = -1.0
            else:
                P = S + 1
                Dg[I, P, P] = -1.0
                L[I, P, P] = 1.0
                Dg[I, P, S] = self.him[1:-1]
                Dg[n - 1, S, P] = -1.0
                L[n - 1, S, P] = 1.0
                Dg[n - 1, S, S] = self.him[n - 1]
                Dg[n - 1, P, P] = -1.0
                Dg[0, P, S] = -(s.eps / s.C_stern)
                Dg[0, P, P] = 1.0
        else:
            Dg[:, S, S] = -1.0
            if self.stern:
                Dg[:, S + 1, S + 1] = -1.0
        return F, L, Dg, U, E0

    # ------------------------------------------------------------------
    def to_sparse(self, L, Dg, U, E0):
        n, b = self.n, self.b
        rows = []; cols = []; vals = []
        rb, cb = np.meshgrid(np.arange(b), np.arange(b), indexing='ij')
        for (blk, off) in ((L, -1), (Dg, 0), (U, 1)):
            i0 = max(0, -off); i1 = n - max(0, off)
            for i in range(i0, i1):
                m = blk[i]
                nz = m != 0
                rows.append((i * b + rb[nz])); cols.append(((i + off) * b + cb[nz])); vals.append(m[nz])
        nz = E0 != 0
        rows.append(rb[nz]); cols.append(2 * b + cb[nz]); vals.append(E0[nz])
        return sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                             shape=(n * b, n * b))

    def mass_diag(self):
        m = np.zeros((self.n, self.b))
        m[:-1, :self.S] = 1.0          # transport rows of nodes 0..n-2 carry d/dt
        return m

    # ------------------------------------------------------------------
    def weights(self, y, rtol, atol_c, atol_g):
        w = np.empty_like(y)
        w[:, :self.S] = rtol * np.abs(y[:, :self.S]) + atol_c
        w[:, self.S:] = rtol * np.abs(y[:, self.S:]) + atol_g
        return w

    def solve_steady(self, y0=None, dt0=1e-6, rtol=1e-10, max_steps=400, growth=8.0,
                     verbose=False, pure_newton=False):
        """pseudo-transient continuation (implicit Euler, one Newton iteration per
        step, switched-evolution-relaxation dt control) to the root F(y)=0,
        finished by full Newton steps.  Returns (y, info)."""
        s = self.s
        y = self.y_bulk() if y0 is None else y0.copy()
        M = self.mass_diag()
        cscale = max(np.max(np.abs(s.c_bulk)), 1e-30)
        atol_c = 1e-12 * cscale
        atol_g = 1e-12 * max(1.0, cscale * np.max(np.abs(s.q)) * s.x[-1] / s.eps) if s.use_migration else 1.0
        inv_dt = 0.0 if pure_newton else 1.0 / dt0
        n_newton = 0
        fnorm_old = None
        info = {'steps': 0, 'converged': False}
        for it in range(max_steps):
            F, L, Dg, U, E0 = self.residual(y, blocks=True)
            fn = self._fnorm(F, y, atol_c)
            if fnorm_old is not None and inv_dt > 0.0:
                ratio = fnorm_old / max(fn, 1e-300)
                inv_dt = inv_dt / min(growth, max(ratio, 0.1))
                if inv_dt < 1e-9:           # dt > 1e9 s: switch to pure Newton
                    inv_dt = 0.0
            fnorm_old = fn
            J = self.to_sparse(L, Dg, U, E0)
            A = sp.diags((M * inv_dt).reshape(-1)) - J
            delta = spla.splu(A.tocsc()).solve(F.reshape(-1)).reshape(y.shape)
            n_newton += 1
            w = self.weights(y, rtol, atol_c, atol_g)
            dn = np.max(np.abs(delta) / w)
            y = y + delta
            if verbose:
                print('it %3d  1/dt %.3e  |F| %.3e  |d|/w %.3e' % (it, inv_dt, fn, dn))
            if not np.all(np.isfinite(y)):
                info['failed'] = 'nan'
                break
            if inv_dt == 0.0 and dn < 1.0:
                info['converged'] = True
                break
        info['steps'] = n_newton
        return y, info

    def _fnorm(self, F, y, atol_c):
        S = self.S
        return np.max(np.abs(F[:-1, :S]) / (np.abs(y[:-1, :S]) + 1e-6 * np.max(np.abs(self.s.c_bulk))))

    # ------------------------------------------------------------------
    def unpack(self, y):
        """(C[S,n], v[n], g[n]) in the reference's output conventions
        (efield=-g, potential=v; calculator_old.py:816-818)."""
        s = self.s
        C = y[:, :self.S].T.copy()
        g = y[:, self.S].copy()
        if self.stern:
            v = y[:, self.S + 1].copy()
        elif s.use_migration:
            v, _, _ = s.poisson(C)
        else:
            v = np.zeros(self.n)
        return C, v, g
