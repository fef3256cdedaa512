"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU variable-order, variable-step BDF (orders 1..5, Nordsieck history array,
Newton corrector with the analytic block-tridiagonal Jacobian, sparse LU) on
the local (c,g[,phi]) form of oracle/pnp_local.py.  Oracle ladder step (ii)
(SURVEY 8c): validated against scipy ``odeint`` on 101-node grids
(tests/test_oracle.py), then used where dense ``odeint`` is impractical
(1001/5001 nodes, Stern extension) and as the algorithmic prototype of the
CUDA integrator (same step/order control, so both can be compared step count
for step count).

The step/order selection follows the published fixed-leading-coefficient-free
Nordsieck BDF scheme of the ODEPACK/CVODE family (Brown, Byrne, Hindmarsh,
"VODE: a variable-coefficient ODE solver", SIAM J. Sci. Stat. Comput. 10
(1989); Hindmarsh et al., "SUNDIALS", ACM TOMS 31 (2005)) -- the same family
``scipy.integrate.odeint`` (LSODA) belongs to, which the reference calls at
/root/reference/catint/calculator_old.py:947.  Error control uses the weighted
max-norm over the concentration unknowns with weights rtol*|c|+atol (LSODA's
``vmnorm``); the algebraic unknowns (g, phi) are not error controlled (they
are eliminated in the reference's ODE).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .pnp_local import LocalForm

QMAX = 5
ADDON = 1e-6
BIAS1, BIAS2, BIAS3 = 6.0, 6.0, 10.0
ETAMX1, ETAMX2, ETAMXF, ETAMIN, ETACF = 1e4, 10.0, 0.2, 0.1, 0.25
THRESH = 1.5
MXNCF, MXNEF, MXNEF1, SMALL_NEF, LONG_WAIT = 10, 7, 3, 2, 10
MAXCOR, CRDOWN, RDIV, NLSCOEF = 3, 0.3, 2.0, 0.1


class BdfStats(object):
    def __init__(self):
        self.nst = 0; self.nni = 0; self.ncfn = 0; self.netf = 0; self.nsolve = 0
        self.qhist = []


class BdfIntegrator(object):
    def __init__(self, lf: LocalForm, rtol=1.49012e-8, atol=1.49012e-8, fresh_jacobian=True):
        self.lf = lf
        self.S, self.n, self.b = lf.S, lf.n, lf.b
        self.rtol, self.atol = rtol, atol
        self.M = lf.mass_diag()
        self.cmask = self.M > 0            # error-controlled unknowns
        self.stats = BdfStats()
        self.fresh = fresh_jacobian
        self._lu = None
        self._gammap = 1.0
        self._nstlp = 0

    # -- norms -----------------------------------------------------------
    def ewt(self, y):
        return 1.0 / (self.rtol * np.abs(y) + self.atol)

    def norm(self, v, w):
        return float(np.max(np.abs(v[self.cmask]) * w[self.cmask]))

    # -- BDF coefficients (variable step) ----------------------------------
    def set_bdf(self):
        q, h, tau = self.q, self.h, self.tau
        l = np.zeros(QMAX + 2)
        l[0] = l[1] = 1.0
        xi_inv = xistar_inv = 1.0
        alpha0 = alpha0_hat = -1.0
        hsum = h
        if q > 1:
            for j in range(2, q):
                hsum += tau[j - 1]
                xi_inv = h / hsum
                alpha0 -= 1.0 / j
                for i in range(j, 0, -1):
                    l[i] += l[i - 1] * xi_inv
            alpha0 -= 1.0 / q
            xistar_inv = -l[1] - alpha0
            hsum += tau[q - 1]
            xi_inv = h / hsum
            alpha0_hat = -l[1] - xi_inv
            for i in range(q, 0, -1):
                l[i] += l[i - 1] * xistar_inv
        tq = np.zeros(6)
        A1 = 1.0 - alpha0_hat + alpha0
        A2 = 1.0 + q * A1
        tq[2] = abs(A1 / (alpha0 * A2))
        tq[5] = abs(A2 * xistar_inv / (l[q] * xi_inv))
        if self.qwait == 1:
            if q > 1:
                C = xistar_inv / l[q]
                A3 = alpha0 + 1.0 / q
                A4 = alpha0_hat + xi_inv
                tq[1] = abs(C * (1.0 - A4 + A3) / A3)
            else:
                tq[1] = 1.0
            hsum += tau[q]
            xi_inv = h / hsum
            A5 = alpha0 - 1.0 / (q + 1)
            A6 = alpha0_hat - xi_inv
            tq[3] = abs(((1.0 - A6 + A5) / A2) / (xi_inv * (q + 2) * A5))
        tq[4] = NLSCOEF / tq[2]
        self.l, self.tq = l, tq

    def rescale(self, eta):
        f = eta
        for j in range(1, self.q + 1):
            self.zn[j] *= f
            f *= eta
        self.h = self.hscale * eta
        self.hscale = self.h

    def predict(self):
        for k in range(1, self.q + 1):
            for j in range(self.q, k - 1, -1):
                self.zn[j - 1] += self.zn[j]

    def restore(self):
        for k in range(1, self.q + 1):
            for j in range(self.q, k - 1, -1):
                self.zn[j - 1] -= self.zn[j]

    def increase_order(self):
        q, tau, hscale = self.q, self.tau, self.hscale
        l = np.zeros(QMAX + 2)
        l[2] = alpha1 = prod = xiold = 1.0
        alpha0 = -1.0
        hsum = hscale
        if q > 1:
            for j in range(1, q):
                hsum += tau[j + 1]
                xi = hsum / hscale
                prod *= xi
                alpha0 -= 1.0 / (j + 1)
                alpha1 += 1.0 / xi
                for i in range(j + 2, 1, -1):
                    l[i] = l[i] * xiold + l[i - 1]
                xiold = xi
        A1 = (-alpha0 - alpha1) / prod
        self.zn[q + 1] = A1 * self.zn[QMAX]
        for j in range(2, q + 1):
            self.zn[j] += l[j] * self.zn[q + 1]

    def decrease_order(self):
        q, tau, hscale = self.q, self.tau, self.hscale
        l = np.zeros(QMAX + 2)
        l[2] = 1.0
        hsum = 0.0
        for j in range(1, q - 1):
            hsum += tau[j]
            xi = hsum / hscale
            for i in range(j + 2, 1, -1):
                l[i] = l[i] * xi + l[i - 1]
        for j in range(2, q):
            self.zn[j] -= l[j] * self.zn[q]

    # -- the nonlinear corrector ------------------------------------------
    def _setup(self, y, gamma):
        """assemble the Jacobian at y and factor  M/gamma - J  (CVODE's lsetup)."""
        lf = self.lf
        F, L, Dg, U, E0 = lf.residual(y, blocks=True)
        J = lf.to_sparse(L, Dg, U, E0)
        A = sp.diags((self.M / gamma).reshape(-1)) - J
        self._lu = spla.splu(A.tocsc())
        self._gammap = gamma
        self._nstlp = self.stats.nst
        self.stats.nsetups = getattr(self.stats, 'nsetups', 0) + 1
        return F

    def newton(self, force_setup=False):
        """solve  M*(rl1*zn1 + (y-zn0)) - gamma*F(y) = 0 ; returns (ok, acor, acnrm).
        fresh_jacobian=True: true Newton (setup in every iteration);
        False: modified Newton with the VODE/CVODE re-use policy (setup on the first step, after
        20 steps, when gamma drifted by more than 30 %, or after a failure with a stale matrix)."""
        lf = self.lf
        y0 = self.zn[0]
        gamma = self.h / self.l[1]
        rl1 = 1.0 / self.l[1]
        w = self.ewt_vec
        MSBP, DGMAX = 20, 0.3
        call_setup = (self.fresh or force_setup or self._lu is None or self.stats.nst >= self._nstlp + MSBP
                      or abs(gamma / self._gammap - 1.0) > DGMAX)
        while True:
            acor = np.zeros_like(y0)
            y = y0.copy()
            crate = 1.0
            delp = 0.0
            jcur = False
            F = None
            if call_setup:
                F = self._setup(y, gamma)
                jcur = True
            gamrat = gamma / self._gammap
            result = None
            for m in range(MAXCOR):
                if self.fresh and m > 0:
                    F = self._setup(y, gamma)
                elif F is None:
                    F = lf.residual(y)
                rhs = F - self.M * (rl1 * self.zn[1] + acor) / gamma
                d = self._lu.solve(rhs.reshape(-1)).reshape(y.shape)
                if gamrat != 1.0:
                    d *= 2.0 * gamrat / (1.0 + gamrat)
                self.stats.nsolve += 1
                self.stats.nni += 1
                dl = self.norm(d, w)
                acor += d
                y = y0 + acor
                F = None
                if m > 0:
                    crate = max(CRDOWN * crate, dl / delp)
                dcon = dl * min(1.0, crate) / self.tq[4]
                if dcon <= 1.0:
                    result = (True, acor, (dl if m == 0 else self.norm(acor, w)))
                    break
                if m + 1 == MAXCOR or (m >= 1 and dl > RDIV * delp):
                    break
                delp = dl
            if result is not None:
                return result
            if jcur or self.fresh:
                return False, acor, 0.0
            call_setup = True            # stale matrix: refresh and try once more

    # -- driver ------------------------------------------------------------
    def integrate(self, t_out, y0=None, h0=None, max_steps=200000, callback=None):
        """returns list of y at the requested times (interpolated from the
        Nordsieck array like ODEPACK's intdy)."""
        lf = self.lf
        y = lf.y_bulk() if y0 is None else y0.copy()
        t_out = list(t_out)
        self.q = 1
        self.qwait = 2
        self.tau = np.zeros(QMAX + 2)
        self.zn = np.zeros((QMAX + 2,) + y.shape)
        self.zn[0] = y
        F = lf.residual(y)
        f0 = F * self.M
        self.ewt_vec = self.ewt(y)
        if h0 is None:
            fn = self.norm(f0, self.ewt_vec)
            h0 = 1.0 / max(fn, 1e-300) if fn > 0 else 1e-6
            h0 = min(h0, 1e-3 * (t_out[-1] if t_out[-1] > 0 else 1.0))
        self.h = self.hscale = h0
        self.zn[1] = self.h * f0
        self.t = 0.0
        etamax = ETAMX1
        self.saved_tq5 = 0.0
        outs = []
        iout = 0
        st = self.stats
        while iout < len(t_out):
            if st.nst >= max_steps:
                raise RuntimeError('bdf: too many steps')
            # ---- one step -------------------------------------------
            ncf = nef = 0
            saved_t = self.t
            while True:
                if saved_t + self.h == saved_t:
                    raise RuntimeError('bdf: step size underflow at t=%g (h=%g): the solution blows up' % (saved_t, self.h))
                self.predict()
                self.t = saved_t + self.h
                self.set_bdf()
                ok, acor, acnrm = self.newton(force_setup=(ncf > 0 or nef > 0))
                if not ok:
                    st.ncfn += 1
                    ncf += 1
                    etamax = 1.0
                    self.restore(); self.t = saved_t
                    if ncf == MXNCF:
                        raise RuntimeError('bdf: repeated corrector failures at t=%g h=%g' % (self.t, self.h))
                    self.rescale(ETACF)
                    continue
                dsm = acnrm * self.tq[2]
                if dsm <= 1.0:
                    break
                st.netf += 1
                nef += 1
                self.restore(); self.t = saved_t
                etamax = 1.0
                if nef == MXNEF:
                    raise RuntimeError('bdf: repeated error test failures at t=%g' % self.t)
                if nef <= MXNEF1:
                    eta = 1.0 / ((BIAS2 * dsm) ** (1.0 / (self.q + 1)) + ADDON)
                    eta = max(ETAMIN, eta)
                    if nef >= SMALL_NEF:
                        eta = min(eta, ETAMXF)
                    self.rescale(eta)
                elif self.q > 1:
                    self.decrease_order()
                    self.q -= 1
                    self.qwait = self.q + 1
                    self.rescale(ETAMIN)
                else:
                    self.h *= ETAMIN
                    self.hscale = self.h
                    self.qwait = LONG_WAIT
                    self.zn[1] = self.h * (lf.residual(self.zn[0]) * self.M)
            # ---- complete step ----------------------------------------
            st.nst += 1
            st.qhist.append(self.q)
            q = self.q
            for i in range(q, 1, -1):
                self.tau[i] = self.tau[i - 1]
            if q == 1 and st.nst > 1:
                self.tau[2] = self.tau[1]
            self.tau[1] = self.h
            for j in range(q + 1):
                self.zn[j] += self.l[j] * acor
            # algebraic unknowns are solved exactly each step; keep their history consistent
            self.qwait -= 1
            if self.qwait == 1 and q != QMAX:
                self.zn[QMAX] = acor.copy()
                self.saved_tq5 = self.tq[5]
            self.ewt_vec = self.ewt(self.zn[0])
            # ---- next order / step size -------------------------------
            qprime = q
            if etamax == 1.0:
                self.qwait = max(self.qwait, 2)
                eta = 1.0
            else:
                etaq = 1.0 / ((BIAS2 * dsm) ** (1.0 / (q + 1)) + ADDON)
                if self.qwait != 0:
                    eta = etaq
                else:
                    self.qwait = 2
                    etaqm1 = 0.0
                    if q > 1:
                        ddn = self.norm(self.zn[q], self.ewt_vec) * self.tq[1]
                        etaqm1 = 1.0 / ((BIAS1 * ddn) ** (1.0 / q) + ADDON)
                    etaqp1 = 0.0
                    if q != QMAX and self.saved_tq5 != 0.0:
                        cquot = (self.tq[5] / self.saved_tq5) * (self.h / self.tau[2]) ** (q + 1)
                        dup = self.norm(acor - cquot * self.zn[QMAX], self.ewt_vec) * self.tq[3]
                        etaqp1 = 1.0 / ((BIAS3 * dup) ** (1.0 / (q + 2)) + ADDON)
                    etam = max(etaqm1, etaq, etaqp1)
                    if etam < THRESH:
                        eta = 1.0
                    elif etam == etaq:
                        eta = etaq
                    elif etam == etaqm1:
                        eta = etaqm1; qprime = q - 1
                    else:
                        eta = etaqp1; qprime = q + 1
                        self.zn[QMAX] = acor.copy()
                if eta < THRESH:
                    eta = 1.0
                else:
                    eta = min(eta, etamax)
            etamax = ETAMX2
            if callback is not None:
                callback(self)
            # ---- output (dense output through the Nordsieck polynomial) ----
            while iout < len(t_out) and self.t >= t_out[iout] * (1 - 1e-14):
                s = (t_out[iout] - self.t) / self.h
                yo = self.zn[q].copy()
                for j in range(q - 1, -1, -1):
                    yo = yo * s + self.zn[j]
                outs.append(yo)
                iout += 1
            # ---- apply order change and rescale for the next step -----------
            if qprime != q:
                if qprime > q:
                    self.increase_order()
                else:
                    self.decrease_order()
                self.q = qprime
                self.qwait = self.q + 1
            if eta != 1.0:
                self.rescale(eta)
        return outs
