"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Restatement of the reference's Poisson routine and of its two fixed-step steppers
(/root/reference/catint/calculator_old.py):

    get_potential_and_gradient   :680-819  (every pb_bound pair of :776-803; scale_pb_grid=None)
    integrate_FTCS               :976-1029
    integrate_Crank_Nicolson     :457-564

Deliberately literal: the same loops, the same index quirks (Crank-Nicolson indexes grad_v / lapl_v by the interior
index, multiplies with np.dot(C, B1), has no reaction term; FTCS takes the rates before the wall update), dense
np.linalg.solve where the reference has it.  "vzeta" -- the reference reads system['vzeta'], a key its Transport
rejects, so upstream these steppers cannot run -- is taken as the wall potential.

PARITY UNPINNED for the steppers: the reference file is Python 2 and cannot execute these loops here (and would
stop at the missing key), so there is no reference-generated golden; the restatement is line by line and the
Poisson routine is additionally pinned, for the default pair, by the reference-executed ode_func fixtures through
oracle/pnp_oracle.py (tests/test_explicit.py).
"""
import numpy as np
from scipy.sparse import diags

UNIT_F = 96485.33289


def potential_and_gradient(C, dx, charges, eps, pb_bound):
    """C [S,nx]; charges = z*F; pb_bound = {'potential': {'wall','bulk'}, 'gradient': {'wall','bulk'}} -> v, grad_v, lapl_v"""
    nx = C.shape[1]
    bounds = pb_bound
    if bounds['gradient']['wall'] is not None and bounds['gradient']['bulk'] is not None:
        raise ValueError('Cannot use two boundary conditions for gradient')                    # :703-705

    def solve_poisson(q, sol0):                                                                # :707-722
        A = diags([1., -2., 1.], [-1, 0, 1], shape=(nx - 2, nx - 2)).toarray()
        b = np.array(q[1:nx - 1]) * dx ** 2
        b[0] -= sol0[0]
        b[-1] -= sol0[-1]
        x = np.linalg.solve(A, b)
        return np.array([sol0[0]] + list(x) + [sol0[-1]])

    def integrate_1d(var0, integrand, inv=False):                                              # :751-759
        var = np.array(var0, dtype=float)
        it = reversed(range(1, nx - 1)) if inv else range(1, nx - 1)
        for i in it:
            if inv:
                var[i] = var[i + 1] - integrand[i] * dx
            else:
                var[i] = var[i - 1] + integrand[i] * dx
        return var

    rhs = np.zeros([nx])
    for i in range(nx):                                                                        # :762-766
        for k in range(C.shape[0]):
            rhs[i] -= charges[k] * C[k, i] / eps
    lapl_v = rhs
    v = np.zeros([nx])
    grad_v = np.zeros([nx])
    if bounds['potential']['wall'] is not None:
        v[0] = bounds['potential']['wall']
    if bounds['potential']['bulk'] is not None:
        v[-1] = bounds['potential']['bulk']
    if bounds['potential']['wall'] is not None and bounds['potential']['bulk'] is not None:    # :775-780
        v = solve_poisson(rhs, v)
        for i in range(1, nx - 1):
            grad_v[i] = 1. / (2 * dx) * (v[i + 1] - v[i - 1])
        grad_v[0] = grad_v[1] + (grad_v[1] - grad_v[2])
        grad_v[-1] = grad_v[-2] + (grad_v[-2] - grad_v[-3])
    else:
        if bounds['gradient']['wall'] is not None:                                             # :783-786
            grad_v[0] = bounds['gradient']['wall']
            grad_v = integrate_1d(grad_v, rhs)
            grad_v[-1] = grad_v[-2] + (grad_v[-2] - grad_v[-3])
        if bounds['gradient']['bulk'] is not None:                                             # :787-790
            grad_v[-1] = bounds['gradient']['bulk']
            grad_v = integrate_1d(grad_v, rhs, inv=True)
            grad_v[0] = grad_v[1] + (grad_v[1] - grad_v[2])
        if bounds['potential']['wall'] is not None:                                            # :792-794
            v = integrate_1d(v, grad_v)
            v[-1] = v[-2] + (v[-2] - v[-3])
        if bounds['potential']['bulk'] is not None:                                            # :795-797
            v = integrate_1d(v, grad_v, inv=True)
            v[0] = v[1] + (v[1] - v[2])
    return v, grad_v, lapl_v


class ExplicitModel(object):
    """the attributes of the reference's Transport the steppers read"""

    def __init__(self, z, D, c_bulk, flux, dx, nx, beta, eps, pb_bound, reactions=(), nu=None, use_migration=True):
        self.z = np.asarray(z, dtype=float)
        self.charges = self.z * UNIT_F
        self.D = np.asarray(D, dtype=float)
        self.c_bulk = np.asarray(c_bulk, dtype=float)
        self.flux = np.asarray(flux, dtype=float)          # flux_bound[:,0] as the steppers use it
        self.dx, self.nx, self.beta, self.eps = float(dx), int(nx), float(beta), float(eps)
        self.mu = self.D * self.charges * self.beta        # transport.py:436
        self.pb_bound = pb_bound
        self.vzeta = pb_bound['potential']['wall'] if pb_bound['potential']['wall'] is not None else 0.0
        self.reactions = list(reactions)
        self.nu = np.zeros((len(self.z), 0)) if nu is None else np.asarray(nu, dtype=float)
        self.use_migration = bool(use_migration)
        self.S = len(self.z)

    def c0(self):
        return np.repeat(self.c_bulk[:, None], self.nx, axis=1)

    def get_rates(self, C):
        rates = np.zeros([self.S, C.shape[1]])
        for r, (ed, pr, kf, kr) in enumerate(self.reactions):
            net = kf * np.prod(C[ed, :], axis=0) - kr * np.prod(C[pr, :], axis=0)
            rates += self.nu[:, r:r + 1] * net[None, :]
        return rates

    def field(self, C):
        return potential_and_gradient(C, self.dx, self.charges, self.eps, self.pb_bound)


def integrate_ftcs(m, dt, nt, itout, lax_friedrich=False, C0=None):
    """integrate_FTCS (:976-1029) -> list of (C[S,nx], v, grad_v) after the steps in itout"""
    nx, dx = m.nx, m.dx
    C = m.c0() if C0 is None else np.array(C0, dtype=float)
    Cb = C[:, -1].copy()                                   # C0[(k+1)*nx-1]
    out = []
    for n in range(0, nt):
        if m.use_migration:
            v, grad_v, lapl_v = m.field(C)
        else:
            v = np.zeros([nx]); grad_v = np.zeros([nx]); lapl_v = np.zeros([nx])
        rates = m.get_rates(C)
        for k in range(m.S):
            flux = m.flux[k]
            divisor = 2 * m.D[k] - m.mu[k] * (v[1] - m.vzeta)
            C[k, 0] = ((2 * m.D[k] + m.mu[k] * (v[1] - m.vzeta)) * C[k, 1] + flux * 2. * dx) / divisor
            C[k, -1] = Cb[k]
            temp = np.zeros([nx])
            temp[0] = C[k, 0]
            temp[-1] = C[k, -1]
            for i in range(1, nx - 1):
                W = m.D[k] * dt / dx ** 2 - dt / (2. * dx) * m.mu[k] * grad_v[i + 1] + 0.5
                M = -2. * m.D[k] * dt / dx ** 2
                E = m.D[k] * dt / dx ** 2 + dt / (2. * dx) * m.mu[k] * grad_v[i - 1] + 0.5
                if not lax_friedrich:
                    W -= 0.5
                    E -= 0.5
                    M += 1
                temp[i] = E * C[k, i - 1] + M * C[k, i] + W * C[k, i + 1] + rates[k, i] * dt
            C[k, :] = temp
        if n in itout:
            out.append((C.copy(), np.array(v), np.array(grad_v)))
    return out


def integrate_crank_nicolson(m, dt, nt, itout, lax_friedrich=False, C0=None):
    """integrate_Crank_Nicolson (:457-564)"""
    nx, dx = m.nx, m.dx

    def a_matrix(s):
        return diags([-0.5 * s, 1 + s, -0.5 * s], [-1, 0, 1], shape=(nx - 2, nx - 2)).toarray()

    def b1_matrix(s):
        return diags([0.5 * s, 1 - s, 0.5 * s], [-1, 0, 1], shape=(nx - 2, nx - 2)).toarray()

    def add_field(B1, A, grad_v, lapl_v, ee):                                                   # :476-494
        for i in range(nx - 2):
            if i == 0:
                jvalues = [i, i + 1]
            elif i == nx - 3:
                jvalues = [i - 1, i]
            else:
                jvalues = [i - 1, i, i + 1]
            for j in jvalues:
                if i == j:
                    B1[i, j] += ee * lapl_v[i]
                if abs(i - j) == 1:
                    if j < i:
                        B1[i, j] -= ee * grad_v[i] / 4. / dx
                        A[i, j] += ee * grad_v[i] / 4. / dx
                    elif i < j:
                        B1[i, j] += ee * grad_v[i] / 4. / dx
                        A[i, j] -= ee * grad_v[i] / 4. / dx
        return B1, A

    C = m.c0() if C0 is None else np.array(C0, dtype=float)
    Cb = C[:, -1].copy()
    COLD = np.zeros([m.S, nx])
    v = np.zeros([nx]); grad_v = np.zeros([nx]); lapl_v = np.zeros([nx])
    out = []
    for n in range(1, nt):
        if m.use_migration:
            v, grad_v, lapl_v = m.field(C)
        for k in range(m.S):
            if n == 1:
                COLD[k, :] = C[k, :].copy()
            C[k, 0] = (-2 * m.D[k] - m.mu[k] * (v[1] - m.vzeta)) / (-2 * m.D[k] + m.mu[k] * (v[1] - m.vzeta)) * C[k, 1] \
                - 2 * m.flux[k] * dx / (-2 * m.D[k] + m.mu[k] * (v[1] - m.vzeta))
            C[k, -1] = Cb[k]
            s = m.D[k] * dt / dx ** 2
            if lax_friedrich:
                s += 0.5
            ee = m.charges[k] * m.beta * dt * m.D[k] if m.use_migration else 0.0
            A = a_matrix(s)
            B1 = b1_matrix(s)
            if m.use_migration:
                B1, A = add_field(B1, A, grad_v, lapl_v, ee)
            B = np.dot(C[k, 1:-1], B1)
            B[0] += (0.5 * s + ee * grad_v[0] / 4. / dx) * (C[k, 0] + COLD[k, 0])               # :496-503
            B[-1] += (0.5 * s - ee * grad_v[-1] / 4. / dx) * (C[k, -1] + COLD[k, -1])
            C[k, 1:-1] = np.linalg.solve(A, B)
            COLD[k, :] = C[k, :]
        if n in itout:
            out.append((C.copy(), np.array(v), np.array(grad_v)))
    return out
