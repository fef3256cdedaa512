"""Quick GPU sanity script (development aid): K1/K2/K3 against the oracle on the C1 fixture."""
import os, sys, time
import numpy as np
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from catint_b200 import backend as be
from oracle.fixtures import system_from_setup, parse_rx
from oracle.pnp_local import LocalForm

def batch_from_setup(su, B=1, rate_mode='summed', fluxes=None):
    S = len(su['z'])
    rx = parse_rx(su)
    nu = be.stoichiometry(S, rx, rate_mode)
    par = np.zeros((B, be.npar(S)))
    for c in range(B):
        par[c, 0:S] = su['c_bulk']
        par[c, S:2 * S] = su['flux'] if fluxes is None else fluxes[c]
        par[c, 2 * S:3 * S] = su['D']
        par[c, 3 * S + 0] = su['beta']; par[c, 3 * S + 1] = su['eps']
        par[c, 3 * S + 2] = su['phi_wall']; par[c, 3 * S + 3] = su['g_bulk']
        par[c, 3 * S + 4] = 0.2; par[c, 3 * S + 5] = su['dx']
    nx = np.full(B, int(su['nx']), dtype=np.int32)
    return be.CellBatch(su['z'], rx, nu, par, nx, use_migration=bool(su['use_migration']))

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'c1'
    su = dict(np.load(os.path.join(os.path.dirname(HERE), 'golden', 'ref_%s.npz' % name)))
    S, n = len(su['z']), int(su['nx'])
    bk = be.PnpBackend('cuda:0')
    batch = batch_from_setup(su, B=4)
    db = bk.upload(batch)
    sys_ = system_from_setup(su, 'summed', False)
    lf = LocalForm(sys_)
    # ---- K1
    states = su['rhs_states'] if 'rhs_states' in su else su['c0'][None]
    worst = 0
    for st in states:
        C = st.reshape(S, n)
        c = torch.tensor(np.ascontiguousarray(np.broadcast_to(C.T[None], (4, n, S))), device='cuda:0')
        dcdt, g, phi = bk.rhs(db, c)
        ref, v, gg, lapl = sys_.rhs(st, with_field=True)
        ref = ref.reshape(S, n).T
        got = dcdt[0].cpu().numpy()
        e = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        eg = np.max(np.abs(g[0].cpu().numpy() - gg)) / max(np.max(np.abs(gg)), 1e-300)
        ev = np.max(np.abs(phi[0].cpu().numpy() - v)) / max(np.max(np.abs(v)), 1e-300)
        worst = max(worst, e)
        print('K1 rhs rel-to-max err %.2e  g %.2e  phi %.2e' % (e, eg, ev))
    # ---- K2
    C = states[min(2, len(states) - 1)].reshape(S, n)
    y = lf.y_from_c(C)
    F, L, Dg, U, E0 = lf.residual(y, blocks=True)
    yt = torch.tensor(np.ascontiguousarray(np.broadcast_to(y[None], (4,) + y.shape)), device='cuda:0')
    Fg, Lg, Dgg, Ug = [a[1].cpu().numpy() for a in bk.jacobian(db, yt)]
    def rel(a, b):
        return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
    print('K2 F %.2e  D %.2e  U %.2e  L[1:] %.2e  E0 %.2e' % (rel(Fg, F), rel(Dgg, Dg), rel(Ug, U), rel(Lg[1:], L[1:]), rel(Lg[0], E0)))
    # ---- K3
    torch.cuda.synchronize()
    for mode in (be.MODE_STEADY, be.MODE_TRANSIENT):
        t0 = time.time()
        out = bk.solve(db, [10.0, 200.0], mode=mode)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print('K3 mode %d: %.3fs status %s steps %s newton %s' % (mode, dt, out['status'].tolist(), out['n_steps'].tolist(), out['n_newton'].tolist()))
        cfin = out['c'][-1, 0].cpu().numpy()
        print('   surface c:', cfin[0])
        gp = os.path.join(os.path.dirname(HERE), 'golden', 'oracle_%s_summed.npz' % name)
        if os.path.exists(gp):
            go = dict(np.load(gp))
            key = 'newton_c' if mode == be.MODE_STEADY else 'c_end'
            ref = go[key].reshape(S, n).T
            cs = np.max(np.abs(su['c_bulk']))
            print('   vs oracle %s: max rel %.3e' % (key, np.max(np.abs(cfin - ref) / (np.abs(ref) + 1e-12 * cs))))
            ref10 = go['c_t10'].reshape(S, n).T
            c10 = out['c'][0, 0].cpu().numpy()
            print('   t=10 vs odeint: max rel %.3e' % np.max(np.abs(c10 - ref10) / (np.abs(ref10) + 1e-9 * cs)))
            print('   flux_out', out['flux'][0].cpu().numpy(), 'J', su['flux'])

if __name__ == '__main__':
    main()
