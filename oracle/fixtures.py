"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

Builds oracle ``PnpSystem`` objects from the setup arrays stored in
tests/golden/ref_*.npz (written by tests/golden/make_golden.py from the
reference's own ``Transport``) or from plain arrays.
"""
import numpy as np

from .pnp_oracle import PnpSystem


def parse_rx(su):
    rx = []
    for e, p, kf, kr in zip(su['rx_educts'], su['rx_products'], su['rx_kf'], su['rx_kr']):
        ed = [int(a) for a in str(e).split(',') if a != '']
        pr = [int(a) for a in str(p).split(',') if a != '']
        rx.append((ed, pr, float(kf), float(kr)))
    return rx


def system_from_setup(su, rate_mode='summed', literal_sign=False, **over):
    """literal_sign=True reproduces the reference's wall stencil exactly
    (it subtracts flux_bound, calculator_old.py:902-909): J=-flux_bound[:,0].
    literal_sign=False is the product convention J=+species[sp]['flux']."""
    J = -su['flux_bound'][:, 0] if literal_sign else su['flux']
    kw = dict(z=su['z'], D=su['D'], c_bulk=su['c_bulk'], J=J, x=su['xmesh'], beta=float(su['beta']),
              eps=float(su['eps']), reactions=parse_rx(su), rate_mode=rate_mode,
              use_migration=bool(su['use_migration']), phi_wall=float(su['phi_wall']),
              g_bulk=float(su['g_bulk']), uniform=True)
    kw.update(over)
    return PnpSystem(**kw)


def system_from_batch(batch, c, rate_mode='summed'):
    """oracle PnpSystem of cell c of a catint_b200 CellBatch (uniform or table mesh, default or Stern
    Poisson boundary) -- the inverse of catint_b200.calculator.build_cell_batch's parameter records
    (include/catint_pnp.h: CATINT_PNP_P_*)."""
    S = batch.S
    p = batch.par[c]
    n = int(batch.nx[c])
    uniform = batch.mesh_id is None or int(batch.mesh_id[c]) < 0
    if uniform:
        x = np.arange(n) * p[3 * S + 5]
    else:
        x = p[3 * S + 5] * np.asarray(batch.mesh_xi[int(batch.mesh_id[c])][:n], dtype=float)
    stern = int(batch.poisson_bc) == 1
    return PnpSystem(z=batch.z, D=p[2 * S:3 * S], c_bulk=p[0:S], J=p[S:2 * S], x=x, beta=p[3 * S], eps=p[3 * S + 1],
                     reactions=batch.reactions, rate_mode=rate_mode, use_migration=batch.use_migration,
                     poisson_bc='stern_robin' if stern else 'dirichlet_wall_neumann_bulk',
                     phi_wall=p[3 * S + 2], g_bulk=p[3 * S + 3], phiM=p[3 * S + 2], phiPZC=0.0,
                     C_stern=p[3 * S + 4], uniform=uniform)
