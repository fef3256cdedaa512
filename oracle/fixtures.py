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
