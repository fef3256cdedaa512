import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault('CATINT_QUIET', '1')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


REF_CASES = ['c1', 'c1_nomig', 'c1_norx', 'c1_pH7p5', 'c1_L30', 'c1_T320', 'c4']
RHS_CASES = ['c1', 'c1_nomig', 'c1_norx', 'c1_L30', 'c4']


def case_inputs(name):
    """Transport keyword dicts that reproduce the reference-generated fixture `name`."""
    from catint_b200 import workloads as w
    return {
        'c1': lambda: w.co2r_inputs(),
        'c1_nomig': lambda: w.co2r_inputs(migration=False),
        'c1_norx': lambda: w.co2r_inputs(reactions=False),
        'c1_pH7p5': lambda: w.co2r_inputs(pH=7.5, i_CO=-30., i_H2=-12.),
        'c1_L30': lambda: w.co2r_inputs(L=30e-6),
        'c1_T320': lambda: w.co2r_inputs(temperature=320.),
        'c4': lambda: w.co2r_inputs(extra_species=True),
    }[name]()


def batch_from_setup(su, B=1, rate_mode='summed', literal_sign=False, fluxes=None):
    """CellBatch straight from a reference setup fixture (no Transport involved)."""
    from catint_b200 import backend as be
    from oracle.fixtures import parse_rx
    S = len(su['z'])
    rx = parse_rx(su)
    nu = be.stoichiometry(S, rx, rate_mode)
    par = np.zeros((B, be.npar(S)))
    J = -su['flux_bound'][:, 0] if literal_sign else su['flux']
    for c in range(B):
        par[c, 0:S] = su['c_bulk']
        par[c, S:2 * S] = J if fluxes is None else fluxes[c]
        par[c, 2 * S:3 * S] = su['D']
        par[c, 3 * S + 0] = su['beta']
        par[c, 3 * S + 1] = su['eps']
        par[c, 3 * S + 2] = su['phi_wall']
        par[c, 3 * S + 3] = su['g_bulk']
        par[c, 3 * S + 4] = 0.2
        par[c, 3 * S + 5] = su['dx']
    nx = np.full(B, int(su['nx']), dtype=np.int32)
    return be.CellBatch(su['z'], rx, nu, par, nx, use_migration=bool(su['use_migration']),
                        species=[str(s) for s in su['species']])


def oracle_system_of_cell(batch, c, rate_mode='summed'):
    """oracle PnpSystem for cell c of a (uniform mesh) CellBatch."""
    from oracle.pnp_oracle import PnpSystem
    S = batch.S
    p = batch.par[c]
    n = int(batch.nx[c])
    x = np.arange(n) * p[3 * S + 5]
    return PnpSystem(z=batch.z, D=p[2 * S:3 * S], c_bulk=p[0:S], J=p[S:2 * S], x=x, beta=p[3 * S],
                     eps=p[3 * S + 1], reactions=batch.reactions, rate_mode=rate_mode,
                     use_migration=batch.use_migration, phi_wall=p[3 * S + 2], g_bulk=p[3 * S + 3], uniform=True)


@pytest.fixture
def resultsdir(tmp_path):
    return str(tmp_path)
