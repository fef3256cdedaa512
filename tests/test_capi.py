"""The C-ABI library loads and exports every symbol include/catint_pnp.h declares; host-only
entry points behave; compute entry points fail loudly (no CPU fallback) when no B200 is there."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, batch_from_setup
from catint_b200 import backend as be

HEADER = os.path.join(ROOT, 'include', 'catint_pnp.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(catint_pnp_[a-z_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    return be.load_library()


def test_header_symbols_are_exported(lib):
    names = declared_functions()
    assert set(names) == set(be.EXPORTS)
    for name in names:
        assert hasattr(lib, name), name


def test_header_cites_the_reference():
    src = open(HEADER).read()
    for cite in ('calculator_old.py:827-935', 'calculator_old.py:680-819', 'calculator_old.py:159-208',
                 'calculator_old.py:946-973', 'calculator.py:204-226'):
        assert cite in src


def test_version_and_struct_layout(lib):
    assert lib.catint_pnp_version() >= 100
    # the ctypes mirrors must have the C layout: 6 int32 + 14 int32 + 2*12*4 int32 + 2*12 + 14*12 doubles
    # + the host pointer to the flux equations
    assert ctypes.sizeof(be.CatintPnpShared) == 4 * (6 + 14 + 96) + 8 * (24 + 168) + ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(be.CatintPnpCells) == 6 * ctypes.sizeof(ctypes.c_void_p)
    # flux equations: 2 + 4 int32, 4 x 96 int32 code words, 4 x 32 constants, 14 x 4 coefficients
    assert ctypes.sizeof(be.CatintPnpFluxEq) == 4 * (2 + 4 + 4 * 96) + 8 * (4 * 32 + 14 * 4)
    assert ctypes.sizeof(be.CatintPnpControl) == 16 + 32 + ctypes.sizeof(ctypes.c_void_p)


def test_workspace_query_is_host_only(lib):
    batch = batch_from_setup(load_golden('ref_c1.npz'), B=3)
    sh = batch.shared_struct()
    one = lib.catint_pnp_workspace_bytes(ctypes.byref(sh), 1)
    three = lib.catint_pnp_workspace_bytes(ctypes.byref(sh), 3)
    nb, n = 9, 101
    per_cell = (6 * n * nb + n * nb + n * nb * nb + nb * nb + 3 * n * nb) * 8
    assert one >= per_cell and three - one == 2 * (one - be.MAX_OUTPUT_TIMES * 8)
    assert lib.catint_pnp_workspace_bytes(None, 3) == 0


def test_shared_struct_tables():
    su = load_golden('ref_c1.npz')
    batch = batch_from_setup(su, B=2)
    sh = batch.shared_struct()
    assert (sh.S, sh.R, sh.nx_max, sh.use_migration, sh.poisson_bc) == (8, 5, 101, 1, 0)
    assert list(sh.z)[:8] == [1, 0, -1, 0, 0, -1, -2, 1]
    assert list(sh.educt[0]) == [1, 2, -1, -1] and list(sh.product[0]) == [5, -1, -1, -1]      # CO2 + OH- <-> HCO3-
    assert list(sh.educt[2]) == [-1, -1, -1, -1] and list(sh.product[2])[:2] == [2, 7]          # H2O <-> OH- + H+
    assert sh.nu[2][0] == -1.0 and sh.nu[2][1] == -1.0 and sh.nu[2][2] == 1.0                  # OH-
    legacy = batch_from_setup(su, B=1, rate_mode='legacy_overwrite').shared_struct()
    assert sum(1 for r in range(5) if legacy.nu[2][r] != 0.0) == 1


@pytest.mark.skipif(os.environ.get('CUDA_VISIBLE_DEVICES', None) not in (None, '') and False, reason='')
def test_compute_entry_points_fail_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert lib.catint_pnp_device_count() == 0
    batch = batch_from_setup(load_golden('ref_c1.npz'), B=1)
    sh = batch.shared_struct()
    cells = be.CatintPnpCells()
    par = np.ascontiguousarray(batch.par)
    nx = np.ascontiguousarray(batch.nx)
    cells.par = par.ctypes.data
    cells.nx = nx.ctypes.data
    dummy = np.zeros(8)
    rc = lib.catint_pnp_rhs_batch(ctypes.byref(sh), ctypes.byref(cells), 1, dummy.ctypes.data, dummy.ctypes.data,
                                  None, None, None)
    assert rc == -4 and b'sm_100' in lib.catint_pnp_last_error()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        be.PnpBackend()


def test_argument_errors_are_codes_not_exits(lib):
    assert lib.catint_pnp_rhs_batch(None, None, 1, None, None, None, None, None) == -1
    assert b'NULL' in lib.catint_pnp_last_error()
