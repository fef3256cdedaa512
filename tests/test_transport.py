"""Host-side mirror of catint.transport.Transport against fixtures written by the reference's
own Transport (tests/golden/ref_*.npz) plus the reference's documented error behaviour."""
import copy
import os

import numpy as np
import pytest

from conftest import load_golden, REF_CASES, case_inputs
from catint_b200.transport import Transport, parse_reaction, charge_from_symbol, reaction_species
from catint_b200 import workloads


@pytest.mark.parametrize('name', REF_CASES)
def test_setup_parity_with_reference_transport(name, resultsdir):
    su = load_golden('ref_%s.npz' % name)
    tp = Transport(resultsdir=resultsdir, **case_inputs(name))
    names = list(tp.species)
    assert names == [str(s) for s in su['species']]          # species order IS the species index
    cb = np.array([tp.species[s]['bulk_concentration'] for s in names])
    fl = np.array([tp.species[s]['flux'] for s in names])
    # integer/bit-level quantities
    assert np.array_equal(np.array([tp.species[s]['charge'] for s in names]), su['z'])
    assert tp.nx == int(su['nx']) and tp.nspecies == len(names)
    for got, want in [(cb, su['c_bulk']), (fl, su['flux']), (tp.D, su['D']), (tp.charges, su['charges']),
                      (tp.c0, su['c0']), (tp.xmesh, su['xmesh']), (tp.flux_bound, su['flux_bound'])]:
        assert got.shape == want.shape
        assert np.array_equal(got, want)
    assert tp.beta == float(su['beta']) and tp.eps == float(su['eps']) and tp.dx == float(su['dx'])
    assert tp.debye_length == float(su['debye_length'])
    assert tp.ionic_strength == float(su['ionic_strength'])
    assert tp.system['bulk_pH'] == float(su['bulk_pH'])
    assert tp.pb_bound['potential']['wall'] == float(su['phi_wall'])
    assert tp.pb_bound['gradient']['bulk'] == float(su['g_bulk'])
    # parsed reaction table in library order
    if len(su['rx_names']):
        rx = [r for r in tp.electrolyte_reactions if 'rates' in tp.electrolyte_reactions[r]]
        assert rx == [str(r) for r in su['rx_names']]
        for r, e, p, kf, kr in zip(rx, su['rx_educts'], su['rx_products'], su['rx_kf'], su['rx_kr']):
            ent = tp.electrolyte_reactions[r]
            ed = ','.join(str(names.index(s)) for s in ent['reaction'][0] if s in names)
            pr = ','.join(str(names.index(s)) for s in ent['reaction'][1] if s in names)
            assert (ed, pr) == (str(e), str(p))
            assert ent['rates'] == [float(kf), float(kr)]


def test_ragged_node_count(resultsdir):
    """nx counts intervals; np.arange(0,xmax+dx,dx) gives 101 or 102 nodes (SURVEY C-6)."""
    assert Transport(resultsdir=resultsdir, **workloads.co2r_inputs(L=50e-6)).nx == 101
    tp = Transport(resultsdir=resultsdir, **workloads.co2r_inputs(L=30e-6))
    assert tp.nx == 102 and tp.xmesh[-1] > 30e-6


def test_reentrant_and_inputs_not_mutated(resultsdir):
    """the reference dies on a second Transport in one process and mutates its inputs (SURVEY C-7)."""
    kw = workloads.c1()
    snapshot = copy.deepcopy({k: kw[k] for k in ('species', 'electrode_reactions', 'electrolyte_reactions')})
    a = Transport(resultsdir=resultsdir, **kw)
    b = Transport(resultsdir=resultsdir, **kw)
    assert np.array_equal(a.c0, b.c0)
    assert kw['species'] == snapshot['species']
    assert kw['electrode_reactions'] == snapshot['electrode_reactions']
    assert kw['electrolyte_reactions'] == snapshot['electrolyte_reactions']
    from catint_b200.data import tp_ref_data
    assert isinstance(tp_ref_data['electrolyte_reactions']['bicarbonate-base']['buffer-base']['reaction'], str)


def test_results_folder_numbering_and_pickles(resultsdir):
    a = Transport(resultsdir=resultsdir, model_name='CO2R', **workloads.c1())
    b = Transport(resultsdir=resultsdir, model_name='CO2R', **workloads.c1())
    assert a.outputfoldername.endswith('CO2R_results')
    assert b.outputfoldername.endswith('CO2R_results_0002')
    b.tmesh = np.arange(0, 1, 0.1)
    b.save()
    for name in ('alldata', 'species', 'system', 'descriptors', 'xmesh', 'tmesh', 'electrode_reactions',
                 'electrolyte_reactions', 'comsol_outputs'):
        assert os.path.isfile(os.path.join(b.outputfoldername, name + '.pkl'))
    from catint_b200.catint_io import read_all
    c = Transport(only_plot=True)
    read_all(c, b.outputfoldername)
    assert np.array_equal(c.xmesh, b.xmesh) and c.nx == b.nx


@pytest.mark.parametrize('bad', [
    dict(species={'K+': {'bulk_concentration': 1.0, 'no such key': 1}}),
    dict(system={'not a system key': 1}),
])
def test_unknown_keys_exit_like_the_reference(bad, resultsdir):
    with pytest.raises(SystemExit):
        Transport(resultsdir=resultsdir, **bad)


def test_electrolyte_reactions_requested_but_missing_exits(resultsdir):
    with pytest.raises(SystemExit):
        Transport(resultsdir=resultsdir, species={'K+': {'bulk_concentration': 1.0}, 'Cl-': {'bulk_concentration': 1.0}},
                  system={'electrolyte reactions': True})


def test_two_fluxes_for_one_reaction_exits(resultsdir):
    kw = workloads.c1()
    kw['species']['CO2']['flux'] = 1e-5
    with pytest.raises(SystemExit):
        Transport(resultsdir=resultsdir, **kw)


def test_descriptor_grid(resultsdir):
    kw = workloads.c1()
    kw['descriptors'] = {'phiM': [-0.5, -0.6, -0.7]}
    tp = Transport(resultsdir=resultsdir, **kw)
    assert list(tp.descriptors) == ['phiM', 'temperature']          # padded with a dummy second descriptor
    assert tp.alldata_names == [[-0.5, 298.], [-0.6, 298.], [-0.7, 298.]]
    assert len(tp.alldata) == 3 and set(tp.alldata[0]) == {'species', 'system'}
    with pytest.raises(SystemExit):
        Transport(resultsdir=resultsdir, descriptors={'phiM': -0.5}, **workloads.c1())
    kw = workloads.c1()
    kw['descriptors'] = {'phiM': [0.], 'temperature': [1.], 'bulk_pH': [7.]}
    with pytest.raises(SystemExit):
        Transport(resultsdir=resultsdir, **kw)


def test_no_descriptors_is_a_single_point(resultsdir):
    tp = Transport(resultsdir=resultsdir, **workloads.c1())
    assert tp.descriptors == {'phiM': [-0.9], 'temperature': [298.]}


def test_parsers():
    assert parse_reaction('CO2 + 6 H2O + 8 e- -> CH4 + 8 OH-') == \
        ([['CO2'] + ['H2O'] * 6 + ['e-'] * 8, ['CH4'] + ['OH-'] * 8], 8)
    assert parse_reaction('HCO3- + OH- <-> CO32- + H2O') == ([['HCO3-', 'OH-'], ['CO32-', 'H2O']], None)
    assert reaction_species('CO2 + H2O + 2 e- -> CO + 2 OH-') == ['CO2', 'H2O', 'e-', 'CO', 'OH-']
    assert [charge_from_symbol(s) for s in ('K^+', 'CO_3^{2-}', 'H_2', 'PO_4^{3-}', 'Ca^{2+}', 'OH^-')] == \
        [1, -2, 0, -3, 2, -1]


def test_default_species_and_debye_mesh(resultsdir):
    tp = Transport(resultsdir=resultsdir)
    assert tp.nspecies == 2 and tp.nx in (101, 102)
    assert tp.xmax == pytest.approx(tp.debye_length * 10)
    phi, grad = tp.gouy_chapman(0.0, phiM=-0.05)
    assert phi == pytest.approx(-0.05, rel=1e-9)


def test_derive_for_refreshes_what_the_reference_does_not(resultsdir):
    """sweeping phiM / bulk_pH / temperature / boundary thickness re-derives the dependent data
    (the reference leaves it stale, SURVEY C-9)."""
    kw = workloads.c2(n_potentials=4)
    tp = Transport(resultsdir=resultsdir, **kw)
    m = tp.derive_for(phiM=-1.2)
    assert m.pb_bound['potential']['wall'] == -1.2
    i_co = min(10. * 10 ** (-(-1.2 + 0.9) / 0.12), 150.)
    names = list(m.species)
    assert m.flux_bound[names.index('CO'), 0] == pytest.approx(i_co / 2 / 96485.33289, rel=1e-14)
    m2 = tp.derive_for(bulk_pH=7.5)
    su = load_golden('ref_c1_pH7p5.npz')
    assert m2.species['H+']['bulk_concentration'] == su['c_bulk'][names.index('H+')]
    m3 = tp.derive_for(temperature=320.)
    assert m3.beta == float(load_golden('ref_c1_T320.npz')['beta'])
    m4 = tp.derive_for(**{'boundary thickness': 30e-6})
    assert m4.nx == 102
