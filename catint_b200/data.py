"""Reaction library and species constants for the transport model.

Same numerical content as the reference's data sources -- buffer reactions
with equilibrium constants and forward/backward rate constants
(/root/reference/catint/data.py:1-123), diffusion coefficients
(/root/reference/data/diffusion_constants.txt:10-44) and Henry constants
(/root/reference/data/henry_constants.txt:3-14) -- because bulk buffer
equilibria and homogeneous rates must agree with the reference to the digit.

Layout differs from the reference: reactions are built from compact rows and
the two text tables are embedded (a user-supplied ``catint_path`` with
``data/*.txt`` files still overrides them, see transport.py).

Units: concentrations mol/m^3, so second-order rate constants are m^3/mol/s
and equilibrium constants carry (mol/m^3)^(dn).
"""
import copy


def _rx(reaction, constant, rates=None):
    d = {'reaction': reaction, 'constant': constant}
    if rates is not None:
        d['rates'] = list(rates)
    return d


def _library():
    lib = {}
    # CO2/bicarbonate/carbonate buffer, alkaline route (pure water)
    lib['bicarbonate-base'] = {
        'buffer-base': _rx('CO2 + OH- <-> HCO3-', 44400.0, (5.93, 0.00013355855855855855)),
        'buffer-base2': _rx('HCO3- + OH- <-> CO32- + H2O', 4.66, (1.0e5, 21459.227467811157)),
    }
    # acidic route
    lib['bicarbonate-acid'] = {
        'buffer-acid': _rx('CO2 + H2O <-> HCO3- + H+', 0.000444, (3.7e-2, 83.33333333333333)),
        'buffer-acid2': _rx('HCO3- <-> CO32- + H+', 4.66e-8, (59.44, 1275536480.6866953)),
    }
    # phosphate (Ryu et al., 10.1002/anie.201802756)
    lib['phosphate-acid'] = {
        'phosphate-1': _rx('H3PO4 + H+ <-> H2PO4-', 0.00629 * 1000., (5.6e8, 8.9e10 / 1000.)),
        'phosphate-2': _rx('H2PO4- + H+ <-> HPO42-', 6.32e-8 * 1000., (6.32e2, 1e10 / 1000.)),
        'phosphate-3': _rx('HPO42- + H+ <-> PO43-', 4.47e-13 * 1000., (4.47e-3, 1e10 / 1000.)),
    }
    # citrate (same source, diffusion-limited estimates)
    lib['citrate-acid'] = {
        'citrate-1': _rx('H3Cit + H+ <-> H2Cit-', 0.000745 * 1000, (7.45e6, 1e10 / 1000.)),
        'citrate-2': _rx('H2Cit- + H+ <-> HCit2-', 1.73e-5 * 1000, (1.73e5, 1e10 / 1000.)),
        'citrate-3': _rx('HCit2- + H+ <-> Cit3-', 4.02e-7 * 1000, (4.02e3, 1e10 / 1000.)),
    }
    # borate: equilibrium only, no kinetics -> not part of PNP dynamics
    lib['borate-base'] = {
        'borate-1': _rx('H3BO3 + OH- <-> H2BO3- + H2O', 5.75e-10 * 1000),
        'borate-2': _rx('H2BO3- + OH- <-> HBO32- + H2O', 3.98e-13 * 1000),
        'borate-3': _rx('HBO32- + OH- <-> BO33- + H2O', 5.01e-14 * 1000),
    }
    lib['water-diss'] = {
        'self-dissociation of water': _rx('H2O <-> OH- + H+', 1e-8,
                                          (2.4e-5 * 1000., 2.4e-5 / 1e-14 / 1000.)),
    }
    return lib


# name kept for scripts that do ``from catint.data import tp_ref_data``
tp_ref_data = {'electrolyte_reactions': _library()}


def electrolyte_reaction_library():
    """fresh deep copy -- the Transport never mutates the module-level table
    (the reference does, SURVEY C-7, which makes it non re-entrant)."""
    return copy.deepcopy(tp_ref_data['electrolyte_reactions'])


# species -> (name, D [m^2/s] at 25 C, LaTeX-ish symbol carrying the charge)
DIFFUSION_CONSTANTS = {
    # gases
    'H2': ('hydrogen', 5.11e-9, 'H_2'),
    'CO2': ('carbon_dioxide', 1.91e-9, 'CO_2'),
    'CO': ('carbon_monoxide', 2.23e-9, 'CO'),
    'O2': ('oxygen', 2.42e-9, 'O_2'),
    # acids / buffers
    'H3PO4': ('phosphoric_acid', 8.8e-10, 'H_3PO_4'),
    'H2PO4-': ('h2_phosphate', 9.59e-10, 'H_2PO_4^-'),
    'HPO42-': ('h_phosphate', 7.59e-10, 'HPO_4^{2-}'),
    'PO43-': ('phophate', 8.24e-10, 'PO_4^{3-}'),
    'H3Cit': ('h3_citrate', 8.87e-10, 'H_3Cit'),
    'H2Cit-': ('h2_citrate', 7.99e-10, 'H_2Cit^-'),
    'HCit2-': ('h_citrate', 7.e-10, 'HCit^{2-}'),
    'Cit3-': ('citrate', 6.23e-10, 'Cit^{3-}'),
    'HCO3-': ('bicarbonate', 1.185e-9, 'HCO_3^-'),
    'CO32-': ('carboxylate', 0.923e-9, 'CO_3^{2-}'),
    # cations
    'Cs+': ('cesium', 2.056e-9, 'Cs^+'),
    'D+': ('deuterium', 6.655e-9, 'D^+'),
    'H+': ('hydronium', 9.311e-9, 'H^+'),
    'K+': ('potassium', 1.957e-9, 'K^+'),
    'Na+': ('sodium', 1.334e-9, 'Na^+'),
    'NH4+': ('ammonium', 1.957e-9, 'NH_4^+'),
    'Li+': ('lithium', 1.029e-9, 'Li^+'),
    'Ca2+': ('calcium', 0.792e-9, 'Ca^{2+}'),
    # anions
    'OH-': ('hydroxide', 5.273e-9, 'OH^-'),
    'Cl-': ('chloride', 2.032e-9, 'Cl^-'),
    'I-': ('iodide', 2.045e-9, 'I^-'),
    'Br-': ('bromide', 2.080e-9, 'Br^-'),
    'ClO4-': ('perchlorate', 1.792e-9, 'ClO_4^-'),
    # hydrocarbons
    'CH4': ('methane', 1.49e-9, 'CH_4'),
    'C2H4': ('ethylene', 1.51e-9, 'C_2H_4'),
    'CH3CO2H': ('acetic_acid', 1.29e-9, 'CH_3CO_2H'),
    'CH3CH2OH': ('ethanol', 1.24e-9, 'CH_3CH_2OH'),
}

# species -> Henry constant in mol/m^3/Pa (converted to mol/m^3/bar by *1e5 at use)
HENRY_CONSTANTS = {
    'CH4': 1.4e-5,
    'C2H6': 1.9e-5,
    'CH3OH': 2.0,
    'CH3CH2OH': 1.9,
    'CO': 9.7e-6,
    'CO2': 3.3e-4,
    'N2': 6.4e-6,
    'H2': 7.8e-6,
    'NH3': 5.9e-1,
    'O2': 1.2e-5,
    'CH2O': 3.2e1,
    'NO': 1.9e-5,
}


def read_diffusion_file(path):
    """same 5-column text format as the reference's data/diffusion_constants.txt"""
    out = {}
    for line in open(path, 'r', encoding='utf8'):
        if line.startswith('#') or not line.strip():
            continue
        ls = line.split()
        out[ls[1]] = (ls[0], float(ls[3]), ls[2])
    return out


def read_henry_file(path):
    out = {}
    for line in open(path, 'r', encoding='utf8'):
        if line.startswith('#') or not line.strip():
            continue
        ls = line.split()
        out[ls[1]] = float(ls[2])
    return out
