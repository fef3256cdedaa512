"""Transport: definition of the 1D mass-transport model (host side).

Drop-in for the reference's ``catint.transport.Transport``
(/root/reference/catint/transport.py:38-1521): same constructor keywords,
same species/system/descriptor dict keys and semantics, same attributes a
solver backend reads (species, nspecies, charges, D, mu, beta, eps,
xmesh/dx/nx/xmax, c0, flux_bound, pb_bound, descriptors, alldata[_names], ...).

Design differences (deliberate, B200 batch backend):
  * model derivation is a pure function of the input dicts
    (``derive_model``) so that ``Calculator.run()`` can re-derive the
    per-cell parameters (bulk equilibria, beta, mesh, fluxes, Poisson BCs) for
    every point of a descriptor sweep without constructing 64k Transports
    -- the reference does not refresh them at all;
  * the module-level reaction library is never mutated (reference: C-7), input
    dicts are copied, the object is re-entrant;
  * errors are logged and then raise ``SystemExit`` exactly like the
    reference's ``logger.error(...); sys.exit()``.
"""
import collections
import copy
import functools
import logging
import math
import os
import re
import subprocess
import sys
from shutil import copy as _copyfile

import numpy as np
from scipy.optimize import fsolve

from .units import unit_R, unit_F, unit_eps0, unit_NA, unit_T
from . import data as _data
from .catint_io import save_all

use_mpi = False     # the reference hard-disables MPI (transport.py:27-36); sharding is done by torch.distributed

SPECIES_KEYS = ['bulk_concentration', 'diffusion', 'name', 'symbol', 'flux', 'current density',
                'flux-equation', 'MPB_radius', 'catmap_symbol', 'Henry constant']
# derived per-species entries a re-used dict may already carry
_SPECIES_DERIVED_KEYS = ['charge', 'surface_concentration', 'surface_activity_coefficient']

SYSTEM_KEYS = [
    'phiM',                    # V
    'Stern capacitance',       # micro F/cm^2
    'Stern epsilon',
    'bulk_pH',
    'phiPZC',                  # V
    'temperature',             # K
    'pressure',                # bar
    'water viscosity',
    'electrolyte viscosity',
    'epsilon',                 # eps_0
    'migration',
    'field dependence',
    'electrode reactions',
    'electrolyte reactions',
    'boundary thickness',      # m
    'exclude species',
    'active site density',     # mol/m^2
    'current density',
    'flow rate',
    'RF',
    'potential drop',
    'Stern_efield',
    'charging_scheme',
    'use_activities',
    'Stern_potential',
    'init_folder',
]

SYSTEM_DEFAULTS = {
    'epsilon': 78.36,
    'Stern epsilon': 2.0,
    'Stern capacitance': 18e-2,
    'temperature': 298.14,
    'phiM': 0.0,
    'phiPZC': 0.0,
    'electrode reactions': False,
    'electrolyte reactions': False,
    'exclude species': ['H2O', 'e-'],
    'pressure': 1,
    'field dependence': None,
    'init_folder': None,
    'Stern_efield': 0.0,
    'charging_scheme': 'comsol',
    'use_activities': True,
    'Stern_potential': 0.0,
    'potential drop': 'Stern',
}

_SPECIES_TOKEN = re.compile('([a-zA-Z]{1,10}[a-zA-Z-+0-9]+)')
_SPECIES_TOKEN0 = re.compile('([a-zA-Z]{0,10}[a-zA-Z-+0-9]+)')


class _Fatal(object):
    """reference error protocol: log, then exit (transport.py:199-200 etc.)."""

    def __init__(self, logger):
        self.logger = logger

    def __call__(self, *lines):
        for ln in lines:
            self.logger.error('| CI | -- | ' + ln)
        sys.exit()


# ----------------------------------------------------------------------
# parsing helpers
# ----------------------------------------------------------------------
@functools.lru_cache(maxsize=4096)
def _reaction_species(string):
    out = []
    for side in string.split('->'):
        for term in side.split(' + '):
            out.append(_SPECIES_TOKEN.findall(term.strip())[0])
    return tuple(out)


def reaction_species(string):
    """species tokens of a reaction string, in order of appearance, as the
    reference discovers them (transport.py:553-555,582-586).  (Parsed once per string: a sweep re-derives
    the model for every cell.)"""
    return list(_reaction_species(string))


def parse_reaction(string):
    """'A + 2 B <-> C' -> ([[A,B,B],[C]], nel) (transport.py:1098-1132);
    nel = number of electrons if an 'n e-' term is present else None."""
    sides, nel = _parse_reaction(string)
    return [list(names) for names in sides], nel


@functools.lru_cache(maxsize=4096)
def _parse_reaction(string):
    if '<->' in string:
        sides = sum([part.split('->') for part in string.split('<->')], [])
    else:
        sides = string.split('->')
    nel = None
    parsed = []
    for side in sides:
        names = []
        for term in side.strip().split(' + '):
            m_el = re.findall('([0-9]+)[ ]+e-', term)
            if m_el:
                nel = int(m_el[0])
            m_n = re.findall('([0-9]+)[ ]+[*A-Za-z]+', term)
            if not m_n:
                names.append(term.strip())
            else:
                mult = int(m_n[0])
                names.extend([term[len(str(mult)) + 1:].strip()] * mult)
        parsed.append(tuple(names))
    return tuple(parsed), nel


def charge_from_symbol(symbol):
    """integer charge from 'K^+', 'CO_3^{2-}', 'H_2' (transport.py:1240-1276)."""
    parts = symbol.split('^')
    if len(parts) == 1:
        return 0
    tail = parts[-1].replace('{', '').replace('}', '')
    if tail[-1] == '-':
        return -int(tail[:-1]) if len(tail) > 1 else -1
    if tail[-1] == '+':
        return int(tail[:-1]) if len(tail) > 1 else 1
    return int(tail)


# ----------------------------------------------------------------------
# the derivation: dicts -> model arrays
# ----------------------------------------------------------------------
class DerivedModel(object):
    """everything the solver backend needs for ONE set of system values."""
    pass


def _prepare_system(system, fatal):
    if system is None:
        system = copy.deepcopy(SYSTEM_DEFAULTS)
    else:
        for key in system:
            if key not in SYSTEM_KEYS:
                fatal('No such key "' + key + '" in system list. Quitting here.',
                      'Current system list = {}'.format(SYSTEM_KEYS))
    for key in SYSTEM_DEFAULTS:
        if key not in system:
            system[key] = copy.deepcopy(SYSTEM_DEFAULTS[key])
    for never in ('e-', 'H2O'):
        if never not in system['exclude species']:
            system['exclude species'] = list(system['exclude species']) + [never]
    return system


def _split_electrolyte_request(electrolyte_reactions):
    """['name', {'constraints':..}, {'additional_cell_reactions': name}] ->
    (names for the bulk equilibrium, constraints, additional name)
    (transport.py:563-577).  Does not touch the caller's list."""
    names, constraints, additional = [], None, None
    for entry in electrolyte_reactions:
        if isinstance(entry, dict):
            if 'constraints' in entry:
                constraints = entry['constraints']
            if 'additional_cell_reactions' in entry:
                additional = entry['additional_cell_reactions']
        else:
            names.append(entry)
    return names, constraints, additional


def _solve_bulk_equilibria(species, unknowns, eq_names, library, constraints, electrolyte_species,
                           exclude, logger, fatal):
    """buffer equilibria  prod(c_products)/prod(c_educts) = K  for the species
    without a bulk concentration (transport.py:637-735): same residuals, same
    initial guess (1,...,1), scipy fsolve (MINPACK hybrd) like the reference.
    Unlike the reference the convergence flag is inspected (SURVEY C-8)."""
    eqs = []
    for group in eq_names:
        for rname in library[group]:
            entry = library[group][rname]
            lhs, rhs = entry['reaction'].split('->')
            educts = [_SPECIES_TOKEN0.findall(t.strip())[0] for t in lhs.split(' + ')]
            products = [_SPECIES_TOKEN0.findall(t.strip())[0] for t in rhs.split(' + ')]
            eqs.append(([e for e in educts if e not in exclude],
                        [p for p in products if p not in exclude], entry['constant']))
    n_con = len(constraints) if constraints is not None else 0
    if len(unknowns) != len(eqs) + n_con:
        fatal('Number of unknown concentrations {} does not match the number of buffer equilibria equations {}. '
              'Cannot determine missing concentrations'.format(len(unknowns), len(eqs) + n_con),
              'These are the unknowns = {}'.format(unknowns))
    if len(unknowns) > 4:
        fatal('More than 4 unknowns in the buffer concentrations are not implemented yet')

    def conc(sp, var):
        if 'bulk_concentration' in species[sp]:
            return species[sp]['bulk_concentration']
        return var[sp]

    def residuals(p):
        p = np.atleast_1d(p)
        var = dict(zip(unknowns, p))
        out = []
        for educts, products, K in eqs:
            num = 1
            for sp in products:
                num *= conc(sp, var)
            den = 1
            for sp in educts:
                den *= conc(sp, var)
            out.append(num / den - K)
        if constraints is not None:
            total = 0.0
            for sp in electrolyte_species:
                if sp in exclude:
                    continue
                total += conc(sp, var) * species[sp]['charge']
            for con in constraints:
                if con == 'counter_ion_concentration':
                    out.append(constraints[con] + total)
        return tuple(out)

    # A sweep re-derives the model for every cell and most descriptors (phiM, fluxes, boundary thickness) leave the
    # bulk composition alone: the solve is a deterministic function of the values below, so it is done once per
    # distinct set of them (the cached numbers are the ones fsolve returned: bit-identical results).
    key = (tuple(unknowns), tuple((tuple(e), tuple(p), K) for e, p, K in eqs),
           tuple((sp, species[sp].get('bulk_concentration', None), species[sp].get('charge', None))
                 for sp in species),
           None if constraints is None else tuple(sorted(constraints.items())),
           tuple(electrolyte_species), tuple(exclude))
    try:
        hit = _BULK_CACHE.get(key)
    except TypeError:                     # an unhashable entry (array-valued input): no caching
        key, hit = None, None
    if hit is None:
        sol, _, ier, msg = fsolve(residuals, (1,) * len(unknowns), full_output=True)
        hit = (tuple(np.atleast_1d(sol)), ier, msg)
        if key is not None:
            if len(_BULK_CACHE) >= 8192:
                _BULK_CACHE.clear()
            _BULK_CACHE[key] = hit
    sol, ier, msg = hit
    if ier != 1:
        logger.warning('| CI | -- | bulk buffer equilibria did not converge ({}); the reference would silently '
                       'continue with these values (SURVEY C-8)'.format(msg.replace('\n', ' ')))
    return dict(zip(unknowns, sol))


_BULK_CACHE = {}


def derive_model(species_in, electrode_reactions_in, electrolyte_reactions_in, system_in, pb_bound_in,
                 nx_in, logger, tables, quiet=False, flux_env=None, owned=False):
    """dicts -> DerivedModel.  Pure: inputs are deep-copied (owned=True: species_in / system_in already are the
    caller's private copies -- derive_for -- and are used as they are).  Follows the
    order of operations of the reference constructor (transport.py:183-509),
    which matters (e.g. the bulk_pH override happens after charge neutrality)."""
    fatal = _Fatal(logger)
    info = (lambda *a: None) if quiet else (lambda m: logger.info('| CI | -- | ' + m))
    m = DerivedModel()
    diff_table, henry_table, library = tables

    # ---- dictionaries -------------------------------------------------
    if species_in is None:
        species = {'species1': {'symbol': r'K^+', 'name': 'potassium', 'diffusion': 1.96e-9,
                                'kind': 'electrolyte', 'bulk_concentration': 0.001 * 1000.},
                   'species2': {'symbol': r'HCO_3^-', 'name': 'bicarbonate', 'diffusion': 1.2e-9,
                                'kind': 'electrolyte', 'bulk_concentration': 0.001 * 1000.}}
    else:
        for sp in species_in:
            for key in species_in[sp]:
                if key not in SPECIES_KEYS and key not in _SPECIES_DERIVED_KEYS:
                    fatal('No such key "' + key + '" in species list. Quitting here.')
        species = species_in if owned else copy.deepcopy(species_in)
    species = collections.OrderedDict(species)
    system = _prepare_system(system_in if owned else copy.deepcopy(system_in), fatal)
    exclude = system['exclude species']
    for es in exclude:
        species.pop(es, None)
    info('Excluding {} from PNP transport. They will also not participate in reactions (activity = 1)'.format(exclude))
    electrode_reactions = copy.deepcopy(electrode_reactions_in)
    if pb_bound_in is None:
        pb_bound_in = {'potential': {'wall': 'phiM'}, 'gradient': {'bulk': 0.0}}

    # ---- species discovery, constants, bulk composition (transport.py:533-786) ----
    m.use_mpb = any('MPB_radius' in species[sp] for sp in species)
    reacting = []
    if electrode_reactions is not None:
        for er in electrode_reactions:
            for sp in reaction_species(electrode_reactions[er]['reaction']):
                reacting.append(sp)
                if sp not in species and sp not in exclude:
                    species[sp] = {}
    eq_names, constraints, additional = [], None, None
    electrolyte_species = []
    if electrolyte_reactions_in is not None:
        eq_names, constraints, additional = _split_electrolyte_request(electrolyte_reactions_in)
        for group in eq_names:
            for rname in library[group]:
                for sp in reaction_species(library[group][rname]['reaction']):
                    if sp not in electrolyte_species:
                        electrolyte_species.append(sp)
                    if sp not in species and sp not in exclude:
                        species[sp] = {}
    for sp in species:
        if 'diffusion' not in species[sp]:
            if sp not in diff_table:
                fatal('No diffusion constant for {}. Either add it to the diffusion constant table '
                      'or manually provide it as an input'.format(sp))
            species[sp]['diffusion'] = diff_table[sp][1]
        if 'name' not in species[sp]:
            species[sp]['name'] = diff_table[sp][0] if sp in diff_table else sp
        if 'symbol' not in species[sp]:
            species[sp]['symbol'] = diff_table[sp][2]
    for sp in henry_table:
        if sp in species:
            species[sp]['Henry constant'] = henry_table[sp] * 1e5     # mol/m^3/bar
    for sp in species:
        if ('Henry constant' not in species[sp] and sp in reacting and sp not in exclude
                and sp not in ['OH-', 'H+']):
            fatal('No Henry constant found for {}. Add it to the Henry constant table'.format(sp))
    for sp in species:
        if species[sp].get('bulk_concentration', None) == 'Henry':
            if 'Henry constant' not in species[sp]:
                fatal('Henry constant was selected for initializing bulk concentrations of {}, '
                      'but no Henry constant is known'.format(sp))
            species[sp]['bulk_concentration'] = species[sp]['Henry constant'] * system['pressure']
    for sp in species:
        species[sp]['charge'] = charge_from_symbol(species[sp]['symbol'])
    m.charges = np.array([species[sp]['charge'] * unit_F for sp in species])

    all_rx_names = list(eq_names)
    if electrolyte_reactions_in is not None:
        unknowns = [sp for sp in electrolyte_species
                    if sp not in exclude and 'bulk_concentration' not in species[sp]]
        if unknowns:
            sol = _solve_bulk_equilibria(species, unknowns, eq_names, library, constraints,
                                         electrolyte_species, exclude, logger, fatal)
            for sp in unknowns:
                species[sp]['bulk_concentration'] = sol[sp]
        if additional is not None:
            all_rx_names.append(additional)
            # species appearing only in the additional reactions must exist too
            for rname in library[additional]:
                for sp in reaction_species(library[additional][rname]['reaction']):
                    if sp not in species and sp not in exclude:
                        fatal('Species {} has not been defined, but is used in the electrolyte reactions, '
                              'define it first!'.format(sp))

    neutral_by = [sp for sp in species if species[sp].get('bulk_concentration', None) == 'charge_neutrality']
    if neutral_by:
        # round so that the bulk is exactly neutral in the stored digits (transport.py:748-755)
        for sp in species:
            c = species[sp].get('bulk_concentration', None)
            if c is not None and not isinstance(c, str):
                rounded = round(c, 8)
                if rounded != c:
                    if not quiet:
                        logger.warning('Rounded concentration of species {} to 8 decimal numbers, '
                                       'new concentration = {}'.format(sp, rounded))
                    species[sp]['bulk_concentration'] = rounded
    if len(neutral_by) > 1:
        fatal('Only a single species can be evaluated by charge neutrality')
    for sp in neutral_by:
        total = 0.
        for sp2 in species:
            c = species[sp2].get('bulk_concentration', None)
            if c is not None and not isinstance(c, str):
                total += species[sp2]['charge'] * c
        species[sp]['bulk_concentration'] = -total / species[sp]['charge']
    for sp in species:
        if 'bulk_concentration' not in species[sp]:
            if not quiet:
                logger.warning('| CI | -- | No bulk_concentration provided for species {}, setting it to zero'.format(sp))
            species[sp]['bulk_concentration'] = 0.0
    info('Updated species bulk concentrations')
    for sp in species:
        info('{} = {} mol/L'.format(sp, species[sp]['bulk_concentration'] / 1000.))
    system['reference_gas_concentration'] = 10 ** 5 / unit_R / unit_T
    m.nspecies = len(species)

    # ---- pH bookkeeping (transport.py:272-291) -------------------------
    if 'bulk_pH' in system:
        if 'H+' in species:
            species['H+']['bulk_concentration'] = 10 ** (-system['bulk_pH']) * 1000.
        elif 'OH-' in species:
            species['OH-']['bulk_concentration'] = 10 ** (-(14 - system['bulk_pH'])) * 1000.
    else:
        if 'H+' in species:
            system['bulk_pH'] = -np.log10(species['H+']['bulk_concentration'] / 1000.)
        elif 'OH-' in species:
            system['bulk_pH'] = 14 + np.log10(species['OH-']['bulk_concentration'] / 1000.)
        else:
            system['bulk_pH'] = 7.0
    system['surface_pH'] = system['bulk_pH']
    system['surface_potential'] = system['phiM']
    system['pH'] = [system['bulk_pH']]
    for sp in species:
        if 'surface_concentration' not in species[sp]:
            species[sp]['surface_concentration'] = species[sp]['bulk_concentration']
    packing = 0.
    for sp in species:
        if 'MPB_radius' in species[sp]:
            packing += species[sp]['MPB_radius'] ** 3 * species[sp]['surface_concentration'] * unit_NA
    for sp in species:
        species[sp]['surface_activity_coefficient'] = 1. / (1. - packing)

    m.eps = system['epsilon'] * unit_eps0
    m.beta = 1. / (system['temperature'] * unit_R)
    m.use_migration = bool(system['migration']) if 'migration' in system else True
    m.use_convection = 'flow rate' in system

    # ---- reactions (transport.py:326-393) ------------------------------
    m.use_electrolyte_reactions = bool(system.get('electrolyte reactions', True))
    if electrolyte_reactions_in is not None:
        rx = collections.OrderedDict()
        for group in all_rx_names:
            for rname in library[group]:
                rx[rname] = copy.deepcopy(library[group][rname])
        if m.use_electrolyte_reactions and any('rates' in rx[r] for r in rx):
            for r in rx:
                rx[r]['reaction'], _ = parse_reaction(rx[r]['reaction'])
        m.electrolyte_reactions = rx
    else:
        m.electrolyte_reactions = None
        if m.use_electrolyte_reactions:
            fatal('Electrolyte reactions were requested by input, but no electrolyte reaction was defined. '
                  'Define electrolyte reaction first.')
    if m.use_electrolyte_reactions:
        for r in m.electrolyte_reactions:
            if 'rates' not in m.electrolyte_reactions[r]:
                info('Reaction {} has no rates given. It will not be considered for PNP dynamics!'.format(r))
            sides = m.electrolyte_reactions[r]['reaction']
            if isinstance(sides, str):
                sides, _ = parse_reaction(sides)
            for sp in sum(sides, []):
                if sp not in species and sp not in exclude:
                    fatal('Species {} has not been defined, but is used in the electrolyte reactions, '
                          'define it first!'.format(sp), 'Current species list: {}'.format(list(species)))

    m.use_electrode_reactions = bool(system.get('electrode reactions', False))
    if electrode_reactions is not None:
        m.use_electrode_reactions = True
        for er in electrode_reactions:
            sides, nel = parse_reaction(electrode_reactions[er]['reaction'])
            electrode_reactions[er]['reaction'] = sides
            if nel is not None:
                electrode_reactions[er]['nel'] = nel
    elif m.use_electrode_reactions:
        fatal('Electrode reactions were requested by input, but no electrode reaction was defined. '
              'Define electrode reaction first.')
    m.electrode_reactions = electrode_reactions
    if m.use_electrode_reactions:
        for er in electrode_reactions:
            for sp in sum(electrode_reactions[er]['reaction'], []):
                if sp not in species and sp not in exclude and not sp.startswith('*'):
                    fatal('Species {} has not been defined, but is used in the electrode reactions, '
                          'define it first!'.format(sp), 'Current species list: {}'.format(list(species)))

    m.product_list, m.educt_list, m.electrolyte_list = [], [], []
    if m.use_electrode_reactions:
        for prod in electrode_reactions:
            m.product_list.append(prod)
            lhs, rhs = electrode_reactions[prod]['reaction'][0], electrode_reactions[prod]['reaction'][1]
            for sp in lhs:
                if sp != prod and sp != 'e-' and sp not in exclude and sp not in m.educt_list:
                    m.educt_list.append(sp)
            for sp in rhs:
                if sp != prod and sp != 'e-' and sp not in exclude and sp not in m.product_list:
                    m.product_list.append(sp)
    for sp in species:
        if sp not in m.product_list and sp not in m.educt_list:
            m.electrolyte_list.append(sp)

    # ---- transport coefficients & mesh (transport.py:422-460) -----------
    D = np.array([species[sp].get('diffusion', 0.0) for sp in species], dtype=float)
    if 'water viscosity' in system and 'electrolyte viscosity' in system:
        D = np.array([d * float(system['water viscosity']) / float(system['electrolyte viscosity']) for d in D])
    m.D = D
    m.mu = m.D * m.charges * m.beta
    ionic = 0.0
    for k, sp in enumerate(species):
        ionic += m.charges[k] ** 2 * species[sp]['bulk_concentration']
    m.ionic_strength = 0.5 * ionic
    with np.errstate(divide='ignore', invalid='ignore'):
        m.debye_length = np.sqrt(m.eps / m.beta / 2. / m.ionic_strength)
    m.nx_intervals = nx_in
    if 'boundary thickness' in system:
        m.boundary_thickness = system['boundary thickness']
        m.xmax = m.boundary_thickness
        m.dx = m.xmax / (nx_in * 1.)
    else:
        nx_mod = max(1., np.ceil(nx_in / 10.))
        m.xmax = m.debye_length * nx_mod
        m.dx = m.debye_length / nx_mod
    m.xmesh = np.arange(0, m.xmax + m.dx, m.dx)       # 101 *or* 102 nodes for nx_in=100 (SURVEY C-6)
    m.nx = len(m.xmesh)

    # ---- boundary fluxes (transport.py:929-1095) ------------------------
    m.use_catmap = False
    m.flux_eq = None
    m.fpar = None
    eq_owners = [sp for sp in species if 'flux-equation' in species[sp]]
    if not eq_owners or not m.use_electrode_reactions:
        _derive_fluxes(m, species, system, exclude, fatal, info)
    else:
        _derive_flux_equations(m, species, system, exclude, fatal, info, eq_owners, flux_env)

    m.species = species
    m.system = system
    # ---- BC containers (transport.py:1278-1322, 1388-1487) -------------
    m.c0 = np.repeat(np.array([species[sp]['bulk_concentration'] for sp in species], dtype=float), m.nx)
    m.pb_bound = {}
    for kind in ('potential', 'gradient'):
        m.pb_bound[kind] = {}
        for side in ('wall', 'bulk'):
            val = None
            if kind in pb_bound_in and side in pb_bound_in[kind]:
                val = pb_bound_in[kind][side]
                if isinstance(val, str) and val == 'phiM':
                    val = system['phiM']
            m.pb_bound[kind][side] = val
    m.boundary_type = 'flux'
    fluxes_numeric = not any(isinstance(species[sp]['flux'], str) for sp in species)
    if m.flux_eq is not None:
        # flux equations: flux_bound carries the FIXED part, the expressions are evaluated by the solver backend
        fb = np.zeros([m.nspecies, 2])
        fb[:, 0] = m.flux_fixed
        m.flux_bound = fb
    elif fluxes_numeric:
        fb = np.zeros([m.nspecies, 2])
        fb[:, 0] = [species[sp]['flux'] for sp in species]
        m.flux_bound = fb
    else:
        m.flux_bound = None
    dc = np.zeros([m.nspecies, 2])
    m.dc_dt_bound = dc * 10 ** 3                      # 'all': {'r': 0.0}
    m.efield_bound = np.array([0.0 * 1e10, None], dtype=object)
    return m


def _derive_fluxes(m, species, system, exclude, fatal, info):
    """electrode boundary fluxes from current densities and stoichiometry."""
    if not m.use_electrode_reactions:
        for sp in species:
            species[sp]['flux'] = 0.0
        return
    ers = m.electrode_reactions
    if any(species[sp].get('flux', None) == 'catmap' for sp in species):
        m.use_catmap = True
        for sp in species:
            species[sp].setdefault('flux', 'catmap')
        info('Found flux = catmap, all fluxes will be calculated by CatMAP.')
        return
    flux_keys = ('flux', 'current density', 'flux-equation')
    for sp in species:
        if sum(1 for key in species[sp] if key in flux_keys) > 1:
            fatal('Flux of species {} has been defined by more than one method.'.format(sp))
    products_named = [prod.split('-')[0] for prod in ers]
    for sp in species:
        if 'current density' in species[sp] and sp not in products_named:
            m_logger_error = 'Flux of species {} has been given as current density but this species is not product.'.format(sp)
            logging.getLogger('transport.info').error('| CI | -- | ' + m_logger_error)
    reduction = None
    for er in ers:
        lhs, rhs = ers[er]['reaction'][0], ers[er]['reaction'][1]
        defined = 0
        for sp in lhs + rhs:
            if sp in exclude or '*' in sp:
                continue
            if any(key in flux_keys for key in species[sp]):
                defined += 1
        if defined > 1:
            fatal('More than one flux has been defined for equation {}. Select one of the fluxes, '
                  'the rest will be automatically calculated.'.format(ers[er]['reaction']))
        if defined == 0:
            fatal('No flux defined in equation {}. Define one flux.'.format(ers[er]['reaction']))
        if 'e-' in lhs:
            reduction = True
        elif 'e-' in rhs:
            reduction = False
        else:
            fatal('No electron found in the reactions.')
    symbolic = any('flux-equation' in species[sp] for sp in species)

    # 1. given current densities -> fluxes:  sign * i / nel / F * nprod
    for sp in list(species):
        if symbolic:
            if 'flux' in species[sp]:
                species[sp]['flux'] = str(species[sp]['flux'])
            elif 'current density' in species[sp]:
                for er in ers:
                    if sp == er.split('-')[0]:
                        nprod = ers[er]['reaction'][1].count(sp)
                        species[sp]['flux'] = ('(-1)' if reduction else '1') + '*' + str(
                            species[sp]['current density'] / ers[er]['nel'] / unit_F * nprod)
            elif 'flux-equation' in species[sp]:
                species[sp]['flux'] = species[sp]['flux-equation']
        elif 'current density' in species[sp]:
            for er in ers:
                if sp == er.split('-')[0]:
                    nprod = ers[er]['reaction'][1].count(sp)
                    species[sp]['flux'] = (-1 if reduction else 1) * species[sp]['current density'] \
                        / ers[er]['nel'] / unit_F * nprod

    # 2. remaining reactants by stoichiometric ratio and side
    missing = []
    for er in ers:
        for sp in sum(ers[er]['reaction'], []):
            if sp not in ers and sp != 'e-' and sp not in exclude and sp not in missing and not sp.startswith('*'):
                missing.append(sp)
    if missing:
        info('Calculating fluxes of {} as sum of other fluxes'.format(missing))
    # the reference species of every reaction is fixed before any derived flux is written
    ref_of = {}
    for er in ers:
        for sp in ers[er]['reaction'][0] + ers[er]['reaction'][1]:
            if sp in exclude or '*' in sp:
                continue
            if any(key in flux_keys for key in species[sp]):
                ref_of[er] = sp
                break
    for er in ers:
        lhs, rhs = ers[er]['reaction'][0], ers[er]['reaction'][1]
        ref = ref_of[er]
        for sp in missing:
            if sp not in lhs and sp not in rhs:
                continue
            n_missing = max(lhs.count(sp), rhs.count(sp)) * 1.
            n_ref = max(lhs.count(ref), rhs.count(ref)) * 1.
            same_side = (sp in lhs and ref in lhs) or (sp in rhs and ref in rhs)
            if symbolic:
                species[sp].setdefault('flux', '0')
                species[sp]['flux'] += '+' + ('1' if same_side else '(-1)') + '*' + species[ref]['flux'] \
                    + '*' + str(n_missing / n_ref)
            else:
                species[sp].setdefault('flux', 0.0)
                species[sp]['flux'] += (1 if same_side else -1.) * species[ref]['flux'] * n_missing / n_ref
    for sp in species:
        if 'flux' not in species[sp]:
            species[sp]['flux'] = '0.0' if symbolic else 0.0


def _derive_flux_equations(m, species, system, exclude, fatal, info, eq_owners, flux_env):
    """species[sp]['flux-equation'] (docs/source/topics/flux_definition.rst:100-156): the flux of the owner is
    the expression RF*flux_factor*(...) (comsol_model.py:1000), the fluxes of the other reactants follow from
    the stoichiometry (transport.py:1057-1087).  That propagation is linear in the given fluxes, so it is run
    numerically (same code as for fixed fluxes) once with all expressions at zero -- the fixed part -- and once
    per expression at one: J = fixed + sum_e coef[:, e]*E_e.  The expressions are compiled for the device."""
    from .fluxeq import FluxEquations, FluxEqError, parameter_values
    params, variables = flux_env if flux_env is not None else ({}, {})
    names = list(species)

    def propagate(values):
        sp2 = copy.deepcopy(species)
        for o in eq_owners:
            sp2[o].pop('flux-equation')
            sp2[o]['flux'] = values[o]
        _derive_fluxes(m, sp2, system, exclude, fatal, lambda *a: None)
        return np.array([float(sp2[s]['flux']) for s in names])

    fixed = propagate({o: 0.0 for o in eq_owners})
    fe = FluxEquations(names)
    coef = np.zeros((len(names), len(eq_owners)))
    try:
        for e, o in enumerate(eq_owners):
            coef[:, e] = propagate({oo: (1.0 if oo == o else 0.0) for oo in eq_owners}) - fixed
            fe.add(o, 'RF*flux_factor*(' + species[o]['flux-equation'] + ')', variables)
        fe.coef = coef
        m.fpar = np.array(parameter_values(fe.par_names, system, params, m), dtype=float)
    except FluxEqError as err:
        fatal(str(err))
    m.flux_eq = fe
    m.flux_fixed = fixed
    for k, sp in enumerate(names):
        species[sp]['flux'] = species[sp]['flux-equation'] if sp in eq_owners else float(fixed[k])
    info('Fluxes of {} are given as flux equations, evaluated on the surface state by the solver'.format(eq_owners))


# ----------------------------------------------------------------------
class Transport(object):

    def __init__(self, catint_path=None, species=None, electrode_reactions=None, electrolyte_reactions=None,
                 system=None, pb_bound=None, nx=100,
                 descriptors=None, model_name=None,
                 comsol_args={}, catmap_args={}, only_plot=False, resultsdir=None):
        """Same keywords as the reference (transport.py:40-43).
        only_plot   only initialize transport without creating folders
        resultsdir  working directory where to save all outputs
        """
        if only_plot:
            return
        if catint_path is None:
            catint_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        self.catint_path = catint_path
        # one process per GPU under torch.distributed (the reference's mpi_rank/mpi_size, transport.py:27-36,54-66)
        from . import distributed as _dist
        self.mpi_rank, self.mpi_size = _dist.world()

        self.model_name = 'catint' if model_name is None else model_name
        self._make_results_folder(resultsdir)
        self._setup_logging()

        # keep the user's inputs (copies) so sweeps can re-derive the model
        self._inputs = dict(
            species=copy.deepcopy(species), electrode_reactions=copy.deepcopy(electrode_reactions),
            electrolyte_reactions=copy.deepcopy(electrolyte_reactions), system=copy.deepcopy(system),
            pb_bound=copy.deepcopy(pb_bound), nx=nx)
        self._tables = self._load_tables()
        # parameters / variables flux equations may refer to (the reference's comsol_args, comsol_model.py:962-978)
        ca = comsol_args if comsol_args is not None else {}
        variables = {}
        for key in ('global_variables', 'boundary_variables'):
            for name, val in (ca.get(key, None) or {}).items():
                variables[name] = val[0] if isinstance(val, (list, tuple)) else val
        self._flux_env = (dict(ca.get('parameter', None) or {}), variables)

        model = self.derive_for(_quiet=False)
        self._adopt(model)
        if self.debye_length > self.xmax / 2.:
            self.logger.warning('| CI | -- | Debye length is larger than 1/4th of the xmesh. Take care that the x '
                                'discretization is not too coarse!.')
            self.logger.warning('| CI | -- | Current xmesh: xmax={}, dx={} at debye_length={}'.format(
                self.xmax, self.dx, self.debye_length))
        self.xmesh_init = self.xmesh
        self.nx_init = self.nx
        self.xmax_init = self.xmax
        self.external_charge = np.zeros([len(self.xmesh)])
        self.count = 1
        self.phiM_init = None

        self.system['efield'] = np.zeros([self.nx])
        self.system['potential'] = np.zeros([self.nx])
        self.system['charge_density'] = np.zeros([self.nx])

        if 'RF' not in self.system:
            self.system['RF'] = 1.0                   # roughness factor default (transport.py:844-851)
        self.initialize_descriptors(descriptors)
        self.catmap_args = catmap_args
        self.comsol_args = self._comsol_defaults(comsol_args)
        self.calc = None

    # ------------------------------------------------------------------
    def _make_results_folder(self, resultsdir):
        """<resultsdir>/<model_name>_results[_NNNN] (transport.py:70-106).  With several ranks the folder is
        created by rank 0 only and its name is broadcast (reference: transport.py:54-110 creates it on rank 0,
        broadcasts and barriers); before the process group exists the other ranks get a private
        <model_name>_results.rankNNN folder for their log file, so nothing races."""
        root = os.getcwd()
        if resultsdir is None:
            resultsdir = root
        os.makedirs(resultsdir, exist_ok=True)
        self.inputfilename = sys.argv[0] if len(sys.argv) else ''
        from . import distributed as _dist
        shared = self.mpi_size > 1 and _dist.group_ready()
        if self.mpi_rank == 0:
            self.outputfoldername = resultsdir + '/' + self.model_name + '_results'
            if not os.path.exists(self.outputfoldername):
                os.makedirs(self.outputfoldername)
            else:
                pat = re.compile(self.model_name + '_results_[0-9]+')
                numbered = sorted(f for f in os.listdir(resultsdir) if pat.search(f))
                number = int(numbered[-1].split('_')[-1]) + 1 if numbered else 2
                self.outputfoldername = self.outputfoldername + '_' + str(number).zfill(4)
                os.makedirs(self.outputfoldername)
        elif not shared:
            self.outputfoldername = resultsdir + '/' + self.model_name + '_results.rank' + str(self.mpi_rank).zfill(3)
            os.makedirs(self.outputfoldername, exist_ok=True)
        if shared:
            import torch.distributed as dist
            name = [self.outputfoldername if self.mpi_rank == 0 else None]
            dist.broadcast_object_list(name, src=0)
            self.outputfoldername = name[0]
            dist.barrier()
        self.logfilename = self.outputfoldername + '/transport.log'
        if self.inputfilename and os.path.isfile(self.inputfilename):
            try:
                _copyfile(self.inputfilename, self.outputfoldername + '/' + os.path.basename(self.inputfilename))
            except (OSError, IOError):
                pass

    def _setup_logging(self):
        """loggers 'transport.info' / 'transport.debug', per-rank file
        transport_idNNN.log, INFO console handler (transport.py:112-139)."""
        logfile = '.'.join(self.logfilename.split('.')[:-1]) + '_id' + str(self.mpi_rank).zfill(3) + '.log'
        root_logger = logging.getLogger('')
        root_logger.setLevel(logging.DEBUG)
        for h in list(root_logger.handlers):
            if getattr(h, '_catint_handler', False):
                root_logger.removeHandler(h)
                h.close()
        fh = logging.FileHandler(logfile, mode='w')
        fh.setLevel(logging.DEBUG)
        fh.setFormatter(logging.Formatter('%(asctime)s %(name)-12s %(levelname)-8s %(message)s', datefmt='%m-%d %H:%M'))
        fh._catint_handler = True
        root_logger.addHandler(fh)
        console = logging.StreamHandler()
        console.setLevel(logging.INFO if not os.environ.get('CATINT_QUIET') else logging.WARNING)
        console.setFormatter(logging.Formatter('%(name)-12s: %(levelname)-8s %(message)s'))
        console._catint_handler = True
        root_logger.addHandler(console)
        try:
            version = subprocess.check_output(['git', 'describe', '--always'], stderr=subprocess.DEVNULL,
                                              cwd=os.path.dirname(os.path.abspath(__file__))).strip()
            logging.info('Starting Transport Calculation. Current Version: {}'.format(version))
        except Exception:
            logging.info('Starting Transport Calculation. Current Version not available.')
        self.logger_db = logging.getLogger('transport.debug')
        self.logger = logging.getLogger('transport.info')

    def _load_tables(self):
        diff = dict(_data.DIFFUSION_CONSTANTS)
        henry = dict(_data.HENRY_CONSTANTS)
        dpath = os.path.join(self.catint_path, 'data', 'diffusion_constants.txt')
        hpath = os.path.join(self.catint_path, 'data', 'henry_constants.txt')
        if os.path.isfile(dpath):
            diff.update(_data.read_diffusion_file(dpath))
        if os.path.isfile(hpath):
            henry.update(_data.read_henry_file(hpath))
        return diff, henry, _data.electrolyte_reaction_library()

    def _adopt(self, m):
        for name in ('species', 'system', 'nspecies', 'charges', 'D', 'mu', 'beta', 'eps', 'use_mpb',
                     'use_migration', 'use_convection', 'use_electrolyte_reactions', 'use_electrode_reactions',
                     'electrolyte_reactions', 'electrode_reactions', 'product_list', 'educt_list',
                     'electrolyte_list', 'ionic_strength', 'debye_length', 'xmax', 'dx', 'xmesh', 'nx',
                     'use_catmap', 'c0', 'pb_bound', 'boundary_type', 'flux_bound', 'dc_dt_bound',
                     'efield_bound', 'flux_eq', 'fpar'):
            setattr(self, name, getattr(m, name))
        if hasattr(m, 'boundary_thickness'):
            self.boundary_thickness = m.boundary_thickness

    def _comsol_defaults(self, comsol_args):
        """the COMSOL backend is out of scope; keep the handful of keys other
        code looks at (calculator.py:199 'desc_method', catint_io.py:88 'outputs')."""
        ca = dict(comsol_args) if comsol_args is not None else {}
        ca.setdefault('model_type', 'tp_dilute_species')
        ca.setdefault('studies', ['stat'])
        ca.setdefault('solver', 'parametric')
        for key in ('global_variables', 'boundary_variables', 'parameter'):
            ca.setdefault(key, {})
        ca['outputs'] = []
        ca.setdefault('desc_method', 'external')
        ca.setdefault('par_method', 'external')
        if 'RF' not in self.system:
            self.system['RF'] = 1.0
        return ca

    # ------------------------------------------------------------------
    def derive_for(self, _quiet=True, **system_overrides):
        """model arrays for the same inputs with some system values replaced
        (one sweep point).  Used by Calculator to build the cell batch.

        Backend extension: species[sp]['flux'] / ['current density'] /
        ['bulk_concentration'] may be a callable f(system) -> float, evaluated
        with the system values of the sweep point (descriptor-dependent fixed
        fluxes, e.g. a Tafel law; an OH- concentration that follows bulk_pH)."""
        system = copy.deepcopy(self._inputs['system']) if self._inputs['system'] is not None else None
        if system_overrides:
            system = {} if system is None else system
            system.update(system_overrides)
        species = copy.deepcopy(self._inputs['species'])
        if species is not None:
            view = dict(SYSTEM_DEFAULTS)
            view.update(system or {})
            for sp in species:
                for key in ('flux', 'current density', 'bulk_concentration'):
                    if callable(species[sp].get(key, None)):
                        species[sp][key] = float(species[sp][key](view))
        return derive_model(species, self._inputs['electrode_reactions'], self._inputs['electrolyte_reactions'],
                            system, self._inputs['pb_bound'], self._inputs['nx'], self.logger, self._tables,
                            quiet=_quiet, flux_env=self._flux_env, owned=True)

    # ------------------------------------------------------------------
    def initialize_descriptors(self, descriptors):
        """exactly two descriptor lists, alldata grid (transport.py:1135-1195)."""
        if descriptors is not None:
            if any(type(descriptors[d]) not in [list, np.ndarray] for d in descriptors):
                self.logger.error('| CI | -- | Descriptors must be given as list. Stopping here for safety')
                sys.exit()
            self.descriptors = collections.OrderedDict(
                (k, list(v)) for k, v in descriptors.items())
        else:
            self.descriptors = collections.OrderedDict()
            self.logger.warning('CI No descriptor list given at input, performing single point calculation')
            self.descriptors['phiM'] = [self.system['phiM']]
            self.descriptors['temperature'] = [self.system['temperature']]
        keys = list(self.descriptors)
        if len(keys) == 1:
            if 'temperature' not in keys:
                self.descriptors['temperature'] = [self.system['temperature']]
            else:
                self.descriptors['phiM'] = [self.system['phiM']]
        keys = list(self.descriptors)
        if len(keys) != 2:
            self.logger.error('| CI | -- | Cannot use other than 2 descriptors')
            sys.exit()
        for d in keys:
            if d not in self.system:
                self.logger.error('| CI | -- | ' + d + ' not found in system list, cannot evaluate other than '
                                  'system descriptors, yet')
        self.alldata = []
        self.alldata_names = []
        for v1 in self.descriptors[keys[0]]:
            for v2 in self.descriptors[keys[1]]:
                self.alldata_names.append([v1, v2])
                self.alldata.append({'species': {sp: {} for sp in self.species}, 'system': {}})

    # ------------------------------------------------------------------
    def gouy_chapman(self, x, phiM=None):
        """analytic Gouy-Chapman potential and its gradient (transport.py:1373-1383)."""
        if phiM is None:
            phiM = self.system['phiM']
        gamma = np.tanh(phiM * self.beta * unit_F / 4.)

        def phi(xx):
            decay = np.exp(-xx / self.debye_length)
            return 2. / (self.beta * abs(self.charges[0])) * np.log((1. + gamma * decay) / (1. - gamma * decay))
        return phi(x), (phi(x + 1e-10) - phi(x - 1e-10)) / (2 * 1e-10)

    def set_initial_concentrations(self, func, phiM=None):
        """Boltzmann-distributed initial state (transport.py:1325-1346)."""
        if func != 'Gouy-Chapman':
            return
        if phiM is None:
            phiM = self.system['phiM']
        else:
            self.phiM_init = phiM
        if self.nspecies != 2:
            self.logger.error('| CI | -- | Gouy-Chapman limit only implemented for two species, cationic'
                              'and anionic. Not applying initialization.')
            return
        c0 = np.zeros([self.nspecies * self.nx])
        for k, sp in enumerate(self.species):
            pot = self.gouy_chapman(self.xmesh, phiM=phiM)[0]
            c0[k * self.nx:(k + 1) * self.nx] = self.species[sp]['bulk_concentration'] * \
                np.exp(-self.beta * self.charges[k] * pot)
        self.c0 = c0

    def get_initial_conditions(self):
        return self.c0

    def get_boundary_conditions(self):
        return None, self.dc_dt_bound, self.efield_bound

    def set_calculator(self, calc=None):
        """Attach calculator (name string), transport.py:1514-1516."""
        self.calc = calc

    def save(self):
        if self.mpi_rank == 0:
            self.logger.info('| CI | -- | ' + 'Saving all data into binary pickle files.')
            save_all(self)
