"""Named synthetic workloads (SURVEY 8d / BASELINE.json configs) as Transport keyword dicts.

C1  single cell CO2R at Au in 0.1 M-class KHCO3 (pH 6.8, CO2 saturated), L = 50 um,
    nx = 100 (101 nodes), fixed fluxes i_CO = -10, i_H2 = -5 A/m^2
C2  the same chemistry, 1024-point potential sweep phiM = linspace(-0.5,-1.5,1024) with
    Tafel-law partial currents  i_CO = -10*10^(-(phiM+0.9)/0.12), i_H2 = -5*10^(...)
    clipped at 150 A/m^2 each (the fixed-flux FD model only feels phiM through the fluxes)
C4  10 species (adds CH4 from a third electrode reaction and inert Cl-),
    bulk_pH x boundary-layer-thickness sweep

Fresh dict objects on every call.  Descriptor-dependent currents are given as
callables of the system dict (backend extension, see Transport.derive_for).
"""
import numpy as np


def tafel_current(i_ref, phi_ref=-0.9, slope=0.12, clip=150.0):
    def current(system):
        i = i_ref * 10.0 ** (-(system['phiM'] - phi_ref) / slope)
        return float(np.sign(i_ref) * min(abs(i), clip))
    return current


def co2r_inputs(pH=6.8, L=50e-6, i_CO=-10., i_H2=-5., nx=100, migration=True, reactions=True,
                temperature=298., phiM=-0.9, extra_species=False):
    system = {'temperature': temperature, 'pressure': 1.013, 'bulk_pH': pH, 'boundary thickness': L,
              'epsilon': 78.36, 'migration': migration, 'electrode reactions': True,
              'electrolyte reactions': reactions, 'phiM': phiM, 'phiPZC': 0.16, 'Stern capacitance': 20.}
    electrolyte_reactions = ['bicarbonate-base', 'water-diss',
                             {'additional_cell_reactions': 'bicarbonate-acid'}] if reactions else None
    electrode_reactions = {'CO': {'reaction': 'CO2 + H2O + 2 e- -> CO + 2 OH-'},
                           'H2': {'reaction': '2 H2O + 2 e- -> H2 + 2 OH-'}}
    species = {'K+': {'bulk_concentration': 'charge_neutrality'},
               'CO2': {'bulk_concentration': 'Henry'},
               # follows the bulk_pH of the sweep point (a plain float would stay at the input pH)
               'OH-': {'bulk_concentration': lambda system: 10 ** (system['bulk_pH'] - 14.) * 1000.},
               'CO': {'bulk_concentration': 0.0, 'current density': i_CO},
               'H2': {'bulk_concentration': 0.0, 'current density': i_H2}}
    if extra_species:
        electrode_reactions['CH4'] = {'reaction': 'CO2 + 6 H2O + 8 e- -> CH4 + 8 OH-'}
        species['CH4'] = {'bulk_concentration': 0.0, 'current density': -2.0}
        species['Cl-'] = {'bulk_concentration': 10.0}
    return dict(species=species, electrode_reactions=electrode_reactions,
                electrolyte_reactions=electrolyte_reactions, system=system, nx=nx)


def c1():
    return co2r_inputs()


def c2(n_potentials=1024, phi_min=-0.5, phi_max=-1.5):
    kw = co2r_inputs(i_CO=tafel_current(-10.), i_H2=tafel_current(-5.))
    kw['descriptors'] = {'phiM': list(np.linspace(phi_min, phi_max, n_potentials))}
    return kw


def c4(n_pH=256, n_L=256):
    kw = co2r_inputs(extra_species=True)
    kw['descriptors'] = {'bulk_pH': list(np.linspace(6.0, 7.83, n_pH)),
                         'boundary thickness': list(np.geomspace(10e-6, 200e-6, n_L))}
    return kw


def replicate_batch(batch, n_cells):
    """tile a CellBatch to n_cells (synthetic weak-scaling workloads)."""
    idx = np.arange(n_cells) % batch.B
    return batch.select(idx)


def geometric_mesh(n_nodes=1001, first_spacing=5e-11, L=50e-6):
    """normalised node positions xi in [0,1] of the graded mesh x_i = L*(r^i-1)/(r^(n-1)-1) whose first
    interval is `first_spacing` (0.05 nm = lambda_D/20 at the wall)."""
    from scipy.optimize import brentq
    m = n_nodes - 1
    f = lambda r: L * (r - 1.0) / (r ** m - 1.0) - first_spacing
    r = brentq(f, 1.0 + 1e-9, min(2.0, float(np.exp(600.0 / m))))     # r**m must not overflow
    i = np.arange(n_nodes, dtype=float)
    return (r ** i - 1.0) / (r ** m - 1.0)


def c3(n_phi=128, n_pH=128):
    """CO2R/KHCO3 with the Stern-layer (Robin) Poisson boundary on a 1001-node graded mesh,
    potential x pH sweep (use Calculator(..., poisson_bc='stern', mesh=geometric_mesh()))."""
    kw = co2r_inputs(i_CO=tafel_current(-10.), i_H2=tafel_current(-5.))
    kw['descriptors'] = {'phiM': list(np.linspace(-0.5, -1.5, n_phi)), 'bulk_pH': list(np.linspace(6.0, 7.8, n_pH))}
    return kw


def scaled_tafel_current(i_ref, phi_ref=-0.9, slope=0.12, clip=150.0):
    """Tafel current multiplied by the roughness factor system['RF'] (the reference's flux multiplier,
    /root/reference/catint/transport.py:844-851), clipped like tafel_current."""
    def current(system):
        i = i_ref * float(system.get('RF', 1.0)) * 10.0 ** (-(system['phiM'] - phi_ref) / slope)
        return float(np.sign(i_ref) * min(abs(i), clip))
    return current


def c5(n_phi=64, n_scale=64, nx=5000):
    """C5: CO2R/KHCO3 transient from the bulk state to steady state on a 5001-node mesh, 64 phiM x
    64 flux scalings (descriptor 'RF').  Use Calculator(..., mesh=geometric_mesh(5001, C5_FIRST_SPACING),
    mode='time-dependent') with an output time mesh; the default Poisson boundary (block size 9)."""
    kw = co2r_inputs(i_CO=scaled_tafel_current(-10.), i_H2=scaled_tafel_current(-5.), nx=nx)
    kw['descriptors'] = {'phiM': list(np.linspace(-0.5, -1.5, n_phi)), 'RF': list(np.geomspace(0.25, 4.0, n_scale))}
    return kw


C5_FIRST_SPACING = 5e-11
C5_T_OUT = [1e-6, 1e-3, 1.0, 200.0]


def c2_kinetic(n_potentials=1024, phi_min=-0.6, phi_max=-1.15, stern=False):
    """C2 with the wall kinetics evaluated on the surface state (flux equations) instead of host-side
    Tafel currents: CO2 reduction first order in the SURFACE CO2 concentration, both reactions with a transfer
    coefficient of 0.5 (118 mV/decade), -10 / -5 A/m^2 at phiM = -0.9 V and bulk composition.  With the Stern
    boundary (`stern=True`, use Calculator(..., poisson_bc='stern')) the driving force is the potential drop
    across the Stern layer, phiM - phi(0), as in the reference's docs (flux_definition.rst:108-118).
    Returns (Transport keywords, comsol_args)."""
    kw = co2r_inputs()
    drive = '(phiM-phi-phiEq_%s)' if stern else '(phiM-phiEq_%s)'
    kw['species']['CO'] = {'bulk_concentration': 0.0,
                           'flux-equation': 'k0_CO*[[CO2]]/conc_std*exp(-alpha*F_const*' + drive % 'CO' + '/RT)'}
    kw['species']['H2'] = {'bulk_concentration': 0.0,
                           'flux-equation': 'k0_H2*exp(-alpha*F_const*' + drive % 'H2' + '/RT)'}
    kw['descriptors'] = {'phiM': list(np.linspace(phi_min, phi_max, n_potentials))}
    # rate constants: with the Stern boundary the driving force is the (nearly potential-independent, ~ -0.1 V) drop
    # across the Stern layer, so the constants are chosen to give currents of the same order (-20 / -5 A/m^2)
    k0 = ('4.0e-6', '4.0e-6') if stern else ('3.2e-13', '6.4e-13')
    comsol_args = {'parameter': {'k0_CO': [k0[0] + '[mol/m^2/s]', 'CO2R rate constant'],
                                 'k0_H2': [k0[1] + '[mol/m^2/s]', 'HER rate constant'],
                                 'alpha': ['0.5', 'transfer coefficient'],
                                 'phiEq_CO': ['-0.11[V]', 'equilibrium potential CO2R'],
                                 'phiEq_H2': ['0.0[V]', 'equilibrium potential HER']}}
    return kw, comsol_args
