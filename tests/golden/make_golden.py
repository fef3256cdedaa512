"""
Generates the committed golden fixtures under tests/golden/ .

Run ONLY in the build container (needs /root/reference, which does not exist
on the GPU box):      python tests/golden/make_golden.py [--skip-odeint]

Two kinds of fixtures:

1. ``ref_*.npz``  -- outputs of the REFERENCE's own code, executed here:
   * setup arrays of ``catint.transport.Transport`` (imported from
     /root/reference with a shim for the removed stdlib module ``imp``);
   * values of the reference's own RHS ``ode_func``
     (/root/reference/catint/calculator_old.py:827-935).  That file is
     orphaned Python 2 with no caller; its source is read from where it lies,
     two ``print`` statements and four absolute imports are rewritten IN
     MEMORY (nothing is copied into the repo), the module is executed, and
     ``scipy.integrate.odeint`` is intercepted to capture the RHS callable the
     reference would have integrated.  Attribute names of the older Transport
     it expects are mapped (use_reactions, reactions[r]['reactants']).
   These pin oracle/pnp_oracle.py (legacy_overwrite rates, J=-flux_bound).

2. ``oracle_*.npz`` -- outputs of oracle/pnp_oracle.py (scipy odeint, dense
   Jacobian, default rtol/atol, t_end=200 s) and of oracle/pnp_local.py for
   the named configurations; these are what the CUDA path is compared to on
   the GPU box.
"""
import copy
import importlib.util
import os
import re
import sys
import tempfile
import types
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)


def _install_imp_shim():
    imp = types.ModuleType('imp')

    def find_module(name):
        if importlib.util.find_spec(name) is None:
            raise ImportError(name)
        return name
    imp.find_module = find_module
    sys.modules['imp'] = imp


def _reference_modules():
    _install_imp_shim()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # our own repo also has a top-level 'catint' shim package: make sure the
    # reference one wins inside this generator
    for m in [m for m in sys.modules if m == 'catint' or m.startswith('catint.')]:
        del sys.modules[m]
    sys.path.remove(REF)
    sys.path.insert(0, REF)
    import catint.transport as rt
    import catint.data as rd
    assert rt.__file__.startswith(REF), rt.__file__
    return rt, rd


def c1_inputs(pH=6.8, L=50e-6, i_CO=-10., i_H2=-5., migration=True, rx=True, extra=None,
              temperature=298., phiM=-0.9):
    """SURVEY 8(d) common chemistry; fresh dict objects on every call
    (the reference mutates them, SURVEY C-7)."""
    system = {'temperature': temperature, 'pressure': 1.013, 'bulk_pH': pH, 'boundary thickness': L,
              'epsilon': 78.36, 'migration': migration, 'electrode reactions': True,
              'electrolyte reactions': rx, 'phiM': phiM, 'phiPZC': 0.16, 'Stern capacitance': 20.}
    electrolyte_reactions = ['bicarbonate-base', 'water-diss',
                             {'additional_cell_reactions': 'bicarbonate-acid'}] if rx else None
    electrode_reactions = {'CO': {'reaction': 'CO2 + H2O + 2 e- -> CO + 2 OH-'},
                           'H2': {'reaction': '2 H2O + 2 e- -> H2 + 2 OH-'}}
    species = {'K+': {'bulk_concentration': 'charge_neutrality'},
               'CO2': {'bulk_concentration': 'Henry'},
               'OH-': {'bulk_concentration': 10 ** (pH - 14.) * 1000.},
               'CO': {'bulk_concentration': 0.0, 'current density': i_CO},
               'H2': {'bulk_concentration': 0.0, 'current density': i_H2}}
    if extra == 'c4':
        electrode_reactions['CH4'] = {'reaction': 'CO2 + 6 H2O + 8 e- -> CH4 + 8 OH-'}
        species['CH4'] = {'bulk_concentration': 0.0, 'current density': -2.0}
        species['Cl-'] = {'bulk_concentration': 10.0}
    return dict(species=species, electrode_reactions=electrode_reactions,
                electrolyte_reactions=electrolyte_reactions, system=system, nx=100)


def build_reference_transport(rt, rd, kwargs, workdir):
    saved = copy.deepcopy(rd.tp_ref_data)
    cwd = os.getcwd()
    os.chdir(workdir)
    open(os.path.join(workdir, 'input.py'), 'w').write('# golden generator\n')
    argv0 = sys.argv[0]
    sys.argv[0] = 'input.py'
    try:
        tp = rt.Transport(comsol_args={'bin_version': 5.3}, catint_path=REF, **kwargs)
    finally:
        sys.argv[0] = argv0
        os.chdir(cwd)
        # the reference rewrites its module-level reaction library (SURVEY C-7)
        live = copy.deepcopy(tp_reactions(tp)) if 'tp' in locals() else None
        rd.tp_ref_data.clear()
        rd.tp_ref_data.update(saved)
    tp._parsed_reactions = live
    return tp


def tp_reactions(tp):
    if getattr(tp, 'electrolyte_reactions', None) is None or not tp.use_electrolyte_reactions:
        return {}
    return tp.electrolyte_reactions


def setup_arrays(tp):
    names = list(tp.species.keys())
    rx = tp._parsed_reactions or {}
    rlist = []
    for r in rx:
        if 'rates' not in rx[r]:
            continue
        ed = [names.index(s) for s in rx[r]['reaction'][0] if s in names]
        pr = [names.index(s) for s in rx[r]['reaction'][1] if s in names]
        rlist.append((r, ed, pr, rx[r]['rates'][0], rx[r]['rates'][1]))
    out = dict(
        species=np.array(names),
        z=np.array([tp.species[s]['charge'] for s in names], dtype=np.int64),
        charges=np.asarray(tp.charges, dtype=float),
        D=np.asarray(tp.D, dtype=float),
        c_bulk=np.array([tp.species[s]['bulk_concentration'] for s in names], dtype=float),
        flux=np.array([tp.species[s]['flux'] for s in names], dtype=float),
        flux_bound=np.asarray(tp.flux_bound, dtype=float),
        c0=np.asarray(tp.c0, dtype=float),
        xmesh=np.asarray(tp.xmesh, dtype=float),
        dx=float(tp.dx), nx=int(tp.nx), xmax=float(tp.xmax),
        beta=float(tp.beta), eps=float(tp.eps),
        debye_length=float(tp.debye_length), ionic_strength=float(tp.ionic_strength),
        bulk_pH=float(tp.system['bulk_pH']),
        phi_wall=float(tp.pb_bound['potential']['wall']),
        g_bulk=float(tp.pb_bound['gradient']['bulk']),
        use_migration=bool(tp.use_migration),
        rx_names=np.array([r[0] for r in rlist]),
        rx_educts=np.array([','.join(map(str, r[1])) for r in rlist]),
        rx_products=np.array([','.join(map(str, r[2])) for r in rlist]),
        rx_kf=np.array([r[3] for r in rlist], dtype=float),
        rx_kr=np.array([r[4] for r in rlist], dtype=float),
    )
    return out


def capture_reference_rhs(tp):
    """returns f(c)->dc/dt evaluated by the reference's own ode_func."""
    src = open(os.path.join(REF, 'catint', 'calculator_old.py')).read()
    src = re.sub(r"^(\s*)print '(\w+)', (.*)$", r"\1print('\2', \3)", src, flags=re.M)
    src = src.replace('from units import *', 'from catint.units import *')
    src = src.replace('from io import sync_mpi,reduce_dict_mpi', '')
    src = src.replace('from comsol_wrapper import Comsol', 'Comsol=None')
    src = src.replace('from catmap_wrapper import CatMAP', 'CatMAP=None')
    mod = types.ModuleType('reference_calculator_old')
    exec(compile(src, os.path.join(REF, 'catint', 'calculator_old.py'), 'exec'), mod.__dict__)

    # attribute names of the older Transport this file was written against
    tp.use_reactions = bool(tp.use_electrolyte_reactions and tp._parsed_reactions)
    tp.reactions = {}
    for r, v in (tp._parsed_reactions or {}).items():
        d = {'reactants': v['reaction']}
        if 'rates' in v:
            d['rates'] = v['rates']
        tp.reactions[r] = d
    tp.calc = 'odeint'
    if not hasattr(tp, 'use_catmap'):
        tp.use_catmap = False
    calc = mod.Calculator(transport=tp, dt=0.1, tmax=10., ntout=1, calc='odeint', desc_method='external')

    class _Captured(Exception):
        pass
    box = {}

    class _FakeIntegrate(object):
        @staticmethod
        def odeint(func, y0, t, args=(), **kw):
            box['func'] = func
            box['args'] = args
            box['y0'] = np.array(y0, dtype=float)
            box['kw'] = kw
            raise _Captured()
    mod.integrate = _FakeIntegrate
    try:
        calc.integrate_pnp(tp.dx, tp.nx, tp.dt, tp.nt, tp.ntout, 'odeint')
    except _Captured:
        pass
    func, args = box['func'], box['args']
    assert box['kw'].get('ml') == tp.nspecies and box['kw'].get('mu') == tp.nspecies

    def f(c):
        import logging
        lvl = logging.getLogger().level
        logging.disable(logging.CRITICAL)
        try:
            return np.array(func(np.array(c, dtype=float), 0.0, *args), dtype=float)
        finally:
            logging.disable(logging.NOTSET)
            logging.getLogger().setLevel(lvl)
    return f, box['y0'], tp


def oracle_system_from_setup(su, rate_mode, literal_sign):
    from oracle.fixtures import system_from_setup
    return system_from_setup(su, rate_mode, literal_sign)


def main():
    skip_odeint = '--skip-odeint' in sys.argv
    rt, rd = _reference_modules()
    rng = np.random.default_rng(0)
    work = tempfile.mkdtemp(prefix='catint_golden_')

    # ---------------- 1. reference-generated fixtures -----------------
    setup_cases = {
        'c1': c1_inputs(),
        'c1_nomig': c1_inputs(migration=False),
        'c1_norx': c1_inputs(rx=False),
        'c1_pH7p5': c1_inputs(pH=7.5, i_CO=-30., i_H2=-12.),
        'c1_L30': c1_inputs(L=30e-6),          # ragged nx: 102 nodes (SURVEY C-6)
        'c1_T320': c1_inputs(temperature=320.),
        'c4': c1_inputs(extra='c4'),
    }
    for name, kw in setup_cases.items():
        tp = build_reference_transport(rt, rd, kw, work)
        su = setup_arrays(tp)
        out = dict(su)
        if name in ('c1', 'c1_nomig', 'c1_norx', 'c4', 'c1_L30'):
            f, y0, _ = capture_reference_rhs(tp)
            assert np.array_equal(y0, su['c0'])
            S, n = len(su['z']), su['nx']
            states = [su['c0'].copy()]
            for amp in (1e-3, 0.1, 0.5):
                st = su['c0'] * (1.0 + amp * rng.standard_normal(S * n))
                st += amp * 1e-3 * rng.standard_normal(S * n)
                states.append(st)
            # a smooth boundary-layer-like state: species deplete/accumulate towards the wall
            xi = np.tile(su['xmesh'] / su['xmesh'][-1], S)
            prof = su['c0'] * (1.0 + 0.3 * np.sin(3.0 * xi) * (1 - xi)) + 0.05 * (1 - xi) ** 2
            states.append(prof)
            states = np.array(states)
            t0 = time.time()
            ref_rhs = np.array([f(s) for s in states])
            out['rhs_states'] = states
            out['rhs_ref'] = ref_rhs
            print('  reference ode_func evaluated on %d states in %.1fs' % (len(states), time.time() - t0))
        np.savez_compressed(os.path.join(HERE, 'ref_%s.npz' % name), **out)
        print('wrote ref_%s.npz  species=%s nx=%d' % (name, list(su['species']), su['nx']))

    # ---------------- 2. oracle-generated fixtures ---------------------
    if skip_odeint:
        return
    from oracle.pnp_oracle import steady_tmesh
    from oracle.pnp_local import LocalForm
    jobs = [('c1', 'summed', False), ('c1', 'legacy_overwrite', True), ('c1_nomig', 'summed', False),
            ('c1_norx', 'summed', False), ('c1_pH7p5', 'summed', False)]
    for name, mode, literal in jobs:
        su = dict(np.load(os.path.join(HERE, 'ref_%s.npz' % name)))
        sys_ = oracle_system_from_setup(su, mode, literal)
        tmesh = steady_tmesh()
        t0 = time.time()
        sol, info = sys_.integrate_odeint(tmesh, full_output=True)
        wall = time.time() - t0
        C = sol[-1].reshape(sys_.S, sys_.n)
        v, g, lapl = sys_.poisson(C) if sys_.use_migration else (np.zeros(sys_.n),) * 3
        # Newton root of the same discrete residual, started from the odeint end state
        # (the discrete steady problem has several roots, only the one odeint reaches counts)
        lf = LocalForm(sys_)
        y, inf2 = lf.solve_steady(y0=lf.y_from_c(C), pure_newton=True)
        Cn, vn, gn = lf.unpack(y)
        tag = '%s_%s' % (name, mode)
        np.savez_compressed(os.path.join(HERE, 'oracle_%s.npz' % tag),
                            tmesh=tmesh, c_end=sol[-1], c_t10=sol[np.searchsorted(tmesh, 10.0)],
                            potential=v, grad=g, lapl=lapl, nfe=int(info['nfe'][-1]), wall_s=wall,
                            newton_c=Cn.reshape(-1), newton_potential=vn, newton_grad=gn,
                            J=sys_.J, rate_mode=mode, literal_sign=literal)
        rel = np.max(np.abs(Cn - C) / (np.abs(C) + 1e-12 * np.max(sys_.c_bulk)))
        print('wrote oracle_%s.npz  nfe=%d  %.1fs  newton-vs-odeint rel=%.2e (newton steps %d, conv %s)'
              % (tag, info['nfe'][-1], wall, rel, inf2['steps'], inf2['converged']))


if __name__ == '__main__':
    main()
