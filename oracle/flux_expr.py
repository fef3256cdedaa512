"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Flux equations (wall kinetics as expressions of the surface state): the reference hands the strings
``RF*flux_factor*(<species[sp]['flux-equation']>)`` to COMSOL with ``[[name]]`` replaced by the surface
concentration variable (/root/reference/catint/comsol_model.py:986-1000; semantics in
/root/reference/docs/source/topics/flux_definition.rst:100-156).  This module evaluates the same strings with
Python's own expression evaluator -- deliberately independent of the product's postfix compiler
(catint_b200/fluxeq.py) and of the device interpreter -- and provides exact derivatives by the complex-step
method for the oracle's Newton/BDF solvers.
"""
import re

import numpy as np

_FUNCS = dict(exp=np.exp, log=np.log, ln=np.log, sqrt=np.sqrt, log10=np.log10, tanh=np.tanh, pi=np.pi)


def _pythonise(expr, names):
    out, i = [], 0
    while i < len(expr):                                   # drop [unit] suffixes, keep [[species]]
        if expr.startswith('[[', i):
            j = expr.index(']]', i)
            out.append('c_[%d]' % names.index(expr[i + 2:j].strip()))
            i = j + 2
        elif expr[i] == '[':
            i = expr.index(']', i) + 1
        else:
            out.append(expr[i])
            i += 1
    return ''.join(out).replace('^', '**')


class WallKinetics(object):
    """J_add(c_wall, phi0) = coef @ [RF*flux_factor*(expr_e)]  with parameters `params` (name -> float) and
    variables `variables` (name -> expression string, may use parameters, [[species]] and earlier variables)."""

    def __init__(self, species_names, expressions, coef, params, variables=None):
        self.names = list(species_names)
        self.coef = np.asarray(coef, dtype=float)
        self.params = dict(params)
        self.params.setdefault('RF', 1.0)
        self.params.setdefault('flux_factor', 1.0)
        self.vars = [(k, compile(_pythonise(v, self.names), '<var %s>' % k, 'eval')) for k, v in (variables or {}).items()]
        self.codes = [compile(_pythonise('RF*flux_factor*(' + e + ')', self.names), '<flux>', 'eval') for e in expressions]

    def equations(self, c_wall, phi0):
        env = dict(_FUNCS)
        env.update(self.params)
        env['c_'] = c_wall
        env['phi'] = phi0
        for _ in range(len(self.vars) + 1):                # variables may refer to each other in any order
            pending = False
            for k, code in self.vars:
                try:
                    env[k] = eval(code, {'__builtins__': {}}, env)
                except NameError:
                    pending = True
            if not pending:
                break
        return np.array([eval(code, {'__builtins__': {}}, env) for code in self.codes])

    def __call__(self, c_wall, phi0):
        return self.coef @ self.equations(np.asarray(c_wall), phi0)

    def jacobian(self, c_wall, phi0):
        """(dJ/dc_wall [S,S], dJ/dphi0 [S]) by the complex-step method (exact to rounding)."""
        S = len(self.names)
        h = 1e-30
        dc = np.zeros((S, S))
        for j in range(S):
            cw = np.asarray(c_wall, dtype=complex).copy()
            cw[j] += 1j * h
            dc[:, j] = np.imag(self.coef @ self.equations(cw, phi0)) / h
        dphi = np.imag(self.coef @ self.equations(np.asarray(c_wall, dtype=complex), phi0 + 1j * h)) / h
        return dc, dphi
