"""Flux equations: wall kinetics given as an expression of the local surface
concentrations and the local potential, compiled to a small postfix program the
CUDA kernels evaluate (value and derivatives) inside the wall rows.

Reference semantics (/root/reference/docs/source/topics/flux_definition.rst:100-156,
/root/reference/catint/comsol_model.py:962-1000, :1093-1175): a species may carry
``species[sp]['flux-equation'] = '<expression>'`` (mol m^-2 s^-1); ``[[name]]`` is the
surface concentration of species ``name``; identifiers are COMSOL parameters
(``comsol_args['parameter'][name] = [value, description]``, value a number or a string with
a ``[unit]`` suffix), COMSOL variables (``comsol_args['global_variables'|'boundary_variables'][name]
= [expression, description]``, expanded recursively) or the built-ins the reference defines for
every model: ``phiM, phiPZC, phi`` (potential at the reaction plane), ``T, RT, R_const, F_const, CS``
(F/m^2), ``eps_r, RF, flux_factor, PZC_factor, conc_std, lambdaD, L_cell``.  The flux the reference hands to
COMSOL is ``RF*flux_factor*(expression)`` (comsol_model.py:1000); fluxes of the other species of the
electrode reaction follow from the stoichiometry (transport.py:1057-1087), which is linear, so that
    J_k = J_k^fixed + sum_e coef[k][e] * E_e(c(0), phi(0))
with one E_e per flux equation.  In the reference only the COMSOL backend can evaluate these; here the
finite-difference kernels do.

Program encoding (include/catint_pnp.h, CatintPnpFluxEq): int32 words ``opcode | operand << 8``.
"""
import math
import re

OP_CONST, OP_PAR, OP_CONC, OP_PHI, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_NEG, OP_EXP, OP_LOG, OP_SQRT, \
    OP_LOG10, OP_TANH = range(15)
FUNCTIONS = {'exp': OP_EXP, 'log': OP_LOG, 'ln': OP_LOG, 'sqrt': OP_SQRT, 'log10': OP_LOG10, 'tanh': OP_TANH}
MAX_EQ, MAX_CODE, MAX_CONST, MAX_PAR, MAX_STACK = 4, 96, 32, 24, 16

_TOKEN = re.compile(r'\s*(?:(\[\[[^\]]+\]\])|(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)|([A-Za-z_][A-Za-z_0-9.]*)|'
                    r'(\*\*|[-+*/^(),]))')


class FluxEqError(ValueError):
    pass


def strip_units(text):
    """drop COMSOL unit suffixes ``[V]``, ``[mol/m^2]`` (single brackets), keep ``[[species]]``"""
    out, i = [], 0
    while i < len(text):
        if text.startswith('[[', i):
            j = text.index(']]', i)
            out.append(text[i:j + 2])
            i = j + 2
        elif text[i] == '[':
            j = text.index(']', i)
            i = j + 1
        else:
            out.append(text[i])
            i += 1
    return ''.join(out)


def tokenize(text):
    text = strip_units(text)
    pos, toks = 0, []
    while pos < len(text):
        if text[pos:].strip() == '':
            break
        m = _TOKEN.match(text, pos)
        if not m:
            raise FluxEqError('cannot parse flux equation at "%s"' % text[pos:pos + 20])
        conc, num, ident, op = m.groups()
        if conc:
            toks.append(('conc', conc[2:-2].strip()))
        elif num:
            toks.append(('num', float(num)))
        elif ident:
            toks.append(('id', ident))
        else:
            toks.append(('op', '^' if op == '**' else op))
        pos = m.end()
    return toks


_PREC = {'+': 1, '-': 1, '*': 2, '/': 2, 'u-': 3, '^': 4}
_RIGHT = {'^', 'u-'}


def to_postfix(toks, variables, depth=0):
    """shunting-yard; identifiers that name a variable are expanded in place (as a parenthesised
    sub-expression).  Returns a list of ('num', v) | ('conc', name) | ('id', name) | ('op', o) | ('fn', f)."""
    if depth > 16:
        raise FluxEqError('variables nested too deeply (circular definition?)')
    out, stack = [], []
    prev = None
    for kind, val in toks:
        if kind in ('num', 'conc'):
            out.append((kind, val))
        elif kind == 'id':
            if val in FUNCTIONS:
                stack.append(('fn', val))
            elif val in variables:
                out.extend(to_postfix(tokenize(variables[val]), variables, depth + 1))
            else:
                out.append(('id', val))
        elif val == ',':
            raise FluxEqError('functions of several arguments are not available in flux equations')
        elif val == '(':
            stack.append(('op', '('))
        elif val == ')':
            while stack and stack[-1] != ('op', '('):
                out.append(stack.pop())
            if not stack:
                raise FluxEqError('unbalanced parentheses in flux equation')
            stack.pop()
            if stack and stack[-1][0] == 'fn':
                out.append(stack.pop())
        else:
            o = val
            unary = prev is None or (prev[0] == 'op' and prev[1] != ')')
            if o == '+' and unary:
                prev = (kind, val)
                continue
            if o == '-' and unary:
                o = 'u-'
            while stack and stack[-1][0] == 'op' and stack[-1][1] != '(':
                top = stack[-1][1]
                if _PREC[top] > _PREC[o] or (_PREC[top] == _PREC[o] and o not in _RIGHT):
                    out.append(stack.pop())
                else:
                    break
            stack.append(('op', o))
        prev = (kind, val)
    while stack:
        t = stack.pop()
        if t == ('op', '('):
            raise FluxEqError('unbalanced parentheses in flux equation')
        out.append(t)
    return out


class FluxProgram(object):
    """one compiled expression: ``code`` (int32 words), ``consts``; parameter names are shared by all
    equations of a model (``FluxEquations.par_names``)."""

    def __init__(self, code, consts):
        self.code, self.consts = code, consts


class FluxEquations(object):
    """all flux equations of a model + the stoichiometric coefficients that distribute them."""

    def __init__(self, species_names):
        self.species = list(species_names)
        self.programs = []          # FluxProgram per equation
        self.owners = []            # species that carries the equation
        self.par_names = []         # per-cell parameters referenced by the programs
        self.coef = None            # [S][n_eq]

    @property
    def n_eq(self):
        return len(self.programs)

    def add(self, owner, expression, variables):
        post = to_postfix(tokenize(expression), variables)
        code, consts = [], []
        depth = maxdepth = 0
        for kind, val in post:
            if kind == 'num':
                if val not in consts:
                    consts.append(val)
                code.append(OP_CONST | consts.index(val) << 8)
                depth += 1
            elif kind == 'conc':
                if val not in self.species:
                    raise FluxEqError('flux equation refers to [[%s]], which is not a transported species' % val)
                code.append(OP_CONC | self.species.index(val) << 8)
                depth += 1
            elif kind == 'id':
                if val == 'phi':
                    code.append(OP_PHI)
                else:
                    if val not in self.par_names:
                        self.par_names.append(val)
                    code.append(OP_PAR | self.par_names.index(val) << 8)
                depth += 1
            elif kind == 'fn':
                code.append(FUNCTIONS[val])
            else:
                o = val
                if o == 'u-':
                    code.append(OP_NEG)
                else:
                    code.append({'+': OP_ADD, '-': OP_SUB, '*': OP_MUL, '/': OP_DIV, '^': OP_POW}[o])
                    depth -= 1
            if depth < 1:
                raise FluxEqError('malformed flux equation "%s"' % expression)
            maxdepth = max(maxdepth, depth)
        if depth != 1:
            raise FluxEqError('malformed flux equation "%s"' % expression)
        if len(code) > MAX_CODE or len(consts) > MAX_CONST or maxdepth > MAX_STACK:
            raise FluxEqError('flux equation too long for the device program (%d ops, %d constants, stack %d)'
                              % (len(code), len(consts), maxdepth))
        if len(self.par_names) > MAX_PAR:
            raise FluxEqError('too many parameters in the flux equations (max %d)' % MAX_PAR)
        if len(self.programs) >= MAX_EQ:
            raise FluxEqError('at most %d flux equations per model' % MAX_EQ)
        self.programs.append(FluxProgram(code, consts))
        self.owners.append(owner)

    # host-side evaluation of the compiled program (tests; the product evaluates on the device)
    def evaluate(self, e, conc, phi, par):
        prog = self.programs[e]
        st = []
        for w in prog.code:
            op, arg = w & 0xff, w >> 8
            if op == OP_CONST:
                st.append(prog.consts[arg])
            elif op == OP_PAR:
                st.append(par[arg])
            elif op == OP_CONC:
                st.append(conc[arg])
            elif op == OP_PHI:
                st.append(phi)
            elif op == OP_NEG:
                st.append(-st.pop())
            elif op in (OP_EXP, OP_LOG, OP_SQRT, OP_LOG10, OP_TANH):
                a = st.pop()
                st.append({OP_EXP: math.exp, OP_LOG: math.log, OP_SQRT: math.sqrt, OP_LOG10: math.log10,
                           OP_TANH: math.tanh}[op](a))
            else:
                b = st.pop()
                a = st.pop()
                st.append(a + b if op == OP_ADD else a - b if op == OP_SUB else a * b if op == OP_MUL
                          else a / b if op == OP_DIV else a ** b)
        return st[0]


def parameter_values(par_names, system, user_params, model):
    """numeric value of every referenced identifier for one cell (one set of system values)."""
    from .units import unit_R, unit_F
    T = float(system['temperature'])
    builtin = {
        'phiM': float(system['phiM']), 'phiPZC': float(system['phiPZC']), 'T': T, 'R_const': unit_R,
        'RT': unit_R * T, 'F_const': unit_F, 'CS': float(system['Stern capacitance']) * 1e-2,
        'eps_r': float(system['epsilon']), 'RF': float(system.get('RF', 1.0)), 'flux_factor': 1.0,
        'PZC_factor': 1.0, 'conc_std': 1.0, 'lambdaD': float(model.debye_length), 'L_cell': float(model.xmax),
        'pi': math.pi,
    }
    vals = []
    for name in par_names:
        if name in user_params:
            v = user_params[name]
            if isinstance(v, (list, tuple)):
                v = v[0]
            if isinstance(v, str):
                v = float(strip_units(v))
            vals.append(float(v))
        elif name in builtin:
            vals.append(builtin[name])
        elif name in system and isinstance(system[name], (int, float)):
            vals.append(float(system[name]))
        else:
            raise FluxEqError('unknown identifier "%s" in a flux equation (define it in comsol_args["parameter"])' % name)
    return vals
