"""Physical constants and energy-unit conversion.

Values are those of the reference (/root/reference/catint/units.py:4-27); the
FD-PNP path only needs unit_R, unit_F, unit_eps0, unit_NA and unit_T but the
constants must agree to the last digit for 1e-6 parity, so the whole table is
kept and ``convert_unit`` (units.py:31-91) is provided for user scripts.
"""
import sys

unit_R = 8.3144598            # J/mol/K
unit_e = 1.6021766208e-19     # C
eVToJ = unit_e
unit_eps0 = 8.854187817e-12   # F/m
unit_NA = 6.022140857e23      # 1/mol
calToJ = 4.182
unit_F = 96485.33289          # C/mol
JTokcal = 1e-3 / calToJ * unit_NA
BohrToAA = 0.52917721
HaToeV = 27.21138602
eVTokcal = 23.0609
eVTokcal2 = eVToJ * JTokcal
HaTokcal = 627.5095
HaToJ = 4.359744650e-18
unit_kB = 1.38064852e-23
unit_T = 298.14
unit_h = 6.626070040e-34
unit_c = 299792458
Rydberg = 0.5 * HaToeV

# energy-unit graph: every unit expressed in kJ/mol at the reference
# temperature unit_T, using the same factors the reference multiplies by
_kT_kJ = 0.001 * unit_R * unit_T


def _table():
    t = {}
    t[('kcal', 'eV')] = 1. / eVTokcal
    t[('kcal', 'kJ')] = calToJ
    t[('kcal', 'kT')] = 1000. / unit_R / unit_T * calToJ
    t[('eV', 'kcal')] = eVTokcal
    t[('eV', 'meV')] = 1000.
    t[('eV', 'Ha')] = 1. / HaToeV
    t[('eV', 'kT')] = eVTokcal * calToJ * 1000. / unit_R / unit_T
    t[('eV', 'kJ')] = eVTokcal * calToJ
    t[('Ha', 'eV')] = HaToeV
    t[('Ha', 'meV')] = HaToeV * 1000.
    t[('Ha', 'kJ')] = HaTokcal * calToJ
    t[('Ha', 'kcal')] = HaTokcal
    t[('Ha', 'kT')] = HaTokcal * calToJ * 1000. / unit_R / unit_T
    t[('kJ', 'kcal')] = 1. / calToJ
    t[('kJ', 'kT')] = 1000. / unit_R / unit_T
    t[('kT', 'kJ')] = 1. / 1000. * unit_R * unit_T
    t[('kT', 'kcal')] = 0.001 * unit_R * unit_T / calToJ
    t[('kT', 'eV')] = 0.001 * unit_R * unit_T / calToJ / eVTokcal
    t[('J', 'Ha')] = 1. / HaToJ
    t[('J', 'eV')] = 1. / eVToJ
    t[('J', 'kcal')] = 1. / JTokcal
    return t


_FACTORS = _table()


def convert_unit(val, start, end):
    """val [start] -> val [end]; same conversion pairs as the reference."""
    if (start, end) in _FACTORS:
        return val * _FACTORS[(start, end)]
    if start == end:
        return val * 1.
    print('unexpected conversion units, check units module for available conversions')
    sys.exit()
