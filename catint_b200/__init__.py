"""catint_b200 -- B200-native batched solver backend for CatINT's 1D
finite-difference Poisson-Nernst-Planck transport path.

Public surface = the reference's: ``Transport`` (catint/transport.py) and
``Calculator`` (catint/calculator.py).  See DESIGN.md.
"""
from .transport import Transport
from .calculator import Calculator

__all__ = ['Transport', 'Calculator']
