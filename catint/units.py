"""reference module path catint.units -> catint_b200.units"""
from catint_b200.units import *  # noqa: F401,F403
