"""reference module path catint.calculator -> catint_b200.calculator"""
from catint_b200.calculator import *  # noqa: F401,F403
