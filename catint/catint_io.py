"""reference module path catint.catint_io -> catint_b200.catint_io"""
from catint_b200.catint_io import *  # noqa: F401,F403
