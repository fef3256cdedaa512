"""reference module path catint.data -> catint_b200.data"""
from catint_b200.data import *  # noqa: F401,F403
from catint_b200.data import tp_ref_data  # noqa: F401
