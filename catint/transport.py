"""reference module path catint.transport -> catint_b200.transport"""
from catint_b200.transport import *  # noqa: F401,F403
