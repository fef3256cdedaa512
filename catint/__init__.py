"""``catint`` import path of the reference (user scripts do
``from catint.transport import Transport`` / ``from catint.calculator import
Calculator``, /root/reference/examples/02_CO2R_Au_CatMAP/run.py:2-3), served by
the catint_b200 package."""
from catint_b200 import Transport, Calculator
