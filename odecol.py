"""Short alias: ``import odecol`` == the package in ``ode-column_b200/``."""
import sys
import ode_column_b200 as _pkg  # noqa: F401

sys.modules[__name__] = sys.modules["ode_column_b200"]
