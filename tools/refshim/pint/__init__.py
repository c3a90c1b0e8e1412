"""Unit-transparent stand-in for `pint`, used ONLY to import the reference planner
in the build container when generating golden vectors (tools/gen_golden.py).

`pint` is not installed here and there is no network.  The reference's solve path
only ever works with SI magnitudes (with the real pint the solve
would raise DimensionalityError), so a Quantity that *is* an ndarray and ignores its
unit string reproduces the only executable semantics.  Not part of the product.
"""
import numpy as np

from . import errors  # noqa: F401
from .errors import DimensionalityError  # noqa: F401


class Quantity(np.ndarray):
    def __new__(cls, value, units=None):
        obj = np.asarray(value, dtype=float).view(cls)
        obj._units = units
        return obj

    def __array_finalize__(self, obj):
        self._units = getattr(obj, "_units", None)

    @property
    def magnitude(self):
        a = np.asarray(self)
        return float(a) if a.ndim == 0 else a

    m = magnitude

    @property
    def units(self):
        return self._units

    def to(self, unit):
        return self

    def to_base_units(self):
        return self

    def m_as(self, unit):
        return self.magnitude

    def check(self, dim):
        return True

    # pint's Quantity is hashable (the reference uses Quantities as dataclass field defaults,
    # planning/global_mission_planner.py:45-55); an ndarray is not
    def __hash__(self):
        return id(self)


class UnitRegistry:
    Quantity = Quantity

    def setup_matplotlib(self, *a, **k):
        pass

    def __contains__(self, item):
        return True

    def define(self, *a, **k):
        pass
