"""Drop-in for the reference's Python package ``gtn`` (bindings/python/gtn/__init__.py:13-20),
restricted to the hot path: ``import gtn_b200.gtn as gtn``.  Everything comes from the
pybind11 module ``gtn_b200._gtn`` built on the C-ABI library; nothing computes on the CPU."""
from .._gtn import *  # noqa: F401,F403
from .._gtn import __version__, epsilon  # noqa: F401
