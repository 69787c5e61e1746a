"""gtn_b200 -- B200-native (sm_100a) implementation of gtn's CTC/ASG hot path.

The product is the C-ABI library ``gtn_b200/lib/libgtn_b200.so`` declared in
``include/gtn_b200.h``; ``gtn_b200.capi`` is its ctypes binding.  Nothing in
this package computes on the CPU.
"""
from . import capi  # noqa: F401

__version__ = "0.1.0"
