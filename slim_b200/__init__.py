"""slim_b200 -- Blackwell-native SLIM model learning behind the reference C ABI.

The product is ``slim_b200/lib/libslim.so`` (CUDA engine + C ABI, built by ``slim_b200.build``).
This package is the host-side mirror of the reference ``python-package/SLIM`` wrapper (same class
and method names) plus the column-sharded multi-GPU driver (``slim_b200.dist``).
"""
from .core import SLIM, SLIMatrix, Staged, learn_columns  # noqa: F401

__all__ = ["SLIM", "SLIMatrix", "Staged", "learn_columns"]
