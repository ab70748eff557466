"""recur_b200 — Python host-side view of librecur_b200.so.

The product is the C-ABI shared library built from recur_b200/csrc (C host
code + sm_100a CUDA kernels) that serves recur's RecurNN API (include/
recur-nn.h) and the array-of-nets calls (include/recur_b200.h).  This package
only binds it with ctypes, mirroring the C names one to one, so that tests
and bench.py read like callers of the reference.  Nothing here computes.
"""
from .api import lib, load_library, LIB_PATH  # noqa: F401
from . import abi  # noqa: F401
