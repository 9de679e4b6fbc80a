"""diverseseq_b200 — B200 (sm_100a) implementation of diverse-seq's data-parallel hot path.

`_dvs` mirrors the reference's PyO3 module `diverse_seq._dvs`; `distance` mirrors the pair-matrix
functions of `diverse_seq/distance.py`; `_lib` is the ctypes binding of the C ABI
(include/dvs_b200.h, libdvs_b200.so).  Nothing here imports the CPU oracle.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
