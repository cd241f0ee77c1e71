"""floria_b200 — B200-native implementation of floria's read-to-haplotype scoring / local clustering hot path.

The compute path is the CUDA library floria_b200/libfloria_b200.so behind the C-ABI of include/floria_b200.h;
this package is the thin ctypes host binding used by tests and bench.py.  There is no CPU fallback.
"""
from ._cdefs import default_params  # noqa: F401
from .frags import Frags  # noqa: F401
