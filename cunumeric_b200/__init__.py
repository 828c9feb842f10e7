"""cunumeric_b200 — B200-native (sm_100a) hot path behind the cuNumeric NumPy API.

    import cunumeric_b200 as np

Elementwise ufuncs, where, astype and reductions run as hand-written CUDA kernels reached through
the C ABI in include/cunumeric_b200.h.  There is no CPU fallback: importing works anywhere, the
first array operation needs a B200."""
from numpy import (bool_, complex64, complex128, dtype, e, euler_gamma, float16, float32,  # noqa
                   float64, inf, int8, int16, int32, int64, nan, newaxis, pi, uint8, uint16,
                   uint32, uint64, finfo, iinfo, result_type, can_cast, broadcast_shapes)

from . import config  # noqa: F401
from ._ufunc import *  # noqa: F401,F403
from ._ufunc import ufunc  # noqa: F401
from .array import ndarray  # noqa: F401
from .module import *  # noqa: F401,F403
from .module import all, any, max, min, sum  # noqa: F401,A004
from ._ufunc.math import abs  # noqa: F401,A004
from .runtime import runtime  # noqa: F401
from .distributed import replicated  # noqa: F401

__version__ = "0.1.0"
