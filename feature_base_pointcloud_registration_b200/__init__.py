"""B200-native scan-to-map registration hot path (LOAM / LIO-SAM) behind a C ABI.

The compute lives in ``libfbpr_b200.so`` (hand-written sm_100a CUDA, ``include/fbpr_b200.h``).
This package is the thin ctypes host layer used by the tests and ``bench.py``; the C++ host
classes that keep the reference's operator names are in ``host/``.  There is no CPU fallback:
importing :mod:`.api` without the built library, or creating a handle without a B200, raises.
"""
from .api import (  # noqa: F401
    FbprError, Params, Registration, load_library, library_path, RAW_POINT_DTYPE, RESULT_DTYPE,
    FLAG_NOT_ENOUGH_FEATURES, FLAG_TOO_FEW_CORRESPONDENCES, FLAG_DEGENERATE, FLAG_CONVERGED, FLAG_MAP_TRUNCATED, BUF,
)
from .params import load_params_yaml  # noqa: F401
