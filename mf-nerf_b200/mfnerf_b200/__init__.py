"""mfnerf_b200 -- B200 (sm_100a) implementation of MF-NeRF's per-ray / per-sample hot path.

Host side mirrors the reference's operator interface (models/custom_functions.py, models/networks.py,
models/rendering.py, losses.py); device side is libmfnerf_b200.so behind the C ABI in include/mfnerf_b200.h.
"""
from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)

__version__ = "0.1.0"
