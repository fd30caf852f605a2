"""ccl_b200 -- B200-native Limber angular power spectra behind the pyccl API names.

Hot path: ``ccl_angular_cls_limber`` (reference src/ccl_cls.c:265-363) as hand-written sm_100a
CUDA kernels in ``libccl_b200.so``, reached through the C ABI of ``include/ccl_b200.h``.
"""
from . import _lib  # noqa: F401
from .lowlevel import (CCLB200Error, LibBatch, LibCosmology, LibF2D, LibTracer,  # noqa: F401
                       angular_cls_limber, make_params)

__version__ = "0.1.0"
