"""tactilesr_b200: B200-native (sm_100a) implementation of the tactileSR hot path.

Drop-in module classes live in ``tactilesr_b200.model`` (same names / constructor arguments /
state_dict layout as the reference's ``model/*.py``); they call hand-written CUDA kernels through
the C ABI declared in ``include/tactilesr_b200.h``.  There is no CPU or PyTorch fallback.
"""
from .engine import check_fp16_overflow, get_precision, invalidate_packed_weights, set_precision  # noqa: F401
from ._lib import TsrError, launch_count  # noqa: F401

__all__ = ["set_precision", "get_precision", "check_fp16_overflow", "invalidate_packed_weights", "TsrError", "launch_count"]
