"""multicameratracking_b200 -- B200-native appearance-scoring path of MultiCameraTracking.

Hand-written sm_100a CUDA behind a C-ABI (include/mct_b200.h, libmct_b200.so); `capi` is the ctypes
binding, `synth` the synthetic-sequence generator used by the tests and bench.py.  There is no CPU
fallback: importing works anywhere, computing needs a CUDA device.
"""
from . import build  # noqa: F401

__all__ = ["build", "capi", "synth", "host"]
