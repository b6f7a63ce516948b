"""vainplex_openclaw_b200 -- B200-native hot path of @vainplex/openclaw-governance.

(The task names the package `vainplex-openclaw_b200`; a hyphen is not importable in Python, so
the directory uses an underscore.)  Contents: csrc/ (sm_100a CUDA kernels + the C ABI of
include/openclaw_gov.h), _native.py (ctypes binding) and the host-side mirror of the reference's
plugin interface for this path (registry / engine / vault / conditions / hooks).
"""
from . import _native  # noqa: F401

__all__ = ["_native"]
__version__ = "0.1.0"
