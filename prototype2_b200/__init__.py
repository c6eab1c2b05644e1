"""prototype2_b200 -- B200-native (sm_100a) character-collision hot path of ArneStenkrona/prototype2.

The product is the C-ABI shared library `libprtc.so` (include/prtc.h) built from prototype2_b200/csrc; this
package is only its Python face (ctypes) plus the synthetic-scene generator used by the tests and bench.py.
There is no CPU fallback: importing `capi` raises if the library has not been built, and creating a context
raises if no compute-capability-10 device is present.
"""
from . import scenes  # noqa: F401

__all__ = ["scenes", "capi"]
