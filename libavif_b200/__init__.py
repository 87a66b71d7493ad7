"""libavif_b200 -- B200-native (sm_100a) implementation of libavif's pixel-reformat + alpha path.

The product is the C-ABI shared library ``libavif_b200/csrc/libavif_b200.so`` (hand-written CUDA kernels behind
the reference's own ``avifImageYUVToRGB`` / ``avifImageRGBToYUV`` / premultiply / alpha entry points, see
``include/avif_b200.h``).  This Python package is only the loader and a thin host-side mirror of the reference
interface used by the tests and the benchmark; it contains no pixel arithmetic and no CPU fallback.
"""
import ctypes
import os

from . import abi
from .abi import *  # noqa: F401,F403  (enum values, struct classes)

_HERE = os.path.dirname(os.path.abspath(__file__))
# AVIF_B200_LIB: load another build of the same sources (kernel-tuning variants, tools/variant_sweep.sh)
LIB_PATH = os.environ.get("AVIF_B200_LIB") or os.path.join(_HERE, "csrc", "libavif_b200.so")
_lib = None


class ExtensionMissing(RuntimeError):
    pass


def lib():
    """The C-ABI library.  Fails loudly when it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissing(
                f"{LIB_PATH} is missing: build it with `make -C libavif_b200/csrc` (or __graft_entry__.build()). "
                "libavif_b200 has no CPU or PyTorch fallback.")
        _lib = abi.bind(ctypes.CDLL(LIB_PATH))
    return _lib


def device_available():
    return bool(lib().avifB200DeviceAvailable())


from .images import YUVImage, RGBImage, pointer_array  # noqa: E402,F401
