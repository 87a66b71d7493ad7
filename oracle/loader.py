"""Loader for the parity checker -- TEST INFRASTRUCTURE ONLY (see oracle/avif_oracle.c header).

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg.
"""
import ctypes
import os
import subprocess

from libavif_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libavif_ref.so")


def build(quiet=True):
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL if quiet else None)


class _Facade:
    """Exposes avif* names over a library whose symbols may carry another prefix (oracle*)."""

    def __init__(self, cdll, rename):
        self._cdll = cdll
        for name, (res, args) in abi.PROTOTYPES.items():
            sym = rename(name)
            try:
                fn = getattr(cdll, sym)
            except AttributeError:
                continue
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)
        self.raw = cdll


def load_oracle():
    if not os.path.exists(ORACLE_SO):
        build()
    lib = _Facade(ctypes.CDLL(ORACLE_SO), lambda n: "oracle" + n[4:] if n.startswith("avif") else n)
    raw = lib.raw
    raw.oracleXorshiftFill.restype = ctypes.c_uint32
    raw.oracleXorshiftFill.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int]
    raw.oracleFnv1a64.restype = ctypes.c_uint64
    raw.oracleFnv1a64.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
    raw.oracleCalcYUVCoefficients.restype = None
    raw.oracleCalcYUVCoefficients.argtypes = [ctypes.POINTER(abi.avifImage)] + [ctypes.POINTER(ctypes.c_float)] * 3
    raw.oracleImageLimitedToFullAlpha.restype = ctypes.c_int
    raw.oracleImageLimitedToFullAlpha.argtypes = [ctypes.POINTER(abi.avifImage)]
    return lib


def load_reference():
    """The unmodified reference compiled from /root/reference (None when it was never built here)."""
    if not os.path.exists(REF_SO):
        return None
    lib = _Facade(ctypes.CDLL(REF_SO), lambda n: n)
    raw = lib.raw
    raw.avifCalcYUVCoefficients.restype = None
    raw.avifCalcYUVCoefficients.argtypes = [ctypes.POINTER(abi.avifImage)] + [ctypes.POINTER(ctypes.c_float)] * 3
    return lib


REF_B2_SO = os.path.join(_HERE, "_ref", "libavif_ref_b2.so")


def load_reference_b2():
    """The same reference sources minus reformat_libyuv.c / reformat_libsharpyuv.c, linked against libavif_b200_hooks.so: the
    reference's own reformat.c / alpha.c driving the GPU through its internal seam (None when it was never built here)."""
    if not os.path.exists(REF_B2_SO):
        return None
    return _Facade(ctypes.CDLL(REF_B2_SO), lambda n: n)
