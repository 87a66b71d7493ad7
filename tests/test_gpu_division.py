"""The kernels replace divisions by call constants with q=x*r; e=fma(-q,d,x); q'=fma(e,r,q).  That is only legal where
it equals the IEEE quotient: this test re-proves it ON THE DEVICE for every whitelisted divisor, over all 2^32 inputs."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu


def test_every_whitelisted_divisor_exhaustively(gpu_product):
    n = gpu_product.avifB200VerifiedDivisorCount()
    assert n >= 50
    for k in range(n):
        bits = gpu_product.avifB200VerifiedDivisorBits(k)
        bad = ctypes.c_uint64(123)
        assert gpu_product.avifB200SelfTestDivisor(bits, ctypes.byref(bad)) == 0
        assert bad.value == 0, f"divisor 0x{bits:08x}: {bad.value} inputs differ from IEEE division"


@pytest.mark.parametrize("depth", [8, 10, 12, 16])
@pytest.mark.parametrize("sampled", [0, 1], ids=["all-code-pairs", "random-colour"])
def test_shared_reciprocal_alpha_division(gpu_product, depth, sampled):
    """Unmultiply shares one refined reciprocal between the three divisions by alpha: must equal IEEE division for every
    (colour code, alpha code) pair of each depth, and on random colour values for every alpha code."""
    bad = ctypes.c_uint64(1)
    assert gpu_product.avifB200SelfTestAlphaDivision(depth, sampled, ctypes.byref(bad)) == 0
    assert bad.value == 0


def test_integer_numerator_division_on_device(gpu_product):
    """div_int (2-op quotient for sample / channel codes): host proof and device re-check agree, for every divisor the path
    uses: code ranges of every depth x range, and 2^depth - 1."""
    import struct
    ds = [255, 1023, 4095, 65535] + [219 << k for k in (0, 2, 4, 8)] + [224 << k for k in (0, 2, 4, 8)]
    for d in ds:
        bits = struct.unpack("<I", struct.pack("<f", float(d)))[0]
        verdict, bad = ctypes.c_int(-1), ctypes.c_uint64(99)
        assert gpu_product.avifB200SelfTestIntDivisor(bits, ctypes.byref(verdict), ctypes.byref(bad)) == 0
        assert verdict.value == 1 and bad.value == 0, (d, verdict.value, bad.value)
    # host verdict and device count must agree for arbitrary divisors too (whatever the verdict is)
    for d in (3.0, 0.7152, 1000.0, 12345.678, 7.0):
        bits = struct.unpack("<I", struct.pack("<f", d))[0]
        verdict, bad = ctypes.c_int(-1), ctypes.c_uint64(99)
        assert gpu_product.avifB200SelfTestIntDivisor(bits, ctypes.byref(verdict), ctypes.byref(bad)) == 0
        assert (verdict.value == 1) == (bad.value == 0), (d, verdict.value, bad.value)


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_premultiply_fma_form_on_device(gpu_product, depth):
    """The 2-fma integer premultiply of the fused YUV->RGB kernels (verified_premul.h) against the reference's
    multiply / IEEE divide / add / floor sequence, on the device, for every (colour code, alpha code) pair."""
    bad = ctypes.c_uint64(5)
    assert gpu_product.avifB200SelfTestPremulFma(depth, ctypes.byref(bad)) == 0
    assert bad.value == 0
