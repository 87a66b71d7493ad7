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


def test_two_operation_division_on_device(gpu_product):
    """fma(x, rh, RN(x*rl)) for arbitrary float numerators (tools/verify_quant.c `div`): every whitelisted divisor over all 2^32
    inputs on the device; the host's verdict (AbDiv.exact2) must be exactly "no input differs"."""
    n = gpu_product.avifB200VerifiedDivisorCount()
    refused = 0
    for k in range(n):
        bits = gpu_product.avifB200VerifiedDivisorBits(k)
        verdict, bad = ctypes.c_int(-1), ctypes.c_uint64(77)
        assert gpu_product.avifB200SelfTestDivisor2(bits, ctypes.byref(verdict), ctypes.byref(bad)) == 0
        assert (verdict.value == 1) == (bad.value == 0), f"divisor 0x{bits:08x}: host verdict {verdict.value}, {bad.value} inputs differ"
        refused += verdict.value == 0
    assert refused == 1  # 0x3fee5bb7


def test_merged_rounding_constant_through_the_kernels(gpu_product, oracle):
    """floor(p + (bias + 0.5)) for floor((p + bias) + 0.5) in the RGB->YUV quantiser (tools/verify_quant.c `quant`): every channel code
    triple of an 8-bit RGB cube slice through the specialised kernel against the oracle, limited and full range."""
    import numpy as np
    from libavif_b200 import abi
    from tests import util
    for rng in (0, 1):
        for depth in (8, 10, 12):
            spec = util.Spec(w=4096, h=64, depth=depth, fmt=abi.AVIF_PIXEL_FORMAT_YUV444, rng=rng, mc=abi.MC_BT709, rgb_depth=8, rgb_fmt=abi.AVIF_RGB_FORMAT_RGB)
            outs = []
            for lib in (oracle, gpu_product):
                rgb = util.make_rgb_src(spec, 1)
                r, g = np.meshgrid(np.arange(64) * 4 + 1, np.arange(4096) % 256, indexing="ij")
                b = (np.arange(4096)[None, :] // 256 * 17) % 256 + 0 * r
                rgb.set_payload(np.stack([r, g, b], axis=-1).astype(np.uint8).reshape(64, -1))
                img = util.make_yuv_dst(spec)
                assert lib.avifImageRGBToYUV(img.ptr(), rgb.ptr()) == 0
                outs.append(img.payload_bytes().copy())
            assert util.first_mismatch(outs[0], outs[1]) is None, (rng, depth)
