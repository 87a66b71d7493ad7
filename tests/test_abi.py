"""Boundary checks that need no GPU: the C-ABI library loads, exports every symbol include/avif_b200.h declares, and its
struct layouts equal the reference's public header (compiled against /root/reference when present, else against the
committed snapshot tests/golden/abi_layout.json)."""
import json
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "avif_b200.h")
REF_INCLUDE = "/root/reference/include"
SNAPSHOT = os.path.join(ROOT, "tests", "golden", "abi_layout.json")

PROBE = r"""
#include <stdio.h>
#include <stddef.h>
%s
#define F(T, m) printf("%%s.%%s %%zu\n", #T, #m, offsetof(T, m))
int main(void) {
    printf("sizeof.avifImage %%zu\n", sizeof(avifImage));
    printf("sizeof.avifRGBImage %%zu\n", sizeof(avifRGBImage));
    F(avifImage, width); F(avifImage, height); F(avifImage, depth); F(avifImage, yuvFormat); F(avifImage, yuvRange);
    F(avifImage, yuvChromaSamplePosition); F(avifImage, yuvPlanes); F(avifImage, yuvRowBytes); F(avifImage, imageOwnsYUVPlanes);
    F(avifImage, alphaPlane); F(avifImage, alphaRowBytes); F(avifImage, imageOwnsAlphaPlane); F(avifImage, alphaPremultiplied);
    F(avifImage, icc); F(avifImage, colorPrimaries); F(avifImage, transferCharacteristics); F(avifImage, matrixCoefficients);
    F(avifImage, clli); F(avifImage, transformFlags); F(avifImage, pasp); F(avifImage, clap); F(avifImage, irot); F(avifImage, imir);
    F(avifImage, exif); F(avifImage, xmp); F(avifImage, properties); F(avifImage, numProperties); F(avifImage, gainMap);
    F(avifRGBImage, width); F(avifRGBImage, height); F(avifRGBImage, depth); F(avifRGBImage, format); F(avifRGBImage, chromaUpsampling);
    F(avifRGBImage, chromaDownsampling); F(avifRGBImage, avoidLibYUV); F(avifRGBImage, ignoreAlpha); F(avifRGBImage, alphaPremultiplied);
    F(avifRGBImage, isFloat); F(avifRGBImage, maxThreads); F(avifRGBImage, pixels); F(avifRGBImage, rowBytes);
    printf("sizeof.avifGainMap %%zu\n", sizeof(avifGainMap));
    printf("sizeof.avifDiagnostics %%zu\n", sizeof(avifDiagnostics));
    printf("sizeof.avifContentLightLevelInformationBox %%zu\n", sizeof(avifContentLightLevelInformationBox));
    F(avifGainMap, image); F(avifGainMap, gainMapMin); F(avifGainMap, gainMapMax); F(avifGainMap, gainMapGamma); F(avifGainMap, baseOffset);
    F(avifGainMap, alternateOffset); F(avifGainMap, baseHdrHeadroom); F(avifGainMap, alternateHdrHeadroom); F(avifGainMap, useBaseColorSpace);
    F(avifGainMap, altICC); F(avifGainMap, altColorPrimaries); F(avifGainMap, altTransferCharacteristics); F(avifGainMap, altMatrixCoefficients);
    F(avifGainMap, altYUVRange); F(avifGainMap, altDepth); F(avifGainMap, altPlaneCount); F(avifGainMap, altCLLI);
    printf("enum.AVIF_RESULT_INVALID_TONE_MAPPED_IMAGE %%d\n", (int)AVIF_RESULT_INVALID_TONE_MAPPED_IMAGE);
    printf("enum.AVIF_TRANSFER_CHARACTERISTICS_HLG %%d\n", (int)AVIF_TRANSFER_CHARACTERISTICS_HLG);
    printf("enum.AVIF_TRANSFER_CHARACTERISTICS_PQ %%d\n", (int)AVIF_TRANSFER_CHARACTERISTICS_PQ);
    printf("enum.AVIF_TRANSFER_CHARACTERISTICS_SRGB %%d\n", (int)AVIF_TRANSFER_CHARACTERISTICS_SRGB);
    printf("enum.AVIF_COLOR_PRIMARIES_SMPTE432 %%d\n", (int)AVIF_COLOR_PRIMARIES_SMPTE432);
    printf("enum.AVIF_RESULT_REFORMAT_FAILED %%d\n", (int)AVIF_RESULT_REFORMAT_FAILED);
    printf("enum.AVIF_RESULT_INVALID_ARGUMENT %%d\n", (int)AVIF_RESULT_INVALID_ARGUMENT);
    printf("enum.AVIF_RESULT_NOT_IMPLEMENTED %%d\n", (int)AVIF_RESULT_NOT_IMPLEMENTED);
    printf("enum.AVIF_RESULT_OUT_OF_MEMORY %%d\n", (int)AVIF_RESULT_OUT_OF_MEMORY);
    printf("enum.AVIF_PIXEL_FORMAT_YUV400 %%d\n", (int)AVIF_PIXEL_FORMAT_YUV400);
    printf("enum.AVIF_RGB_FORMAT_AGRAY %%d\n", (int)AVIF_RGB_FORMAT_AGRAY);
    printf("enum.AVIF_RGB_FORMAT_RGB_565 %%d\n", (int)AVIF_RGB_FORMAT_RGB_565);
    printf("enum.AVIF_CHROMA_UPSAMPLING_BILINEAR %%d\n", (int)AVIF_CHROMA_UPSAMPLING_BILINEAR);
    printf("enum.AVIF_CHROMA_DOWNSAMPLING_SHARP_YUV %%d\n", (int)AVIF_CHROMA_DOWNSAMPLING_SHARP_YUV);
    printf("enum.AVIF_MATRIX_COEFFICIENTS_YCGCO_RO %%d\n", (int)AVIF_MATRIX_COEFFICIENTS_YCGCO_RO);
    printf("enum.AVIF_RANGE_FULL %%d\n", (int)AVIF_RANGE_FULL);
    return 0;
}
"""


def _probe(include_line, include_dirs):
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        exe = os.path.join(td, "p")
        open(src, "w").write(PROBE % include_line)
        subprocess.run(["gcc", "-std=gnu11", "-o", exe, src] + [f"-I{d}" for d in include_dirs], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return dict(line.split() for line in out.strip().splitlines())


def test_layout_matches_reference_header():
    mine = _probe('#include "avif_b200.h"', [os.path.join(ROOT, "include")])
    if os.path.isdir(REF_INCLUDE):
        ref = _probe('#include "avif/avif.h"', [REF_INCLUDE])
        assert mine == ref
        os.makedirs(os.path.dirname(SNAPSHOT), exist_ok=True)
        if not os.path.exists(SNAPSHOT):
            json.dump(ref, open(SNAPSHOT, "w"), indent=1, sort_keys=True)
    snap = json.load(open(SNAPSHOT))
    assert mine == snap


def test_ctypes_mirror_matches_header():
    import ctypes
    from libavif_b200 import abi
    mine = _probe('#include "avif_b200.h"', [os.path.join(ROOT, "include")])
    assert ctypes.sizeof(abi.avifImage) == int(mine["sizeof.avifImage"])
    assert ctypes.sizeof(abi.avifRGBImage) == int(mine["sizeof.avifRGBImage"])
    assert ctypes.sizeof(abi.avifGainMap) == int(mine["sizeof.avifGainMap"])
    assert ctypes.sizeof(abi.avifDiagnostics) == int(mine["sizeof.avifDiagnostics"])
    for key, off in mine.items():
        t, _, m = key.partition(".")
        if t == "avifGainMap":
            assert getattr(abi.avifGainMap, m).offset == int(off), key
    for key, off in mine.items():
        t, _, m = key.partition(".")
        if t in ("avifImage", "avifRGBImage"):
            assert getattr(getattr(abi, t), m).offset == int(off), key


def test_header_coexists_with_reference_header():
    """A translation unit may include the reference's avif.h first and ours second (only the avifB200* extensions are added)."""
    if not os.path.isdir(REF_INCLUDE):
        pytest.skip("reference headers absent")
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "c.c")
        open(src, "w").write('#include "avif/avif.h"\n#include "avif/internal.h"\n#include "avif_b200.h"\nint main(void){return (int)sizeof(avifImage) == 0;}\n')
        subprocess.run(["gcc", "-std=gnu11", "-c", "-o", os.path.join(td, "c.o"), src, f"-I{REF_INCLUDE}", f"-I{os.path.join(ROOT, 'include')}"], check=True)


def test_library_exports_every_declared_symbol(product):
    decl = re.findall(r"AVIF_B200_API\s+[\w\s\*]+?\b(avif\w+)\s*\(", open(HEADER).read())
    assert len(decl) >= 35
    from libavif_b200 import abi
    for name in decl:
        assert hasattr(product, name), f"{name} declared in include/avif_b200.h but not exported"
        assert name in abi.PROTOTYPES, f"{name} has no ctypes prototype in libavif_b200/abi.py"
    nm = subprocess.run(["nm", "-D", "--defined-only", product._name], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (\w+)", nm))
    assert set(decl) <= exported
    # nothing of the oracle may be reachable from the product
    assert not any(s.startswith("oracle") for s in exported)
    ldd = subprocess.run(["ldd", product._name], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "libavif_ref" not in ldd


def test_no_gpu_means_loud_failure_not_fallback(product):
    """On a machine without a CUDA device conversions must fail (UNKNOWN_ERROR), never silently compute on the CPU."""
    import libavif_b200
    from libavif_b200 import abi
    from tests import util
    if libavif_b200.device_available():
        pytest.skip("a GPU is present")
    spec = util.Spec(w=16, h=8)
    assert util.yuv_to_rgb(product, spec, 1)[0] == abi.AVIF_RESULT_UNKNOWN_ERROR
    assert util.rgb_to_yuv(product, spec, 1)[0] == abi.AVIF_RESULT_UNKNOWN_ERROR


def test_internal_seam_library_exports_the_reference_hook_names():
    """libavif_b200_hooks.so (B-2): the six functions src/reformat.c / src/alpha.c call before their built-in loops
    (include/avif/internal.h:350-387) plus avifLibYUVVersion, and nothing that would interpose the public API."""
    import ctypes
    import os
    import subprocess
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(here, "libavif_b200", "csrc", "libavif_b200_hooks.so")
    assert os.path.exists(path), "run __graft_entry__.build()"
    lib = ctypes.CDLL(path)
    hooks = ["avifImageRGBToYUVLibYUV", "avifImageYUVToRGBLibYUV", "avifImageRGBToYUVLibSharpYUV", "avifRGBImageToF16LibYUV",
             "avifRGBImagePremultiplyAlphaLibYUV", "avifRGBImageUnpremultiplyAlphaLibYUV", "avifLibYUVVersion"]
    for name in hooks:
        assert hasattr(lib, name), name
    lib.avifLibYUVVersion.restype = ctypes.c_uint
    assert lib.avifLibYUVVersion() == 0
    exported = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    names = {line.split()[-1] for line in exported.splitlines() if " T " in line}
    assert names == set(hooks), names ^ set(hooks)


def test_header_coexists_with_the_reference_headers_and_cxx():
    """INTEGRATION.md's seams include avif_b200.h AFTER the reference's own headers (only the avifB200* extensions are then
    declared), and C++ callers include it alone: both must compile."""
    inc = os.path.join(ROOT, "include")
    with tempfile.TemporaryDirectory() as td:
        cxx = os.path.join(td, "a.cpp")
        open(cxx, "w").write('#include "avif_b200.h"\nint main() { avifRGBImage r; avifGainMap g; (void)r; (void)g; return 0; }\n')
        subprocess.run(["g++", "-std=c++17", "-fsyntax-only", f"-I{inc}", cxx], check=True)
        if os.path.isdir(REF_INCLUDE):
            c = os.path.join(td, "b.c")
            open(c, "w").write('#include "avif/avif.h"\n#include "avif/internal.h"\n#include "avif_b200.h"\n'
                               'int main(void) { return avifB200DeviceAvailable() & 0; }\n')
            subprocess.run(["gcc", "-std=gnu11", "-fsyntax-only", f"-I{REF_INCLUDE}", f"-I{inc}", c], check=True)


def test_c_client_example_builds_and_fails_loudly_without_a_device(tmp_path):
    """examples/convert_frame.c is a plain C caller of the public ABI.  It must link against the library as is; on a machine
    without a CUDA device the conversion must FAIL (exit code 3, no CPU fallback), with one it must succeed."""
    csrc = os.path.join(ROOT, "libavif_b200", "csrc")
    exe = str(tmp_path / "convert_frame")
    subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Wextra", "-Werror", f"-I{os.path.join(ROOT, 'include')}", os.path.join(ROOT, "examples", "convert_frame.c"),
                    f"-L{csrc}", "-lavif_b200", f"-Wl,-rpath,{csrc}", "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    import libavif_b200
    if libavif_b200.device_available():
        assert out.returncode == 0, out.stderr
        assert "kernel: y2r_" in out.stdout
    else:
        assert out.returncode == 3
        assert "no CPU fallback" in out.stderr
