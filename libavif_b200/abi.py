"""ctypes mirror of include/avif_b200.h (== the reference's avif.h:777-851, :948-1018 for this path).

Pure declarations: struct layouts, enum values and the function prototypes of the C-ABI.  Nothing here
computes anything; the compute lives in libavif_b200/csrc (CUDA).  The same Structure classes are used by
the tests to drive the oracle (oracle/liboracle.so) and the compiled reference (oracle/_ref/libavif_ref.so),
because all three libraries share one ABI.
"""
import ctypes as C

# --- enums (values: include/avif_b200.h; reference include/avif/avif.h) ------------------------------
AVIF_RESULT_OK = 0
AVIF_RESULT_UNKNOWN_ERROR = 1
AVIF_RESULT_REFORMAT_FAILED = 5
AVIF_RESULT_INVALID_ARGUMENT = 24
AVIF_RESULT_NOT_IMPLEMENTED = 25
AVIF_RESULT_OUT_OF_MEMORY = 26
AVIF_RESULT_INTERNAL_ERROR = 29

AVIF_PIXEL_FORMAT_NONE, AVIF_PIXEL_FORMAT_YUV444, AVIF_PIXEL_FORMAT_YUV422, AVIF_PIXEL_FORMAT_YUV420, \
    AVIF_PIXEL_FORMAT_YUV400, AVIF_PIXEL_FORMAT_COUNT = range(6)
AVIF_RANGE_LIMITED, AVIF_RANGE_FULL = 0, 1

(AVIF_RGB_FORMAT_RGB, AVIF_RGB_FORMAT_RGBA, AVIF_RGB_FORMAT_ARGB, AVIF_RGB_FORMAT_BGR, AVIF_RGB_FORMAT_BGRA,
 AVIF_RGB_FORMAT_ABGR, AVIF_RGB_FORMAT_RGB_565, AVIF_RGB_FORMAT_GRAY, AVIF_RGB_FORMAT_GRAYA, AVIF_RGB_FORMAT_AGRAY,
 AVIF_RGB_FORMAT_COUNT) = range(11)
RGB_FORMAT_NAMES = ["RGB", "RGBA", "ARGB", "BGR", "BGRA", "ABGR", "RGB_565", "GRAY", "GRAYA", "AGRAY"]

AVIF_CHROMA_UPSAMPLING_AUTOMATIC, AVIF_CHROMA_UPSAMPLING_FASTEST, AVIF_CHROMA_UPSAMPLING_BEST_QUALITY, \
    AVIF_CHROMA_UPSAMPLING_NEAREST, AVIF_CHROMA_UPSAMPLING_BILINEAR = range(5)
AVIF_CHROMA_DOWNSAMPLING_AUTOMATIC, AVIF_CHROMA_DOWNSAMPLING_FASTEST, AVIF_CHROMA_DOWNSAMPLING_BEST_QUALITY, \
    AVIF_CHROMA_DOWNSAMPLING_AVERAGE, AVIF_CHROMA_DOWNSAMPLING_SHARP_YUV = range(5)

MC_IDENTITY, MC_BT709, MC_UNSPECIFIED, MC_FCC, MC_BT470BG, MC_BT601, MC_SMPTE240, MC_YCGCO, MC_BT2020_NCL = 0, 1, 2, 4, 5, 6, 7, 8, 9
MC_BT2020_CL, MC_SMPTE2085, MC_CHROMA_DERIVED_NCL, MC_CHROMA_DERIVED_CL, MC_ICTCP, MC_YCGCO_RE, MC_YCGCO_RO = 10, 11, 12, 13, 14, 16, 17

AVIF_PLANES_YUV, AVIF_PLANES_A, AVIF_PLANES_ALL = 1, 2, 0xFF


def rgb_format_has_alpha(fmt):
    return fmt not in (AVIF_RGB_FORMAT_RGB, AVIF_RGB_FORMAT_BGR, AVIF_RGB_FORMAT_RGB_565, AVIF_RGB_FORMAT_GRAY)


def rgb_format_is_gray(fmt):
    return fmt in (AVIF_RGB_FORMAT_GRAY, AVIF_RGB_FORMAT_GRAYA, AVIF_RGB_FORMAT_AGRAY)


def rgb_format_channel_count(fmt):
    if fmt == AVIF_RGB_FORMAT_GRAY:
        return 1
    if fmt in (AVIF_RGB_FORMAT_GRAYA, AVIF_RGB_FORMAT_AGRAY):
        return 2
    return 4 if rgb_format_has_alpha(fmt) else 3


def rgb_pixel_size(fmt, depth):
    if fmt == AVIF_RGB_FORMAT_RGB_565:
        return 2
    return rgb_format_channel_count(fmt) * (2 if depth > 8 else 1)


# --- structs ----------------------------------------------------------------------------------------
class avifRWData(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t)]


class avifImage(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("depth", C.c_uint32),
        ("yuvFormat", C.c_int), ("yuvRange", C.c_int), ("yuvChromaSamplePosition", C.c_int),
        ("yuvPlanes", C.c_void_p * 3), ("yuvRowBytes", C.c_uint32 * 3), ("imageOwnsYUVPlanes", C.c_int),
        ("alphaPlane", C.c_void_p), ("alphaRowBytes", C.c_uint32), ("imageOwnsAlphaPlane", C.c_int),
        ("alphaPremultiplied", C.c_int),
        ("icc", avifRWData),
        ("colorPrimaries", C.c_uint16), ("transferCharacteristics", C.c_uint16), ("matrixCoefficients", C.c_uint16),
        ("clli", C.c_uint16 * 2),
        ("transformFlags", C.c_uint32),
        ("pasp", C.c_uint32 * 2), ("clap", C.c_uint32 * 8),
        ("irot", C.c_uint8), ("imir", C.c_uint8),
        ("exif", avifRWData), ("xmp", avifRWData),
        ("properties", C.c_void_p), ("numProperties", C.c_size_t), ("gainMap", C.c_void_p),
    ]


class avifRGBImage(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("depth", C.c_uint32),
        ("format", C.c_int), ("chromaUpsampling", C.c_int), ("chromaDownsampling", C.c_int),
        ("avoidLibYUV", C.c_int), ("ignoreAlpha", C.c_int), ("alphaPremultiplied", C.c_int), ("isFloat", C.c_int),
        ("maxThreads", C.c_int),
        ("pixels", C.c_void_p), ("rowBytes", C.c_uint32),
    ]


class avifAlphaParams(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("srcDepth", C.c_uint32), ("srcPlane", C.c_void_p), ("srcRowBytes", C.c_uint32),
        ("srcOffsetBytes", C.c_uint32), ("srcPixelBytes", C.c_uint32),
        ("dstDepth", C.c_uint32), ("dstPlane", C.c_void_p), ("dstRowBytes", C.c_uint32),
        ("dstOffsetBytes", C.c_uint32), ("dstPixelBytes", C.c_uint32),
    ]


class avifSignedFraction(C.Structure):
    _fields_ = [("n", C.c_int32), ("d", C.c_uint32)]


class avifUnsignedFraction(C.Structure):
    _fields_ = [("n", C.c_uint32), ("d", C.c_uint32)]


class avifDiagnostics(C.Structure):
    _fields_ = [("error", C.c_char * 256)]


class avifContentLightLevelInformationBox(C.Structure):
    _fields_ = [("maxCLL", C.c_uint16), ("maxPALL", C.c_uint16)]


class avifGainMap(C.Structure):  # avif.h:630-712
    _fields_ = [
        ("image", C.POINTER(avifImage)),
        ("gainMapMin", avifSignedFraction * 3), ("gainMapMax", avifSignedFraction * 3), ("gainMapGamma", avifUnsignedFraction * 3),
        ("baseOffset", avifSignedFraction * 3), ("alternateOffset", avifSignedFraction * 3),
        ("baseHdrHeadroom", avifUnsignedFraction), ("alternateHdrHeadroom", avifUnsignedFraction),
        ("useBaseColorSpace", C.c_int),
        ("altICC", avifRWData),
        ("altColorPrimaries", C.c_uint16), ("altTransferCharacteristics", C.c_uint16), ("altMatrixCoefficients", C.c_uint16),
        ("altYUVRange", C.c_int), ("altDepth", C.c_uint32), ("altPlaneCount", C.c_uint32),
        ("altCLLI", avifContentLightLevelInformationBox),
    ]


class avifPixelFormatInfo(C.Structure):
    _fields_ = [("monochrome", C.c_int), ("chromaShiftX", C.c_int), ("chromaShiftY", C.c_int)]


# --- prototypes of every symbol include/avif_b200.h declares -------------------------------------------
_PI, _PR, _PA = C.POINTER(avifImage), C.POINTER(avifRGBImage), C.POINTER(avifAlphaParams)


class avifCropRect(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32)]


class avifRGBColorSpaceInfo(C.Structure):  # include/avif/internal.h:297-309
    _fields_ = [("channelBytes", C.c_uint32), ("pixelBytes", C.c_uint32), ("offsetBytesR", C.c_uint32), ("offsetBytesG", C.c_uint32),
                ("offsetBytesB", C.c_uint32), ("offsetBytesA", C.c_uint32), ("offsetBytesGray", C.c_uint32), ("maxChannel", C.c_int),
                ("maxChannelF", C.c_float)]


class avifYUVColorSpaceInfo(C.Structure):  # include/avif/internal.h:314-332
    _fields_ = [("kr", C.c_float), ("kg", C.c_float), ("kb", C.c_float), ("channelBytes", C.c_uint32), ("depth", C.c_uint32), ("range", C.c_int),
                ("maxChannel", C.c_int), ("biasY", C.c_float), ("biasUV", C.c_float), ("rangeY", C.c_float), ("rangeUV", C.c_float),
                ("formatInfo", avifPixelFormatInfo), ("mode", C.c_int)]


PROTOTYPES = {
    # reference public ABI
    "avifImageYUVToRGB": (C.c_int, [_PI, _PR]),
    "avifImageRGBToYUV": (C.c_int, [_PI, _PR]),
    "avifRGBImagePremultiplyAlpha": (C.c_int, [_PR]),
    "avifRGBImageUnpremultiplyAlpha": (C.c_int, [_PR]),
    "avifFullToLimitedY": (C.c_int, [C.c_uint32, C.c_int]),
    "avifFullToLimitedUV": (C.c_int, [C.c_uint32, C.c_int]),
    "avifLimitedToFullY": (C.c_int, [C.c_uint32, C.c_int]),
    "avifLimitedToFullUV": (C.c_int, [C.c_uint32, C.c_int]),
    "avifB200ImageYUVToRGBRowsAsync": (C.c_int, [_PI, _PR, C.c_uint32, C.c_uint32, C.c_void_p]),
    "avifB200GridGetUniqueId": (C.c_int, [C.POINTER(C.c_uint8 * 128)]),
    "avifB200GridCommInit": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_uint8 * 128)]),
    "avifB200GridCommDestroy": (C.c_int, [C.c_void_p]),
    "avifB200GridCellRect": (C.c_int, [_PI, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(avifCropRect)]),
    "avifB200GridRGBToYUV": (C.c_int, [_PI, _PR, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float * 4)]),
    "avifB200GridYUVToRGB": (C.c_int, [_PI, _PR, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float * 3)]),
    "avifB200GridLastMode": (C.c_int, []),
    "avifB200SetDevice": (C.c_int, [C.c_int]),
    "avifB200SetDevices": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "avifB200GetDevice": (C.c_int, []),
    "avifGetRGBColorSpaceInfo": (C.c_int, [_PR, C.POINTER(avifRGBColorSpaceInfo)]),
    "avifGetYUVColorSpaceInfo": (C.c_int, [_PI, C.POINTER(avifYUVColorSpaceInfo)]),
    "avifGetRGBAPixel": (None, [_PR, C.c_uint32, C.c_uint32, C.POINTER(avifRGBColorSpaceInfo), C.POINTER(C.c_float * 4)]),
    "avifSetRGBAPixel": (None, [_PR, C.c_uint32, C.c_uint32, C.POINTER(avifRGBColorSpaceInfo), C.POINTER(C.c_float * 4)]),
    "avifRGBFormatChannelCount": (C.c_uint32, [C.c_int]),
    "avifRGBFormatHasAlpha": (C.c_int, [C.c_int]),
    "avifRGBFormatIsGray": (C.c_int, [C.c_int]),
    "avifRGBImageSetDefaults": (None, [_PR, _PI]),
    "avifRGBImagePixelSize": (C.c_uint32, [_PR]),
    "avifRGBImageAllocatePixels": (C.c_int, [_PR]),
    "avifRGBImageFreePixels": (None, [_PR]),
    "avifGetPixelFormatInfo": (None, [C.c_int, C.POINTER(avifPixelFormatInfo)]),
    "avifLibYUVVersion": (C.c_uint, []),
    # reference internal seam
    "avifFillAlpha": (None, [_PA]),
    "avifReformatAlpha": (None, [_PA]),
    # B200 extensions
    "avifB200ImageYUVToRGBAsync": (C.c_int, [_PI, _PR, C.c_void_p]),
    "avifB200ImageRGBToYUVAsync": (C.c_int, [_PI, _PR, C.c_void_p]),
    "avifB200RGBImagePremultiplyAlphaAsync": (C.c_int, [_PR, C.c_void_p]),
    "avifB200RGBImageUnpremultiplyAlphaAsync": (C.c_int, [_PR, C.c_void_p]),
    "avifB200ImageYUVToRGBBatchAsync": (C.c_int, [C.POINTER(_PI), C.POINTER(_PR), C.c_uint32, C.c_void_p]),
    "avifB200ImageRGBToYUVBatchAsync": (C.c_int, [C.POINTER(_PI), C.POINTER(_PR), C.c_uint32, C.c_void_p]),
    "avifB200ImageYUVToRGBHost": (C.c_int, [_PI, _PR]),
    "avifB200ImageRGBToYUVHost": (C.c_int, [_PI, _PR]),
    "avifB200RGBImagePremultiplyAlphaHost": (C.c_int, [_PR]),
    "avifB200RGBImageUnpremultiplyAlphaHost": (C.c_int, [_PR]),
    "avifB200ImageLimitedToFullAlpha": (C.c_int, [_PI]),
    "avifB200ImageCopySamplesAsync": (C.c_int, [_PI, _PI, C.c_uint32, C.c_void_p]),
    "avifB200ImageCopyAndPadAsync": (C.c_int, [_PI, _PI, C.c_void_p]),
    "avifB200AllocPinned": (C.c_void_p, [C.c_size_t]),
    "avifB200FreePinned": (None, [C.c_void_p]),
    "avifB200SetKernelPolicy": (None, [C.c_int]),
    "avifB200KernelLaunchCount": (C.c_uint64, []),
    "avifB200ResetKernelLaunchCount": (None, []),
    "avifB200LastKernelName": (C.c_char_p, []),
    "avifB200VerifiedDivisorCount": (C.c_size_t, []),
    "avifB200VerifiedDivisorBits": (C.c_uint32, [C.c_size_t]),
    "avifB200SelfTestDivisor": (C.c_int, [C.c_uint32, C.POINTER(C.c_uint64)]),
    "avifB200SelfTestIntDivisor": (C.c_int, [C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "avifB200SelfTestDivisor2": (C.c_int, [C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "avifB200SelfTestAlphaDivision": (C.c_int, [C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]),
    "avifRGBImageApplyGainMap": (C.c_int, [_PR, C.c_uint16, C.c_uint16, C.POINTER(avifGainMap), C.c_float, C.c_uint16, C.c_uint16, _PR,
                                           C.POINTER(avifContentLightLevelInformationBox), C.POINTER(avifDiagnostics)]),
    "avifB200DebugScalePlan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "avifB200DebugScalePlanColumns": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "avifB200DebugPrimariesMatrix": (C.c_int, [C.c_uint16, C.c_uint16, C.POINTER(C.c_double)]),
    "avifImageScale": (C.c_int, [_PI, C.c_uint32, C.c_uint32, C.POINTER(avifDiagnostics)]),
    "avifB200ImageScaleToDevice": (C.c_int, [_PI, _PI]),
    "avifB200RGBImageApplyGainMapDevice": (C.c_int, [_PR, C.c_uint16, C.c_uint16, C.POINTER(avifGainMap), C.c_float, C.c_uint16, C.c_uint16, _PR,
                                                     C.POINTER(avifContentLightLevelInformationBox), C.POINTER(avifDiagnostics)]),
    "avifImageApplyGainMap": (C.c_int, [_PI, C.POINTER(avifGainMap), C.c_float, C.c_uint16, C.c_uint16, _PR,
                                        C.POINTER(avifContentLightLevelInformationBox), C.POINTER(avifDiagnostics)]),
    "avifB200HookImageYUVToRGB": (C.c_int, [_PI, _PR, C.c_int, C.POINTER(C.c_int)]),
    "avifB200HookImageRGBToYUV": (C.c_int, [_PI, _PR]),
    "avifB200HookRGBImageToF16": (C.c_int, [_PR]),
    "avifB200VerifyPremulFma": (C.c_int, [C.c_uint32, C.POINTER(C.c_uint64)]),
    "avifB200SelfTestPremulFma": (C.c_int, [C.c_uint32, C.POINTER(C.c_uint64)]),
    "avifB200DeviceAvailable": (C.c_int, []),
    "avifB200Version": (C.c_char_p, []),
}


def bind(lib, prototypes=None, prefix_map=None):
    """Attach restype/argtypes.  prefix_map lets the oracle's `oracle*` names be bound with the same prototypes."""
    for name, (res, args) in (prototypes or PROTOTYPES).items():
        sym = prefix_map(name) if prefix_map else name
        try:
            fn = getattr(lib, sym)
        except AttributeError:
            continue
        fn.restype = res
        fn.argtypes = args
    return lib
