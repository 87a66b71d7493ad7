"""Host-side mirror of the reference's image containers (avifImageCreate + avifImageAllocatePlanes,
src/avif.c:431-489; avifRGBImageSetDefaults + avifRGBImageAllocatePixels, src/avif.c:700-741).

Buffers are numpy arrays (host; optionally CUDA-pinned) or torch uint8 tensors (device-resident planes).
PyTorch is used for device memory only; no arithmetic happens here.
"""
import ctypes

import numpy as np

from . import abi


def _chroma_dims(width, height, yuv_format):
    sx = 0 if yuv_format == abi.AVIF_PIXEL_FORMAT_YUV444 else 1
    sy = 1 if yuv_format == abi.AVIF_PIXEL_FORMAT_YUV420 else 0
    return (width + sx) >> sx, (height + sy) >> sy


class _Buf:
    """One 2-D byte buffer (rows x row_bytes) on the host (numpy) or on a CUDA device (torch)."""

    def __init__(self, rows, row_bytes, device=None, pinned=False, offset=0):
        self.rows, self.row_bytes, self.device, self.offset = rows, row_bytes, device, offset
        n = rows * row_bytes + offset
        if device is None:
            if pinned:
                import torch
                self._t = torch.empty(n, dtype=torch.uint8).pin_memory()
                self.flat = self._t.numpy()
            else:
                self.flat = np.zeros(n, dtype=np.uint8)
            self.ptr = self.flat.ctypes.data + offset
        else:
            import torch
            self._t = torch.zeros(n, dtype=torch.uint8, device=device)
            self.flat = None
            self.ptr = self._t.data_ptr() + offset

    @property
    def nbytes(self):
        return self.rows * self.row_bytes

    def host(self):
        """numpy view/copy of the bytes as (rows, row_bytes)."""
        if self.device is None:
            return self.flat[self.offset:self.offset + self.nbytes].reshape(self.rows, self.row_bytes)
        return self._t[self.offset:self.offset + self.nbytes].cpu().numpy().reshape(self.rows, self.row_bytes)

    def set(self, arr2d):
        arr = np.ascontiguousarray(arr2d, dtype=np.uint8).reshape(-1)
        assert arr.size == self.nbytes
        if self.device is None:
            self.flat[self.offset:self.offset + self.nbytes] = arr
        else:
            import torch
            self._t[self.offset:self.offset + self.nbytes].copy_(torch.from_numpy(arr))


class YUVImage:
    """avifImage + the storage behind its planes."""

    def __init__(self, width, height, depth, yuv_format, yuv_range=abi.AVIF_RANGE_FULL, mc=abi.MC_BT601, alpha=False,
                 alpha_premultiplied=False, color_primaries=1, device=None, pinned=False, row_pad=0, offset=0,
                 allocate=True):
        self.c = abi.avifImage()
        c = self.c
        c.width, c.height, c.depth = width, height, depth
        c.yuvFormat, c.yuvRange = yuv_format, yuv_range
        c.matrixCoefficients, c.colorPrimaries, c.transferCharacteristics = mc, color_primaries, 2
        c.alphaPremultiplied = int(alpha_premultiplied)
        self.device = device
        self.bps = 2 if depth > 8 else 1
        self.bufs = [None, None, None, None]  # Y U V A
        if allocate:
            self.allocate(alpha, pinned=pinned, row_pad=row_pad, offset=offset)

    def plane_dims(self, p):
        c = self.c
        if p in (0, 3):
            return c.width, c.height
        return _chroma_dims(c.width, c.height, c.yuvFormat)

    def allocate(self, alpha, pinned=False, row_pad=0, offset=0):
        c = self.c
        planes = [0] if c.yuvFormat == abi.AVIF_PIXEL_FORMAT_YUV400 else [0, 1, 2]
        if alpha:
            planes.append(3)
        for p in planes:
            w, h = self.plane_dims(p)
            rb = w * self.bps + row_pad
            b = _Buf(h, rb, self.device, pinned, offset)
            self.bufs[p] = b
            if p < 3:
                c.yuvPlanes[p], c.yuvRowBytes[p] = b.ptr, rb
            else:
                c.alphaPlane, c.alphaRowBytes = b.ptr, rb

    def ptr(self):
        return ctypes.byref(self.c)

    def plane_samples(self, p):
        """(h, w) integer array of the samples of plane p (row padding stripped)."""
        w, h = self.plane_dims(p)
        raw = self.bufs[p].host()
        if self.bps == 2:
            return raw[:, :w * 2].copy().view(np.uint16).reshape(h, w)
        return raw[:, :w].copy()

    def set_plane_samples(self, p, samples):
        w, h = self.plane_dims(p)
        raw = self.bufs[p].host().copy()
        s = np.ascontiguousarray(samples, dtype=np.uint16 if self.bps == 2 else np.uint8).reshape(h, w)
        raw[:, :w * self.bps] = s.view(np.uint8).reshape(h, w * self.bps)
        self.bufs[p].set(raw)

    def present_planes(self):
        return [p for p in range(4) if self.bufs[p] is not None]

    def payload_bytes(self):
        """Concatenated visible bytes of every present plane in Y,U,V,A order (the hash domain of SURVEY App. D)."""
        out = []
        for p in self.present_planes():
            w, _ = self.plane_dims(p)
            out.append(self.bufs[p].host()[:, :w * self.bps].reshape(-1))
        return np.concatenate(out)


class RGBImage:
    """avifRGBImage + the storage behind rgb.pixels; defaults follow avifRGBImageSetDefaults (src/avif.c:700-717)."""

    # rgb->avoidLibYUV: irrelevant to this library (always built-in-path results); a reference build only consults its
    # internal libyuv seam when it is 0 (tests of libavif_b200_hooks.so clear it)
    AVOID_LIBYUV = 1

    def __init__(self, width, height, depth, fmt=abi.AVIF_RGB_FORMAT_RGBA, upsampling=abi.AVIF_CHROMA_UPSAMPLING_AUTOMATIC,
                 downsampling=abi.AVIF_CHROMA_DOWNSAMPLING_AUTOMATIC, ignore_alpha=False, alpha_premultiplied=False,
                 is_float=False, max_threads=1, device=None, pinned=False, row_pad=0, offset=0, allocate=True):
        self.c = abi.avifRGBImage()
        c = self.c
        c.width, c.height, c.depth, c.format = width, height, depth, fmt
        c.chromaUpsampling, c.chromaDownsampling = upsampling, downsampling
        c.avoidLibYUV, c.ignoreAlpha, c.alphaPremultiplied, c.isFloat = self.AVOID_LIBYUV, int(ignore_alpha), int(alpha_premultiplied), int(is_float)
        c.maxThreads = max_threads
        self.device = device
        self.pixel_bytes = abi.rgb_pixel_size(fmt, depth) if 0 <= fmt < abi.AVIF_RGB_FORMAT_COUNT else 4
        self.buf = None
        if allocate:
            rb = width * self.pixel_bytes + row_pad
            self.buf = _Buf(height, rb, device, pinned, offset)
            c.pixels, c.rowBytes = self.buf.ptr, rb

    def ptr(self):
        return ctypes.byref(self.c)

    def payload_bytes(self):
        return self.buf.host()[:, :self.c.width * self.pixel_bytes].reshape(-1)

    def full_bytes(self):
        return self.buf.host().reshape(-1)

    def set_payload(self, arr2d):
        raw = self.buf.host().copy()
        a = np.ascontiguousarray(arr2d, dtype=np.uint8).reshape(self.c.height, self.c.width * self.pixel_bytes)
        raw[:, :a.shape[1]] = a
        self.buf.set(raw)


def pointer_array(items):
    """ctypes array of pointers to the structs of YUVImage / RGBImage wrappers (argument of the *BatchAsync entry points)."""
    import ctypes as C
    if not items:
        return None
    t = type(items[0].c)
    return (C.POINTER(t) * len(items))(*[C.pointer(it.c) for it in items])
