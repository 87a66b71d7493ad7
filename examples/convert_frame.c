/*
 * convert_frame.c -- a plain C caller of the C-ABI, exactly as a libavif user writes it: fill an avifImage, ask for RGB.
 *
 *   gcc -std=gnu11 -Iinclude examples/convert_frame.c -Llibavif_b200/csrc -lavif_b200 -Wl,-rpath,$PWD/libavif_b200/csrc -o convert_frame
 *   ./convert_frame            # prints the first pixels and a checksum; exit code 0 = converted on the GPU
 *
 * Host pointers (malloc) are fine: the library stages them over PCIe.  Without a CUDA device the call fails with
 * AVIF_RESULT_UNKNOWN_ERROR -- there is no CPU fallback -- and this program exits with 3.
 */
#include "avif_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(void)
{
    const uint32_t W = 1920, H = 1080;
    avifImage image;
    memset(&image, 0, sizeof(image));
    image.width = W;
    image.height = H;
    image.depth = 10;
    image.yuvFormat = AVIF_PIXEL_FORMAT_YUV420;
    image.yuvRange = AVIF_RANGE_LIMITED;
    image.matrixCoefficients = AVIF_MATRIX_COEFFICIENTS_BT709;
    const uint32_t cw = (W + 1) / 2, ch = (H + 1) / 2;
    image.yuvRowBytes[0] = W * 2;
    image.yuvRowBytes[1] = image.yuvRowBytes[2] = cw * 2;
    image.yuvPlanes[0] = malloc((size_t)image.yuvRowBytes[0] * H);
    image.yuvPlanes[1] = malloc((size_t)image.yuvRowBytes[1] * ch);
    image.yuvPlanes[2] = malloc((size_t)image.yuvRowBytes[2] * ch);
    if (!image.yuvPlanes[0] || !image.yuvPlanes[1] || !image.yuvPlanes[2]) return 2;
    for (uint32_t j = 0; j < H; ++j) { /* a diagonal luma ramp over mid-grey chroma */
        uint16_t * row = (uint16_t *)(image.yuvPlanes[0] + (size_t)j * image.yuvRowBytes[0]);
        for (uint32_t i = 0; i < W; ++i) row[i] = (uint16_t)(64 + ((i + j) * 876u) / (W + H));
    }
    for (uint32_t j = 0; j < ch; ++j) {
        uint16_t * u = (uint16_t *)(image.yuvPlanes[1] + (size_t)j * image.yuvRowBytes[1]);
        uint16_t * v = (uint16_t *)(image.yuvPlanes[2] + (size_t)j * image.yuvRowBytes[2]);
        for (uint32_t i = 0; i < cw; ++i) {
            u[i] = (uint16_t)(512 + (i % 64));
            v[i] = (uint16_t)(512 - (j % 64));
        }
    }

    avifRGBImage rgb;
    avifRGBImageSetDefaults(&rgb, &image); /* RGBA, depth of the image, automatic (bilinear) chroma upsampling */
    avifResult res = avifRGBImageAllocatePixels(&rgb);
    if (res != AVIF_RESULT_OK) return 2;
    res = avifImageYUVToRGB(&image, &rgb);
    if (res != AVIF_RESULT_OK) {
        fprintf(stderr, "avifImageYUVToRGB failed with avifResult %d%s\n", (int)res,
                avifB200DeviceAvailable() ? "" : " (no CUDA device: this library has no CPU fallback)");
        return 3;
    }
    const uint16_t * px = (const uint16_t *)rgb.pixels;
    unsigned long long sum = 0;
    for (size_t k = 0; k < (size_t)W * H * 4; ++k) sum += px[k];
    printf("first pixel RGBA = %u %u %u %u, channel sum = %llu, kernel: %s\n", px[0], px[1], px[2], px[3], sum, avifB200LastKernelName());
    avifRGBImageFreePixels(&rgb);
    free(image.yuvPlanes[0]);
    free(image.yuvPlanes[1]);
    free(image.yuvPlanes[2]);
    return 0;
}
