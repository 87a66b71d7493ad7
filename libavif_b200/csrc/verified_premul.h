/* Constants (r, k) for which trunc(fma(c*a, r, k)) equals the reference's integer premultiply
 * floorf((float)c * (float)a / maxF + 0.5f) (src/alpha.c:187-189) for EVERY pair of codes 0 <= c, a <= 2^depth - 1.
 * Proven exhaustively by tools/verify_premul.c (each line is its output, bad=0), re-checked on the host by
 * avifB200VerifyPremulFma() (tests/test_host_logic.py) and on the device by avifB200SelfTestPremulFma()
 * (tests/test_gpu_division.py).  No entry for 16 bits: the product itself is rounded there. */
#ifndef AVIF_B200_VERIFIED_PREMUL_H
#define AVIF_B200_VERIFIED_PREMUL_H
#include <stdint.h>
typedef struct AbPremulConst { uint32_t depth, rBits, kBits; } AbPremulConst;
static const AbPremulConst kAbVerifiedPremul[] = {
    {  8, 0x3b808081u, 0x3f000000u }, /* r = RN(1/255) +0 ulp, k = 0.5 +0 ulp, bad=0 over 65536 pairs */
    {  9, 0x3b004020u, 0x3f000000u }, /* r = RN(1/511) +0 ulp, k = 0.5 +0 ulp, bad=0 over 262144 pairs */
    { 10, 0x3a802008u, 0x3f000000u }, /* r = RN(1/1023) +0 ulp, k = 0.5 +0 ulp, bad=0 over 1048576 pairs */
    { 11, 0x3a001002u, 0x3f000000u }, /* r = RN(1/2047) +0 ulp, k = 0.5 +0 ulp, bad=0 over 4194304 pairs */
    { 12, 0x39800800u, 0x3f000000u }, /* r = RN(1/4095) -1 ulp, k = 0.5 +0 ulp, bad=0 over 16777216 pairs */
};
#endif
