/*
 * verify_premul.c -- exhaustive proof that the integer premultiply of the reference,
 *     p = (uintN_t)floorf((float)c * (float)a / maxF + 0.5f)            (src/alpha.c:187-189, :215-217)
 * equals, for EVERY pair of codes 0 <= c, a <= max = 2^depth - 1 with depth <= 12,
 *     p = trunc(fma(x, r, k)),   x = (float)c * (float)a   (exact: x < 2^24)
 * for the constants (r, k) printed below.  The kernels use this form (y2r_fast.cuh, AB_MUL_INT_PREMUL_FMA) only with the
 * constants of libavif_b200/csrc/verified_premul.h, each line of which was printed by this program with bad=0.  The
 * library repeats the check on demand (avifB200VerifyPremulFma, host) and on the device (avifB200SelfTestPremulFma).
 *
 * Why it can hold: x/max + 0.5 reaches an integer n only at x = n*max - max/2, and max is odd, so the nearest products are
 * 0.5/max away from the jump -- hundreds of float spacings for depth <= 12 -- which leaves room for the rounding errors of
 * both forms.  At 16 bits x itself is rounded and the spacing of the quotient exceeds that margin: not used there.
 *
 * Build/run:  gcc -O2 -ffp-contract=off tools/verify_premul.c -o /tmp/vp -lm && /tmp/vp
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static uint64_t count_bad(uint32_t max, float r, float k)
{
    const float maxF = (float)max;
    uint64_t bad = 0;
    for (uint32_t a = 0; a <= max; ++a) {
        for (uint32_t c = 0; c <= max; ++c) {
            const float x = (float)c * (float)a;
            bad += ((uint32_t)floorf(x / maxF + 0.5f) != (uint32_t)fmaf(x, r, k));
        }
    }
    return bad;
}
int main(void)
{
    static const int order[] = { 0, -1, 1, -2, 2, -3, 3 };
    for (int depth = 8; depth <= 12; ++depth) {
        const uint32_t max = (1u << depth) - 1;
        const float r0 = 1.0f / (float)max;
        int found = 0;
        for (int i = 0; i < 7 && !found; ++i) {
            for (int j = 0; j < 7 && !found; ++j) {
                const float r = asf(asu(r0) + order[i]), k = asf(asu(0.5f) + order[j]);
                if (count_bad(max, r, k) == 0) {
                    printf("    { %2d, 0x%08xu, 0x%08xu }, /* r = RN(1/%u) %+d ulp, k = 0.5 %+d ulp, bad=0 over %llu pairs */\n", depth, asu(r),
                           asu(k), max, order[i], order[j], (unsigned long long)(max + 1) * (max + 1));
                    found = 1;
                }
            }
        }
        if (!found) printf("    /* depth %d: no constants found */\n", depth);
    }
    return 0;
}
