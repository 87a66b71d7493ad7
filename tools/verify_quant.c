/*
 * verify_quant.c -- exhaustive proofs behind two shorter forms of the RGB->YUV arithmetic (r2y_fast.cuh).
 *
 * (1) merged rounding constant.  The reference quantises  clamp((int)floorf((v*range + bias) + 0.5f), 0, yuvMax)
 *     (src/reformat.c:197-219).  For every binary32 v (all 2^32 bit patterns) and every (range, bias) pair the path can
 *     meet, is   floorf(RN(RN(v*range) + bias) + 0.5)  ==  floorf(RN(v*range) + (bias + 0.5))   after the clamp?
 *     (bias + 0.5 is exactly representable.)  Printed per pair: bad = number of v where the two differ.
 *
 * (2) two-operation division.  For a divisor d and every binary32 x in the domain of verify_divisors.c, is
 *     fma(x, rh, RN(x * rl)) == x / d, with rh + rl the double-precision 1/d split into two floats?
 *
 * Build/run:  gcc -O2 -mfma -fopenmp -ffp-contract=off tools/verify_quant.c -o /tmp/vq -lm
 *             /tmp/vq quant            (all pairs)      /tmp/vq div 3fc9930c 3fed84ce ...
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int clampi(float t, int max)
{
    if (!(t == t)) return 0; /* NaN: cvt.rzi gives 0 on the device; both forms are NaN together */
    if (t < 0.0f) return 0;
    if (t > (float)max + 1.0f) return max;
    int q = (int)floorf(t);
    return q < 0 ? 0 : (q > max ? max : q);
}

int main(int argc, char ** argv)
{
    if (argc >= 2 && !strcmp(argv[1], "quant")) {
        const int depths[4] = { 8, 10, 12, 16 };
        for (int di = 0; di < 4; ++di) {
            const int depth = depths[di], max = (1 << depth) - 1;
            /* (range, bias): limited Y, limited UV, full Y (bias 0), full UV; identity mode quantises UV with Y's pair */
            const float pairs[4][2] = { { (float)(219 << (depth - 8)), (float)(16 << (depth - 8)) },
                                        { (float)(224 << (depth - 8)), (float)(1 << (depth - 1)) },
                                        { (float)max, 0.0f },
                                        { (float)max, (float)(1 << (depth - 1)) } };
            for (int p = 0; p < 4; ++p) {
                const float range = pairs[p][0], bias = pairs[p][1], bh = bias + 0.5f;
                uint64_t bad = 0, bad_fma = 0;
#pragma omp parallel for reduction(+ : bad, bad_fma) schedule(static)
                for (int64_t b = 0; b < (1ll << 32); b++) {
                    uint32_t u = (uint32_t)b;
                    float v;
                    memcpy(&v, &u, 4);
                    const float pr = v * range;
                    const float t1 = (pr + bias) + 0.5f;
                    const float t2 = pr + bh;
                    if (clampi(t1, max) != clampi(t2, max)) bad++;
                    if (clampi(t1, max) != clampi(fmaf(v, range, bh), max)) bad_fma++;
                }
                printf("depth %2d range %g bias %g: bad=%llu  (single fma(v, range, bias + 0.5): bad=%llu)\n", depth, range, bias, (unsigned long long)bad,
                       (unsigned long long)bad_fma);
                fflush(stdout);
            }
        }
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "div")) {
        for (int k = 2; k < argc; k++) {
            uint32_t db = (uint32_t)strtoul(argv[k], 0, 16);
            float d;
            memcpy(&d, &db, 4);
            const double inv = 1.0 / (double)d;
            const float rh = (float)inv, rl = (float)(inv - (double)rh);
            uint64_t bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
            for (int64_t b = 0; b < (1ll << 32); b++) {
                uint32_t u = (uint32_t)b;
                float x;
                memcpy(&x, &u, 4);
                float ax = fabsf(x);
                if (!((u == 0) || (ax >= 0x1p-60f && ax <= 0x1p60f))) continue;
                float a = fmaf(x, rh, x * rl), t = x / d;
                uint32_t ua, ut;
                memcpy(&ua, &a, 4);
                memcpy(&ut, &t, 4);
                if (ua != ut) bad++;
            }
            printf("%08x %.9g bad=%llu\n", db, d, (unsigned long long)bad);
            fflush(stdout);
        }
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "rgbquant")) {
        /* (3) YUV->RGB quantiser (src/reformat.c:952, :1021): (T)(0.5f + c*rgbMaxF) for c in [0,1] -- is the product's own rounding
         * redundant, i.e. floor(RN(0.5 + RN(c*max))) == floor(RN(fma(c, max, 0.5))) for EVERY float c in [0,1] and max = 2^depth - 1? */
        const int depths[4] = { 8, 10, 12, 16 };
        for (int di = 0; di < 4; ++di) {
            const float max = (float)((1 << depths[di]) - 1);
            uint64_t bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
            for (int64_t b = 0; b <= 0x3F800000ll; b++) {
                uint32_t u = (uint32_t)b;
                float c;
                memcpy(&c, &u, 4);
                const float t1 = 0.5f + c * max;
                const float t2 = fmaf(c, max, 0.5f);
                if ((uint32_t)t1 != (uint32_t)t2) bad++;
            }
            printf("rgb depth %2d: bad=%llu\n", depths[di], (unsigned long long)bad);
            fflush(stdout);
        }
        return 0;
    }
    fprintf(stderr, "usage: %s quant | rgbquant | div <hex divisor>...\n", argv[0]);
    return 2;
}
