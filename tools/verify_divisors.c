/*
 * verify_divisors.c -- exhaustive proof, per divisor d, that the 3-operation sequence
 *     q = RN(x * r);  e = fma(-q, d, x);  q' = fma(e, r, q)        with r = RN(1/d)
 * returns the IEEE-754 correctly rounded quotient RN(x / d) for EVERY binary32 x with x == +0 or
 * 2^-60 <= |x| <= 2^60 (all 2^32 bit patterns are visited; out-of-domain ones are skipped).
 * The kernels use this sequence (pixel_math.cuh: div_const) only for divisors listed in
 * libavif_b200/csrc/verified_divisors.h, every entry of which was printed with bad=0 by this program.
 * The same check is repeated on the GPU by tests/test_gpu_division.py through avifB200SelfTestDivisor().
 *
 * Build/run:  gcc -O2 -mfma -fopenmp -ffp-contract=off tools/verify_divisors.c -o /tmp/vd -lm && /tmp/vd 3f371759 ...
 * (-mfma only makes fmaf() a single instruction; fmaf is correctly rounded either way.)
 */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
static inline float mk(float x, float d, float r){ float q = x*r; float e = fmaf(-q,d,x); return fmaf(e,r,q); }
int main(int argc,char**argv){
  for (int k=1;k<argc;k++){
    uint32_t db=(uint32_t)strtoul(argv[k],0,16); float d; memcpy(&d,&db,4); float r=1.0f/d; uint64_t bad=0;
    #pragma omp parallel for reduction(+:bad) schedule(static)
    for (int64_t b=0;b<(1ll<<32);b++){
      uint32_t u=(uint32_t)b; float x; memcpy(&x,&u,4);
      float ax=fabsf(x);
      if (!((u==0) || (ax>=0x1p-60f && ax<=0x1p60f))) continue;
      float a=mk(x,d,r), t=x/d;
      uint32_t ua,ut; memcpy(&ua,&a,4); memcpy(&ut,&t,4);
      if (ua!=ut) bad++;
    }
    printf("%08x %.9g bad=%llu\n", db, d, (unsigned long long)bad); fflush(stdout);
  }
}
