/* Divisors for which div_const()'s 3-op sequence equals IEEE division on the whole domain
 * (x == +0 or 2^-60 <= |x| <= 2^60).  Proven exhaustively by tools/verify_divisors.c (bad=0 for each) and
 * re-checked on the device by tests/test_gpu_division.py.  Values are binary32 bit patterns, sorted.
 * Origin: kg, 2(1-kb), 2(1-kr) of every supported matrix (table MCs and chroma-derived NCL for each
 * known set of primaries), the luma/chroma code ranges of every depth, and (2^depth - 1). */
#ifndef AVIF_B200_VERIFIED_DIVISORS_H
#define AVIF_B200_VERIFIED_DIVISORS_H
#include <stdint.h>
static const uint32_t kAbVerifiedDivisors[] = {
    0x3f161fb4u, 0x3f1645a1u, 0x3f170a3du, 0x3f2c18a0u, 0x3f2d9147u, 0x3f2d9169u, 0x3f2da76au, 0x3f3115c6u,
    0x3f3374bcu, 0x3f3378a8u, 0x3f34e753u, 0x3f37154au, 0x3f371759u, 0x3f38ba77u, 0x3f800000u, 0x3fb33333u,
    0x3fb374bcu, 0x3fb376ecu, 0x3fbcbfadu, 0x3fbcbfb2u, 0x3fbf1507u, 0x3fc4abfeu, 0x3fc561ebu, 0x3fc72ab9u,
    0x3fc9907cu, 0x3fc9930cu, 0x3fc9a1b3u, 0x3fc9ba5eu, 0x3fca5ec0u, 0x3fe2a8c8u, 0x3fe2d0e5u, 0x3fe3d70au,
    0x3fe76ca2u, 0x3fe9ba5eu, 0x3fe9d6f5u, 0x3febb3dbu, 0x3fed844du, 0x3fed84ceu, 0x3fedbc9au, 0x3fee5bb7u,
    0x3fee9263u, 0x3ff0d19au, 0x3ff0d1b7u, 0x40000000u, 0x435b0000u, 0x43600000u, 0x437f0000u, 0x445b0000u,
    0x44600000u, 0x447fc000u, 0x455b0000u, 0x45600000u, 0x457ff000u, 0x475b0000u, 0x47600000u, 0x477fff00u,
};
/* The 2-operation form fma(x, rh, RN(x*rl)) with rh + rl = 1/d in double precision equals IEEE division on the same domain for
 * every divisor above EXCEPT these (tools/verify_quant.c `div`: bad = 240 for 0x3fee5bb7, 0 for the other 55; re-checked on the
 * device by tests/test_gpu_division.py through avifB200SelfTestDivisor2). */
static const uint32_t kAbNotTwoOpDivisors[] = { 0x3fee5bb7u };
#endif
