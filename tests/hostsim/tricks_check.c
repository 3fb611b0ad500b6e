/*
 * TEST INFRASTRUCTURE ONLY.  CPU cross-check (random sampling; the exhaustive proof of the division
 * runs on the GPU, tests/cuda/divcheck.cu) of the three arithmetic identities the production walk
 * (csrc/ray_fast.cuh) relies on, evaluated with the host's IEEE fmaf and rounding modes:
 *   1. floor(x / 2^L)          == significand field of  RD(fma(x, 2^-L, 2^23))       for 0 <= x < 2^23 * 2^L
 *   2. (floor(x / 2^L) + 1)*2^L == fma(RD(..), 2^L, 2^L * (1 - 2^23))                 (exact)
 *   3. a / d (IEEE)            == fma(fma(-d, a*r, a), r, a*r)  with r = RN(1/d)      for normal a, d
 * Returns the number of mismatches.
 */
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#pragma STDC FENV_ACCESS ON

static uint64_t rng_state;
static uint32_t rnd32(void) {
  rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (uint32_t)(rng_state >> 32);
}
static float as_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint32_t as_uint(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

long tricks_check(uint64_t seed, long samples) {
  long bad = 0;
  rng_state = seed;
  for (long s = 0; s < samples; ++s) {
    /* 1 + 2: level L, coordinate x with a random significand and an exponent below 23 + L */
    const int L = (int)(rnd32() % 16);
    const float c = ldexpf(1.0f, L), ic = ldexpf(1.0f, -L), kc = c * -8388607.0f;
    const int e = (int)(rnd32() % (24 + L)) - 1; /* x in [2^(e-1), 2^e) or tiny */
    volatile float x = ldexpf(as_float(0x3f000000u | (rnd32() & 0x7fffffu)), e);
    if ((s & 1023) == 0) x = 0.0f;
    fesetround(FE_DOWNWARD);
    volatile float sfl = fmaf(x, ic, 8388608.0f);
    fesetround(FE_TONEAREST);
    const float want_floor = floorf(x / c);
    if ((as_uint(sfl) & 0x7fffffu) != (uint32_t)want_floor) ++bad;
    volatile float b = fmaf(sfl, c, kc);
    if (b != (want_floor + 1.0f) * c) ++bad;
    /* 3: division with random significands, divisor in [2^-40, 1], numerator in [2^-24, 2^24] */
    volatile float d = ldexpf(as_float(0x3f800000u | (rnd32() & 0x7fffffu)), -(int)(rnd32() % 40) - 1);
    volatile float a = ldexpf(as_float(0x3f800000u | (rnd32() & 0x7fffffu)), (int)(rnd32() % 48) - 24);
    volatile float r = 1.0f / d;
    volatile float q0 = a * r;
    volatile float rem = fmaf(-d, q0, a);
    volatile float q = fmaf(rem, r, q0);
    volatile float want = a / d;
    if (as_uint(q) != as_uint(want)) ++bad;
  }
  return bad;
}
