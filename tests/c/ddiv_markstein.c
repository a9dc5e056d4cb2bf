/* The division by a render-constant divisor used on the device (fused_ops.cuh ddiv_by_const): with y = RN(1 / b),
 *     q0 = a y;  q1 = fma(fma(-b, q0, a), y, q0);  q = fma(fma(-b, q1, a), y, q1)
 * must equal the IEEE quotient RN(a / b) bit for bit on the operand ranges that occur:
 *   (1) polyBLEP (oscillator.rs:50-67): a = t or t - 1 with t in [0, 1), b = delta in [2^-200, 1/8)
 *   (2) the V/oct conversion (oscillator.rs:132): a = 440 * 2^x with a biased exponent in [64, 1984), b = the sample
 *       rate, an integer in [1, 65535] (u16 in the reference, synth.rs:21)
 * Build: gcc -O2 -ffp-contract=off -mfma ddiv_markstein.c -lm ; prints the number of mismatches (0 expected) and exits with it.
 * usage: ddiv_markstein [pairs per range, default 20000000] */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static double u01(void) { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }
static double bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

static double ddiv_by_const(double a, double b, double y) {
  const double q0 = a * y;
  const double q1 = fma(fma(-b, q0, a), y, q0);
  return fma(fma(-b, q1, a), y, q1);
}

static long check(double a, double b) {
  const double y = 1.0 / b;
  const double q = ddiv_by_const(a, b, y), ref = a / b;
  if (memcmp(&q, &ref, 8) != 0) {
    static int shown = 0;
    if (shown++ < 10) printf("MISMATCH a=%a b=%a got %a want %a\n", a, b, q, ref);
    return 1;
  }
  return 0;
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 20000000;
  long bad = 0;
  for (long i = 0; i < n; ++i) {  /* (1) */
    const double t = u01();
    double b;
    switch (i & 3) {
      case 0: b = ldexp(0.5 + 0.5 * u01(), -3 - (int)(rnd() % 20)); break;    /* audio rates */
      case 1: b = ldexp(0.5 + 0.5 * u01(), -3 - (int)(rnd() % 197)); break;   /* down to 2^-200 */
      case 2: b = 440.0 * exp2(-4.0 + 8.0 * u01()) / 48000.0; break;
      default: b = bits((0x3fbull << 52) | (rnd() >> 12)); b = b < 0.125 ? b : 0.1; break;
    }
    if (b >= 0.125) b = 0.124999;
    bad += check((i & 4) ? t : t - 1.0, b);
    if (t < b) bad += check(t, b);
  }
  for (long i = 0; i < n; ++i) {  /* (2) */
    const int sr = 1 + (int)(rnd() % 65535);
    const double b = (i & 1) ? (double)sr : ((i & 2) ? 48000.0 : 44100.0);
    const uint64_t e = 64 + rnd() % 1920;
    const double a = (i & 4) ? bits((e << 52) | (rnd() >> 12)) : 440.0 * exp2(-12.0 + 24.0 * u01());
    bad += check(a, b);
  }
  printf("%ld mismatches in %ld pairs\n", bad, 2 * n);
  return bad ? 1 : 0;
}
