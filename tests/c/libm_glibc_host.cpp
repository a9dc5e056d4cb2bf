// Host build of s-rack_b200/csrc/libm_glibc.cuh for tests/test_libm_glibc.py: the restatements of glibc's exp2 (f64) and powf
// exactly as the device compiles them, minus the intrinsics.  Built with -O2 -ffp-contract=off -mfma (std::fma is then one
// instruction; contraction stays off so that only the explicit fma()s fuse).
#include "../../s-rack_b200/csrc/libm_glibc.cuh"

extern "C" {
void t_exp2_glibc(const double* x, double* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = exp2_glibc(x[i]);
}
void t_powf_glibc(const float* x, const float* y, float* z, long n) {
  for (long i = 0; i < n; ++i) z[i] = powf_glibc(x[i], y[i]);
}
}
extern "C" void t_sin_glibc(const double* x, double* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = sin_glibc(x[i]);
}
// the tie-band narrowing of the sine port, fed a `fast` sine that is `ulps[i]` bit patterns away from glibc's
extern "C" void t_sin_settle(const double* x, const long* ulps, float* y, long n) {
  for (long i = 0; i < n; ++i) {
    const double g = std::sin(x[i]);
    y[i] = lg_sin_settle(x[i], lg_f64(lg_bits(g) + (unsigned long long)ulps[i]));
  }
}
// the fast sine of the common route against the platform's: largest distance in bit patterns over
//   mode 0: n arguments (float64)(pos * PI * 2.0), pos uniform in [0, 1) -- as the oscillator forms them
//   mode 1: n arguments uniform in (-lim, lim)
//   mode 2: every f64 within n patterns of k * pi / 2, k = 1 .. 8 (both signs)
// and whether (float) of the settled value equals (float) of the platform's everywhere (-> *narrow_bad)
#include <cstdint>
extern "C" long t_sin_fast_scan(int mode, long n, double lim, unsigned long long seed, double* worst_x, long* narrow_bad) {
  unsigned long long st = seed * 0x9E3779B97F4A7C15ull + 1;
  auto rnd = [&st]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) * 0x1p-53; };
  long worst = 0, bad = 0;
  auto one = [&](double x) {
    const double g = std::sin(x), f = lg_sin_fast(x);
    long d = (long)(lg_bits(f) - lg_bits(g));
    if ((lg_bits(f) ^ lg_bits(g)) >> 63) d = (long)((lg_bits(f) & ~(1ull << 63)) + (lg_bits(g) & ~(1ull << 63)));  // across zero
    if (d < 0) d = -d;
    if (d > worst) { worst = d; *worst_x = x; }
    const float a = sinf_of_f64_glibc(x), b = (float)g;
    if (lg_bitsf(a) != lg_bitsf(b)) ++bad;
  };
  if (mode == 0) for (long i = 0; i < n; ++i) one(rnd() * 3.14159265358979323846 * 2.0);
  if (mode == 1) for (long i = 0; i < n; ++i) one((rnd() * 2.0 - 1.0) * lim);
  if (mode == 2)
    for (int k = 1; k <= 8; ++k) {
      const double c = k * 1.5707963267948966;
      for (long i = -n; i <= n; ++i) { const double x = lg_f64(lg_bits(c) + (unsigned long long)i); one(x); one(-x); }
    }
  *narrow_bad = bad;
  return worst;
}
extern "C" void t_sinf_of_f64(const double* x, float* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = sinf_of_f64_glibc(x[i]);
}
// the sample-group form of exp2 (main route for all four, one branch for the group)
extern "C" void t_exp2_group(const double* x, double* y, long n) {
  long i = 0;
  for (; i + 4 <= n; i += 4) {
    const double a[4] = {x[i], x[i + 1], x[i + 2], x[i + 3]};
    double b[4];
    exp2_glibc_group(a, b);
    for (int j = 0; j < 4; ++j) y[i + j] = b[j];
  }
  for (; i < n; ++i) { const double a[1] = {x[i]}; double b[1]; exp2_glibc_group(a, b); y[i] = b[0]; }
}
