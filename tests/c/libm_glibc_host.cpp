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
