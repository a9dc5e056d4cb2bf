// Per-module DSP, one voice per thread, for sm_100a.
//
// Arithmetic contract (bit-for-bit with the reference wherever the reference is
// deterministic): every f32/f64 add/sub/mul below is an explicit round-to-nearest
// intrinsic, so nothing can be contracted into an FMA whatever -fmad says (Rust
// never contracts a*b+c); divisions are IEEE (-prec-div=true, default); denormals
// are kept (-ftz=false, default).  The only library calls are the f64 sin/exp2/pow
// of the oscillator and Non-Linear module, where the reference itself goes through
// the platform libm (<= 2 ulp f64 here vs glibc => at most a rare 1-ulp f32 flip).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "program.hpp"

namespace srk {
namespace dsp {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// f64 `x % 1.0` (Rust) == fmod(x, 1.0): exact, sign of x.  x - trunc(x) is exact for
// every finite x (the subtraction of two doubles with the same exponent range is
// representable), NaN for +-inf like fmod.
__device__ __forceinline__ double fmod1(double x) { return dsub(x, trunc(x)); }

// TransitionDetector::is_transition, src/synth.rs:292-297
__device__ __forceinline__ bool transition(bool& last, float val) {
  bool above = val > 0.0f;
  bool t = above && !last;
  last = above;
  return t;
}

// OscillatorModule::poly_blep, src/synth/oscillator.rs:50-67
__device__ __forceinline__ double poly_blep(double t, double dt) {
  if (dt == 0.0) return 0.0;
  if (t < dt) {
    t = __ddiv_rn(t, dt);
    return dsub(dsub(dadd(t, t), dmul(t, t)), 1.0);
  } else if (t > dsub(1.0, dt)) {
    t = __ddiv_rn(dsub(t, 1.0), dt);
    return dadd(dadd(dadd(dmul(t, t), t), t), 1.0);
  }
  return 0.0;
}

// Philox4x32-10 (Salmon et al. 2011): the seeded stand-in for the reference's
// unseeded rand::random (oscillator.rs:385); counter = (sample/4, voice, module).
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// Shared-memory views of one thread's column of the per-block tables.
struct Lane {
  uint32_t* st;        // state   [S][T], this thread's column
  const uint32_t* pr;  // params  [P][T]
  float* wires;        // wires   [W][K][T]
  int T;               // threads per block (column stride)
};

template <int K>
__device__ __forceinline__ float* wire(const Lane& L, int slot) {
  return slot >= 0 ? L.wires + (size_t)slot * K * L.T : nullptr;
}

// ---- OscillatorModule::calc, src/synth/oscillator.rs:108-158 ----------------
template <int K>
__device__ __forceinline__ void op_osc(const Instr& ins, const Lane& L, int kk) {
  const int T = L.T;
  uint32_t* s = L.st + ins.state * T;
  double pos = __hiloint2double((int)s[T], (int)s[0]);
  bool last = s[2 * T] != 0u;
  const uint32_t* p = L.pr + ins.param * T;
  const double val = (double)__uint_as_float(p[0]);
  const double delta_const = __hiloint2double((int)p[2 * T], (int)p[T]);
  const double sr = (double)ins.imm;
  const bool aa = __uint_as_float(p[3 * T]) != 0.0f;
  const float* cv = wire<K>(L, ins.in[0]);
  const float* sync = wire<K>(L, ins.in[1]);
  float* sine = wire<K>(L, ins.out[0]);
  float* square = wire<K>(L, ins.out[1]);
  float* saw = wire<K>(L, ins.out[2]);
  for (int k = 0; k < kk; ++k) {
    if (sync) {  // with no sync input the detector sees 0.0 forever: `last` just goes false
      if (transition(last, sync[k * T])) pos = 0.0;
    } else {
      last = false;
    }
    // get_freq_in_hz (:43-48) then / sample_rate (:132)
    double delta = delta_const;
    if (cv) delta = __ddiv_rn(dmul(440.0, exp2(dadd((double)cv[k * T], val))), sr);
    if (sine) sine[k * T] = __double2float_rn(sin(dmul(dmul(pos, 3.14159265358979323846), 2.0)));
    if (square || saw) {
      const double pb0 = aa ? poly_blep(pos, delta) : 0.0;
      if (square) {
        const float base = pos < 0.5 ? -1.0f : 1.0f;
        const float corr = aa ? __double2float_rn(dsub(pb0, poly_blep(fmod1(dadd(pos, 0.5)), delta))) : 0.0f;
        square[k * T] = fsub(base, corr);
      }
      if (saw) {
        const float corr = aa ? __double2float_rn(pb0) : 0.0f;
        saw[k * T] = fsub(fsub(fmul(__double2float_rn(pos), 2.0f), 1.0f), corr);
      }
    }
    pos = fmod1(dadd(pos, delta));
  }
  s[0] = (uint32_t)__double2loint(pos);
  s[T] = (uint32_t)__double2hiint(pos);
  s[2 * T] = last ? 1u : 0u;
}

// ---- NoiseModule::calc, src/synth/oscillator.rs:381-388 (seeded generator) ----
template <int K>
__device__ __forceinline__ void op_noise(const Instr& ins, const Lane& L, int kk, uint32_t voice, uint32_t seed_lo,
                                         uint32_t seed_hi) {
  const int T = L.T;
  uint32_t* s = L.st + ins.state * T;
  uint64_t n = ((uint64_t)s[T] << 32) | s[0];
  float* out = wire<K>(L, ins.out[0]);
  if (out) {
    uint32_t c[4] = {0, 0, 0, 0};
    for (int k = 0; k < kk; ++k) {
      const uint64_t i = n + k;
      if (k == 0 || (i & 3) == 0) {
        const uint64_t blk = i >> 2;
        c[0] = (uint32_t)blk; c[1] = (uint32_t)(blk >> 32); c[2] = voice; c[3] = ins.aux;
        philox4x32_10(c, seed_lo, seed_hi);
      }
      const uint32_t lane = (uint32_t)(i & 3);
      const uint32_t r = lane == 0 ? c[0] : lane == 1 ? c[1] : lane == 2 ? c[2] : c[3];
      const float u = fmul((float)(r >> 8), 1.0f / 16777216.0f);  // rand 0.8.5 Standard f32
      out[k * T] = fmul(fsub(u, 0.5f), 2.0f);
    }
  }
  n += kk;
  s[0] = (uint32_t)n;
  s[T] = (uint32_t)(n >> 32);
}

// ---- MoogFilterModule::calc, src/synth/filter.rs:182-221 with
//      InternalMoogFilterState::calc :60-83 and clamp_buffers :86-91 -------------
__device__ __forceinline__ float clamp1(float x) { return fmaxf(fminf(x, 1.0f), -1.0f); }

template <int K>
__device__ __forceinline__ void op_moog(const Instr& ins, const Lane& L, int kk) {
  const int T = L.T;
  uint32_t* s = L.st + ins.state * T;
  float f = __uint_as_float(s[0]), p = __uint_as_float(s[T]), q = __uint_as_float(s[2 * T]);
  float b0 = __uint_as_float(s[3 * T]), b1 = __uint_as_float(s[4 * T]), b2 = __uint_as_float(s[5 * T]);
  float b3 = __uint_as_float(s[6 * T]), b4 = __uint_as_float(s[7 * T]);
  float c_freq = __uint_as_float(s[8 * T]), c_res = __uint_as_float(s[9 * T]);
  const uint32_t* pp = L.pr + ins.param * T;
  const float freq = __uint_as_float(pp[0]), res = __uint_as_float(pp[T]), exp_amt = __uint_as_float(pp[2 * T]);
  const float r = fminf(fmaxf(res, 0.0f), 1.0f);  // :214
  const float* audio = wire<K>(L, ins.in[0]);
  const float* cv = wire<K>(L, ins.in[1]);
  float* lowpass = wire<K>(L, ins.out[0]);
  float* bandpass = wire<K>(L, ins.out[1]);
  float* highpass = wire<K>(L, ins.out[2]);
  for (int k = 0; k < kk; ++k) {
    const float a = audio ? audio[k * T] : 0.0f;
    const float c = cv ? cv[k * T] : 0.0f;
    const float fc = fminf(fmaxf(fadd(freq, fmul(c, exp_amt)), 0.0f), 0.9f);  // :213
    if (fc != c_freq || r != c_res) {  // :61-68
      c_freq = fc;
      c_res = r;
      q = fsub(1.0f, fc);
      p = fadd(fc, fmul(fmul(0.8f, fc), q));
      f = fsub(fmul(p, 2.0f), 1.0f);
      q = fmul(r, fadd(1.0f, fmul(fmul(0.5f, q), fadd(fsub(1.0f, q), fmul(fmul(5.6f, q), q)))));
    }
    const float in = fsub(a, fmul(q, b4));  // :69
    float t1 = b1;
    b1 = fsub(fmul(fadd(in, b0), p), fmul(b1, f));
    const float t2 = b2;
    b2 = fsub(fmul(fadd(b1, t1), p), fmul(b2, f));
    t1 = b3;
    b3 = fsub(fmul(fadd(b2, t2), p), fmul(b3, f));
    b4 = fsub(fmul(fadd(b3, t1), p), fmul(b4, f));
    b4 = fsub(b4, fmul(fmul(fmul(b4, b4), b4), 0.166667f));  // powi(3)
    b0 = in;
    b0 = clamp1(b0); b1 = clamp1(b1); b2 = clamp1(b2); b3 = clamp1(b3); b4 = clamp1(b4);
    // calc returns (b4, in - b4, 3*(b3-b4)) assigned to (lowpass, highpass, bandpass), :211
    if (lowpass) lowpass[k * T] = b4;
    if (highpass) highpass[k * T] = fsub(in, b4);
    if (bandpass) bandpass[k * T] = fmul(3.0f, fsub(b3, b4));
  }
  s[0] = __float_as_uint(f); s[T] = __float_as_uint(p); s[2 * T] = __float_as_uint(q);
  s[3 * T] = __float_as_uint(b0); s[4 * T] = __float_as_uint(b1); s[5 * T] = __float_as_uint(b2);
  s[6 * T] = __float_as_uint(b3); s[7 * T] = __float_as_uint(b4);
  s[8 * T] = __float_as_uint(c_freq); s[9 * T] = __float_as_uint(c_res);
}

// ---- ADSRModule::calc, src/synth/adsr.rs:134-217 ------------------------------
template <int K>
__device__ __forceinline__ void op_adsr(const Instr& ins, const Lane& L, int kk) {
  const int T = L.T;
  uint32_t* s = L.st + ins.state * T;
  float phase = __uint_as_float(s[0]), r_val = __uint_as_float(s[T]), from_a_val = __uint_as_float(s[2 * T]);
  uint32_t mode = s[3 * T] & 0xFFu;
  bool last = (s[3 * T] >> 8) & 1u;
  const uint32_t* pp = L.pr + ins.param * T;
  const float a_sec = __uint_as_float(pp[0]), d_sec = __uint_as_float(pp[T]);
  const float s_val = __uint_as_float(pp[2 * T]), r_sec = __uint_as_float(pp[3 * T]);
  const float sr = ins.imm;
  // `1.0 / (self.sample_rate * self.x_sec)` is loop invariant: same IEEE value every sample
  const float inc_a = __fdiv_rn(1.0f, fmul(sr, a_sec));
  const float inc_d = __fdiv_rn(1.0f, fmul(sr, d_sec));
  const float inc_r = __fdiv_rn(1.0f, fmul(sr, r_sec));
  const float* gate = wire<K>(L, ins.in[0]);
  float* out = wire<K>(L, ins.out[0]);
  for (int k = 0; k < kk; ++k) {
    const float g = gate ? gate[k * T] : 0.0f;
    const bool high = gate && g > 0.0f;
    const bool tr = transition(last, g);
    if (mode == ADSR_NONE) {
      if (high) { phase = 0.0f; mode = ADSR_ATTACK; }
    } else if (mode == ADSR_ATTACK) {
      phase = fadd(phase, inc_a);
      if (phase >= 1.0f) { phase = 0.0f; mode = ADSR_DECAY; }
      else if (tr) { phase = 0.0f; r_val = from_a_val; }
    } else if (mode == ADSR_DECAY) {
      phase = fadd(phase, inc_d);
      if (phase >= 1.0f) { phase = 0.0f; mode = ADSR_SUSTAIN; }
      if (tr) { phase = 0.0f; mode = ADSR_ATTACK; }
    } else if (mode == ADSR_SUSTAIN) {
      if (!gate || g <= 0.0f) { phase = 0.0f; mode = ADSR_RELEASE; }
      if (tr) { phase = 0.0f; mode = ADSR_ATTACK; }
    } else {  // Release
      if (high) { phase = 0.0f; mode = ADSR_ATTACK; }
      phase = fadd(phase, inc_r);
      if (phase >= 1.0f) { phase = 0.0f; r_val = 0.0f; mode = ADSR_NONE; }
    }
    float o;
    if (mode == ADSR_NONE) o = 0.0f;
    else if (mode == ADSR_ATTACK) o = fadd(r_val, fmul(fsub(1.0f, r_val), phase));
    else if (mode == ADSR_DECAY) o = fadd(s_val, fmul(fsub(1.0f, s_val), fsub(1.0f, phase)));
    else if (mode == ADSR_SUSTAIN) o = s_val;
    else o = fmul(s_val, fsub(1.0f, phase));
    if (out) out[k * T] = o;
    if (mode != ADSR_ATTACK) r_val = o; else from_a_val = o;
  }
  s[0] = __float_as_uint(phase); s[T] = __float_as_uint(r_val); s[2 * T] = __float_as_uint(from_a_val);
  s[3 * T] = mode | (last ? 1u << 8 : 0u);
}

// ---- VCAModule::calc, src/synth/vca.rs:117-148 ---------------------------------
template <int K>
__device__ __forceinline__ void op_vca(const Instr& ins, const Lane& L, int kk) {
  const int T = L.T;
  const float* audio = wire<K>(L, ins.in[0]);
  const float* cv = wire<K>(L, ins.in[1]);
  float* out = wire<K>(L, ins.out[0]);
  if (!out) return;
  const bool negative = __uint_as_float(L.pr[ins.param * T]) != 0.0f;
  for (int k = 0; k < kk; ++k) {
    float o = 0.0f;
    if (audio && cv) {
      const float c = cv[k * T];
      o = (negative || c > 0.0f) ? fmul(audio[k * T], c) : 0.0f;
    }
    out[k * T] = o;
  }
}

// ---- MonoMixerModule::calc, src/synth/mixer.rs:101-122 -------------------------
template <int K>
__device__ __forceinline__ void op_mixer(const Instr& ins, const Lane& L, int kk) {
  const int T = L.T;
  float* out = wire<K>(L, ins.out[0]);
  if (!out) return;
  const uint32_t* pp = L.pr + ins.param * T;
  const float* in[4];
  float gain[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    in[j] = wire<K>(L, ins.in[j]);
    gain[j] = __uint_as_float(pp[j * T]);
  }
  for (int k = 0; k < kk; ++k) {
    float o = 0.0f;  // output.fill(0.0) then `*dst += src * gain` per connected input, in order
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (in[j]) o = fadd(o, fmul(in[j][k * T], gain[j]));
    out[k * T] = o;
  }
}

// ---- MathModule / NonLinearModule::calc, src/synth/math.rs:139-160, :292-313 ----
__device__ __forceinline__ float math_op(uint8_t which, float a, float b) {
  switch (which) {
    case F_MATH_ADD: return fadd(a, b);
    case F_MATH_SUB: return fsub(a, b);
    case F_MATH_MUL: return fmul(a, b);
    default:
      // math.rs:203-205 `if a > 0.0 { a.powf(b) } else { -(-a).powf(b) }` in f32.  glibc's powf
      // evaluates in f64 and rounds once; f64 pow here then one rounding agrees with it except
      // when the f64 results straddle an f32 rounding boundary.
      return a > 0.0f ? __double2float_rn(pow((double)a, (double)b)) : -__double2float_rn(pow((double)(-a), (double)b));
  }
}

template <int K>
__device__ __forceinline__ void op_math(const Instr& ins, const Lane& L, int kk) {
  const int T = L.T;
  float* out = wire<K>(L, ins.out[0]);
  if (!out) return;
  const float constant = __uint_as_float(L.pr[ins.param * T]);
  const float* i1 = wire<K>(L, ins.in[0]);
  const float* i2 = wire<K>(L, ins.in[1]);
  for (int k = 0; k < kk; ++k) {
    const float a = i1 ? i1[k * T] : 0.0f;          // (None, _) => 0.0
    const float b = i2 ? i2[k * T] : constant;      // (_, None) => constant
    out[k * T] = math_op(ins.flags, a, b);
  }
}

}  // namespace dsp
}  // namespace srk
