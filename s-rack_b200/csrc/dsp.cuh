// Per-module DSP, one voice per lane, for sm_100a.
//
// Arithmetic contract (bit-for-bit with the reference wherever the reference is
// deterministic): every f32/f64 add/sub/mul below is an explicit round-to-nearest
// intrinsic, so nothing can be contracted into an FMA whatever -fmad says (Rust
// never contracts a*b+c); divisions are IEEE (-prec-div=true, default); denormals
// are kept (-ftz=false, default).  The only library calls are the f64 sin/exp2/pow
// of the oscillator and Non-Linear module, where the reference itself goes through
// the platform libm (<= 2 ulp f64 here vs glibc => at most a rare 1-ulp f32 flip).
//
// Shape of every op: the chunk is walked in groups of 8 samples (then single
// samples for a ragged tail).  Inside a group the code is straight-line: inputs are
// loaded first, the *stateless* per-sample work (coefficients, 2^x, sin, polyBLEP)
// is independent across the 8 samples, and only the true recurrence (phase, ladder
// stages, envelope state machine) forms a dependent chain -- ptxas overlaps the two.
// "Is this port connected" is decided once per op or per group, never per sample.
// Wire tiles are [K samples][32 voices] f32 in shared memory: sample k of this lane
// is p[k * 32], so offsets inside a group are immediates.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

#include "program.hpp"

namespace srk {
namespace dsp {

constexpr int L = kVoicesPerGroup;  // lane stride of tiles, state and params

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// f64 `x % 1.0` (Rust) == fmod(x, 1.0): exact, sign of x.  x - trunc(x) is exact for
// every finite x, NaN for +-inf like fmod.  The phase accumulator only ever sees
// x in [0, 2), where the result is x or x - 1 (both exact): that path avoids the
// slow f64 round instruction on the recurrence.
__device__ __forceinline__ double fmod1(double x) {
  if (x >= 0.0 && x < 1.0) return x;
  if (x >= 1.0 && x < 2.0) return dsub(x, 1.0);
  return dsub(x, trunc(x));
}

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

// This lane's view of its voice group's shared-memory tables.
struct Lane {
  uint32_t* st;        // state  [S][32], this lane's column
  const uint32_t* pr;  // params [P][32]
  float* tiles;        // wire tiles, this lane's column of tile 0
  const WireDesc* wd;  // wire slot -> (first tile, ring mask)
  uint32_t tile_elems; // K * 32
  uint32_t chunk;      // chunk the current instruction works on
};

__device__ __forceinline__ float* wire(const Lane& ln, int slot) {
  if (slot < 0) return nullptr;
  const WireDesc d = ln.wd[slot];
  return ln.tiles + (size_t)(d.base + (ln.chunk & d.mask)) * ln.tile_elems;
}

template <int U>
using UC = std::integral_constant<int, U>;

// Runs body(UC<8>, k0) over full groups of 8 samples, body(UC<1>, k) over the tail.
template <class Body>
__device__ __forceinline__ void for_groups(int kk, Body&& body) {
  int k0 = 0;
  for (; k0 + 8 <= kk; k0 += 8) body(UC<8>(), k0);
  for (; k0 < kk; ++k0) body(UC<1>(), k0);
}

// ---- OscillatorModule::calc, src/synth/oscillator.rs:108-158 ----------------
template <bool HAS_CV, bool HAS_SYNC>
__device__ __forceinline__ void op_osc(const Instr& ins, const Lane& ln, int kk) {
  uint32_t* s = ln.st + ins.state * L;
  double pos = __hiloint2double((int)s[L], (int)s[0]);
  bool last = s[2 * L] != 0u;
  const uint32_t* p = ln.pr + ins.param * L;
  const double val = (double)__uint_as_float(p[0]);
  const double delta_const = __hiloint2double((int)p[2 * L], (int)p[L]);
  const double sr = (double)ins.imm;
  const bool aa = __uint_as_float(p[3 * L]) != 0.0f;
  const float* cv = wire(ln, ins.in[0]);
  const float* sync = wire(ln, ins.in[1]);
  float* sine = wire(ln, ins.out[0]);
  float* square = wire(ln, ins.out[1]);
  float* saw = wire(ln, ins.out[2]);
  for_groups(kk, [&](auto u, int k0) {
    constexpr int U = decltype(u)::value;
    float cvv[U], syv[U];
    double ps[U], dl[U];
    if (HAS_CV) {
#pragma unroll
      for (int j = 0; j < U; ++j) cvv[j] = cv[(k0 + j) * L];
    }
    if (HAS_SYNC) {
#pragma unroll
      for (int j = 0; j < U; ++j) syv[j] = sync[(k0 + j) * L];
    }
    // get_freq_in_hz (:43-48) then / sample_rate (:132): stateless
#pragma unroll
    for (int j = 0; j < U; ++j)
      dl[j] = HAS_CV ? __ddiv_rn(dmul(440.0, exp2(dadd((double)cvv[j], val))), sr) : delta_const;
    // the recurrence: sync reset (:125-131), pos += delta; pos %= 1.0 (:151-152)
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (HAS_SYNC && transition(last, syv[j])) pos = 0.0;
      ps[j] = pos;
      pos = fmod1(dadd(pos, dl[j]));
    }
    if (sine) {
#pragma unroll
      for (int j = 0; j < U; ++j)
        sine[(k0 + j) * L] = __double2float_rn(sin(dmul(dmul(ps[j], 3.14159265358979323846), 2.0)));
    }
    if (square || saw) {
      double pb0[U];
#pragma unroll
      for (int j = 0; j < U; ++j) pb0[j] = aa ? poly_blep(ps[j], dl[j]) : 0.0;
      if (square) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const float base = ps[j] < 0.5 ? -1.0f : 1.0f;
          const float corr = aa ? __double2float_rn(dsub(pb0[j], poly_blep(fmod1(dadd(ps[j], 0.5)), dl[j]))) : 0.0f;
          square[(k0 + j) * L] = fsub(base, corr);
        }
      }
      if (saw) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const float corr = aa ? __double2float_rn(pb0[j]) : 0.0f;
          saw[(k0 + j) * L] = fsub(fsub(fmul(__double2float_rn(ps[j]), 2.0f), 1.0f), corr);
        }
      }
    }
  });
  // with no sync input the detector sees 0.0 every sample: `last` just goes false
  if (!HAS_SYNC && kk > 0) last = false;
  s[0] = (uint32_t)__double2loint(pos);
  s[L] = (uint32_t)__double2hiint(pos);
  s[2 * L] = last ? 1u : 0u;
}

__device__ __forceinline__ void op_osc_dispatch(const Instr& ins, const Lane& ln, int kk) {
  const bool cv = ins.in[0] >= 0, sync = ins.in[1] >= 0;
  if (cv) {
    if (sync) op_osc<true, true>(ins, ln, kk);
    else op_osc<true, false>(ins, ln, kk);
  } else {
    if (sync) op_osc<false, true>(ins, ln, kk);
    else op_osc<false, false>(ins, ln, kk);
  }
}

// ---- NoiseModule::calc, src/synth/oscillator.rs:381-388 (seeded generator) ----
__device__ __forceinline__ void op_noise(const Instr& ins, const Lane& ln, int kk, uint32_t voice, uint32_t seed_lo,
                                         uint32_t seed_hi) {
  uint32_t* s = ln.st + ins.state * L;
  uint64_t n = ((uint64_t)s[L] << 32) | s[0];
  float* out = wire(ln, ins.out[0]);
  if (out) {
    auto draw = [&](uint64_t blk, uint32_t (&c)[4]) {
      c[0] = (uint32_t)blk; c[1] = (uint32_t)(blk >> 32); c[2] = voice; c[3] = ins.aux;
      philox4x32_10(c, seed_lo, seed_hi);
    };
    auto shape = [](uint32_t r) {
      const float u = fmul((float)(r >> 8), 1.0f / 16777216.0f);  // rand 0.8.5 Standard f32
      return fmul(fsub(u, 0.5f), 2.0f);
    };
    int k = 0;
    uint32_t c[4];
    // ragged head up to the next multiple of 4 of the absolute sample counter
    if ((n & 3) != 0 && kk > 0) {
      draw(n >> 2, c);
      for (; k < kk && ((n + k) & 3) != 0; ++k) {
        const uint32_t q = (uint32_t)((n + k) & 3);
        out[k * L] = shape(q == 1 ? c[1] : q == 2 ? c[2] : c[3]);
      }
    }
    for (; k + 4 <= kk; k += 4) {  // one Philox block per 4 samples
      draw((n + k) >> 2, c);
      out[(k + 0) * L] = shape(c[0]);
      out[(k + 1) * L] = shape(c[1]);
      out[(k + 2) * L] = shape(c[2]);
      out[(k + 3) * L] = shape(c[3]);
    }
    if (k < kk) {
      draw((n + k) >> 2, c);
      for (int q = 0; k < kk; ++k, ++q) out[k * L] = shape(q == 0 ? c[0] : q == 1 ? c[1] : c[2]);
    }
  }
  n += kk;
  s[0] = (uint32_t)n;
  s[L] = (uint32_t)(n >> 32);
}

// ---- MoogFilterModule::calc, src/synth/filter.rs:182-221 with
//      InternalMoogFilterState::calc :60-83 and clamp_buffers :86-91 -------------
__device__ __forceinline__ float clamp1(float x) { return fmaxf(fminf(x, 1.0f), -1.0f); }

// The coefficient block of :61-68 as a pure function of (frequency, res).
__device__ __forceinline__ void moog_coef(float fc, float r, float& f, float& p, float& q) {
  q = fsub(1.0f, fc);
  p = fadd(fc, fmul(fmul(0.8f, fc), q));
  f = fsub(fmul(p, 2.0f), 1.0f);
  q = fmul(r, fadd(1.0f, fmul(fmul(0.5f, q), fadd(fsub(1.0f, q), fmul(fmul(5.6f, q), q)))));
}

// The reference caches (freq, res) and recomputes (f, p, q) when either changes.  Since
// the block is pure, the cached coefficients always equal moog_coef(current fc, r) --
// except while the state is still the all-zero Default (:48) and (fc, r) == (0, 0) hits
// that zero cache, where they stay 0 (moog_coef(0,0) has f = -1, so `f == 0` with a zero
// cache key identifies it).  That makes the coefficients stateless per sample: they are
// computed off the ladder's dependency chain.
template <bool HAS_AUDIO, bool HAS_CV>
__device__ __forceinline__ void op_moog(const Instr& ins, const Lane& ln, int kk) {
  uint32_t* s = ln.st + ins.state * L;
  float f = __uint_as_float(s[0]), p = __uint_as_float(s[L]), q = __uint_as_float(s[2 * L]);
  float b0 = __uint_as_float(s[3 * L]), b1 = __uint_as_float(s[4 * L]), b2 = __uint_as_float(s[5 * L]);
  float b3 = __uint_as_float(s[6 * L]), b4 = __uint_as_float(s[7 * L]);
  float c_freq = __uint_as_float(s[8 * L]), c_res = __uint_as_float(s[9 * L]);
  const uint32_t* pp = ln.pr + ins.param * L;
  const float freq = __uint_as_float(pp[0]), res = __uint_as_float(pp[L]), exp_amt = __uint_as_float(pp[2 * L]);
  const float r = fminf(fmaxf(res, 0.0f), 1.0f);  // :214
  const float* audio = wire(ln, ins.in[0]);
  const float* cv = wire(ln, ins.in[1]);
  float* lowpass = wire(ln, ins.out[0]);
  float* bandpass = wire(ln, ins.out[1]);
  float* highpass = wire(ln, ins.out[2]);
  bool virgin = c_freq == 0.0f && c_res == 0.0f && f == 0.0f;
  if (!HAS_CV && kk > 0) {  // cutoff is constant over the chunk: one cache check (:61)
    const float fc = fminf(fmaxf(fadd(freq, fmul(0.0f, exp_amt)), 0.0f), 0.9f);  // :213 with cv = 0.0
    if (fc != c_freq || r != c_res) {
      c_freq = fc;
      c_res = r;
      moog_coef(fc, r, f, p, q);
    }
  }
  for_groups(kk, [&](auto u, int k0) {
    constexpr int U = decltype(u)::value;
    float a[U], fj[U], pj[U], qj[U], in_[U], o3[U], o4[U];
#pragma unroll
    for (int j = 0; j < U; ++j) a[j] = HAS_AUDIO ? audio[(k0 + j) * L] : 0.0f;
    if (HAS_CV) {
      float fc[U];
#pragma unroll
      for (int j = 0; j < U; ++j) fc[j] = fminf(fmaxf(fadd(freq, fmul(cv[(k0 + j) * L], exp_amt)), 0.0f), 0.9f);  // :213
#pragma unroll
      for (int j = 0; j < U; ++j) {
        moog_coef(fc[j], r, fj[j], pj[j], qj[j]);
        virgin = virgin && fc[j] == 0.0f && r == 0.0f;
        if (virgin) { fj[j] = 0.0f; pj[j] = 0.0f; qj[j] = 0.0f; }
      }
      if (!virgin) { c_freq = fc[U - 1]; c_res = r; }
      f = fj[U - 1]; p = pj[U - 1]; q = qj[U - 1];
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j) { fj[j] = f; pj[j] = p; qj[j] = q; }
    }
    // the ladder (:69-82): the only dependent chain
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const float in = fsub(a[j], fmul(qj[j], b4));
      float t1 = b1;
      b1 = fsub(fmul(fadd(in, b0), pj[j]), fmul(b1, fj[j]));
      const float t2 = b2;
      b2 = fsub(fmul(fadd(b1, t1), pj[j]), fmul(b2, fj[j]));
      t1 = b3;
      b3 = fsub(fmul(fadd(b2, t2), pj[j]), fmul(b3, fj[j]));
      b4 = fsub(fmul(fadd(b3, t1), pj[j]), fmul(b4, fj[j]));
      b4 = fsub(b4, fmul(fmul(fmul(b4, b4), b4), 0.166667f));  // powi(3)
      b0 = clamp1(in);
      b1 = clamp1(b1); b2 = clamp1(b2); b3 = clamp1(b3); b4 = clamp1(b4);
      in_[j] = in; o3[j] = b3; o4[j] = b4;
    }
    // calc returns (b4, in - b4, 3*(b3-b4)) assigned to (lowpass, highpass, bandpass), :211
    if (lowpass) {
#pragma unroll
      for (int j = 0; j < U; ++j) lowpass[(k0 + j) * L] = o4[j];
    }
    if (highpass) {
#pragma unroll
      for (int j = 0; j < U; ++j) highpass[(k0 + j) * L] = fsub(in_[j], o4[j]);
    }
    if (bandpass) {
#pragma unroll
      for (int j = 0; j < U; ++j) bandpass[(k0 + j) * L] = fmul(3.0f, fsub(o3[j], o4[j]));
    }
  });
  s[0] = __float_as_uint(f); s[L] = __float_as_uint(p); s[2 * L] = __float_as_uint(q);
  s[3 * L] = __float_as_uint(b0); s[4 * L] = __float_as_uint(b1); s[5 * L] = __float_as_uint(b2);
  s[6 * L] = __float_as_uint(b3); s[7 * L] = __float_as_uint(b4);
  s[8 * L] = __float_as_uint(c_freq); s[9 * L] = __float_as_uint(c_res);
}

__device__ __forceinline__ void op_moog_dispatch(const Instr& ins, const Lane& ln, int kk) {
  const bool au = ins.in[0] >= 0, cv = ins.in[1] >= 0;
  if (au) {
    if (cv) op_moog<true, true>(ins, ln, kk);
    else op_moog<true, false>(ins, ln, kk);
  } else {
    if (cv) op_moog<false, true>(ins, ln, kk);
    else op_moog<false, false>(ins, ln, kk);
  }
}

// ---- ADSRModule::calc, src/synth/adsr.rs:134-217 ------------------------------
__device__ __forceinline__ void op_adsr(const Instr& ins, const Lane& ln, int kk) {
  uint32_t* s = ln.st + ins.state * L;
  float phase = __uint_as_float(s[0]), r_val = __uint_as_float(s[L]), from_a_val = __uint_as_float(s[2 * L]);
  uint32_t mode = s[3 * L] & 0xFFu;
  bool last = (s[3 * L] >> 8) & 1u;
  const uint32_t* pp = ln.pr + ins.param * L;
  const float a_sec = __uint_as_float(pp[0]), d_sec = __uint_as_float(pp[L]);
  const float s_val = __uint_as_float(pp[2 * L]), r_sec = __uint_as_float(pp[3 * L]);
  const float sr = ins.imm;
  // `1.0 / (self.sample_rate * self.x_sec)` is loop invariant: same IEEE value every sample
  const float inc_a = __fdiv_rn(1.0f, fmul(sr, a_sec));
  const float inc_d = __fdiv_rn(1.0f, fmul(sr, d_sec));
  const float inc_r = __fdiv_rn(1.0f, fmul(sr, r_sec));
  const float one_minus_s = fsub(1.0f, s_val);
  const float* gate = wire(ln, ins.in[0]);
  float* out = wire(ln, ins.out[0]);
  for_groups(kk, [&](auto u, int k0) {
    constexpr int U = decltype(u)::value;
    float g[U], o[U];
#pragma unroll
    for (int j = 0; j < U; ++j) g[j] = gate ? gate[(k0 + j) * L] : 0.0f;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const bool high = gate && g[j] > 0.0f;   // `gate.is_some() && gate[i] > 0.0`
      const bool low = !gate || g[j] <= 0.0f;  // `gate.is_none() || gate[i] <= 0.0` (NaN is neither)
      const bool tr = transition(last, g[j]);  // None => detector sees 0.0
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
        if (low) { phase = 0.0f; mode = ADSR_RELEASE; }
        if (tr) { phase = 0.0f; mode = ADSR_ATTACK; }
      } else {  // Release
        if (high) { phase = 0.0f; mode = ADSR_ATTACK; }
        phase = fadd(phase, inc_r);
        if (phase >= 1.0f) { phase = 0.0f; r_val = 0.0f; mode = ADSR_NONE; }
      }
      float v;
      if (mode == ADSR_NONE) v = 0.0f;
      else if (mode == ADSR_ATTACK) v = fadd(r_val, fmul(fsub(1.0f, r_val), phase));
      else if (mode == ADSR_DECAY) v = fadd(s_val, fmul(one_minus_s, fsub(1.0f, phase)));
      else if (mode == ADSR_SUSTAIN) v = s_val;
      else v = fmul(s_val, fsub(1.0f, phase));
      o[j] = v;
      if (mode != ADSR_ATTACK) r_val = v; else from_a_val = v;
    }
    if (out) {
#pragma unroll
      for (int j = 0; j < U; ++j) out[(k0 + j) * L] = o[j];
    }
  });
  s[0] = __float_as_uint(phase); s[L] = __float_as_uint(r_val); s[2 * L] = __float_as_uint(from_a_val);
  s[3 * L] = mode | (last ? 1u << 8 : 0u);
}

// ---- VCAModule::calc, src/synth/vca.rs:117-148 ---------------------------------
__device__ __forceinline__ void op_vca(const Instr& ins, const Lane& ln, int kk) {
  const float* audio = wire(ln, ins.in[0]);
  const float* cv = wire(ln, ins.in[1]);
  float* out = wire(ln, ins.out[0]);
  if (!out) return;
  const bool negative = __uint_as_float(ln.pr[ins.param * L]) != 0.0f;
  if (!(audio && cv)) {  // :143 `_ => output.fill(0.0)`
    for (int k = 0; k < kk; ++k) out[k * L] = 0.0f;
    return;
  }
  for_groups(kk, [&](auto u, int k0) {
    constexpr int U = decltype(u)::value;
    float a[U], c[U];
#pragma unroll
    for (int j = 0; j < U; ++j) { a[j] = audio[(k0 + j) * L]; c[j] = cv[(k0 + j) * L]; }
#pragma unroll
    for (int j = 0; j < U; ++j) out[(k0 + j) * L] = (negative || c[j] > 0.0f) ? fmul(a[j], c[j]) : 0.0f;
  });
}

// ---- MonoMixerModule::calc, src/synth/mixer.rs:101-122 -------------------------
__device__ __forceinline__ void op_mixer(const Instr& ins, const Lane& ln, int kk) {
  float* out = wire(ln, ins.out[0]);
  if (!out) return;
  const uint32_t* pp = ln.pr + ins.param * L;
  const float* in[4];
  float gain[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    in[j] = wire(ln, ins.in[j]);
    gain[j] = __uint_as_float(pp[j * L]);
  }
  for_groups(kk, [&](auto u, int k0) {
    constexpr int U = decltype(u)::value;
    float o[U];  // output.fill(0.0) then `*dst += src * gain` per connected input, in order
#pragma unroll
    for (int j = 0; j < U; ++j) o[j] = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (in[i]) {
#pragma unroll
        for (int j = 0; j < U; ++j) o[j] = fadd(o[j], fmul(in[i][(k0 + j) * L], gain[i]));
      }
#pragma unroll
    for (int j = 0; j < U; ++j) out[(k0 + j) * L] = o[j];
  });
}

// ---- MathModule / NonLinearModule::calc, src/synth/math.rs:139-160, :292-313 ----
template <int WHICH>
__device__ __forceinline__ float math_op(float a, float b) {
  if (WHICH == F_MATH_ADD) return fadd(a, b);
  if (WHICH == F_MATH_SUB) return fsub(a, b);
  if (WHICH == F_MATH_MUL) return fmul(a, b);
  // math.rs:203-205 `if a > 0.0 { a.powf(b) } else { -(-a).powf(b) }` in f32.  glibc's powf
  // evaluates in f64 and rounds once; f64 pow here then one rounding agrees with it except
  // when the f64 results straddle an f32 rounding boundary.
  return a > 0.0f ? __double2float_rn(pow((double)a, (double)b)) : -__double2float_rn(pow((double)(-a), (double)b));
}

template <int WHICH>
__device__ __forceinline__ void op_math(const Instr& ins, const Lane& ln, int kk) {
  float* out = wire(ln, ins.out[0]);
  if (!out) return;
  const float constant = __uint_as_float(ln.pr[ins.param * L]);
  const float* i1 = wire(ln, ins.in[0]);
  const float* i2 = wire(ln, ins.in[1]);
  for_groups(kk, [&](auto u, int k0) {
    constexpr int U = decltype(u)::value;
    float a[U], b[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      a[j] = i1 ? i1[(k0 + j) * L] : 0.0f;      // (None, _) => 0.0
      b[j] = i2 ? i2[(k0 + j) * L] : constant;  // (_, None) => constant
    }
#pragma unroll
    for (int j = 0; j < U; ++j) out[(k0 + j) * L] = math_op<WHICH>(a[j], b[j]);
  });
}

__device__ __forceinline__ void op_math_dispatch(const Instr& ins, const Lane& ln, int kk) {
  switch (ins.flags) {
    case F_MATH_ADD: op_math<F_MATH_ADD>(ins, ln, kk); break;
    case F_MATH_SUB: op_math<F_MATH_SUB>(ins, ln, kk); break;
    case F_MATH_MUL: op_math<F_MATH_MUL>(ins, ln, kk); break;
    default: op_math<F_MATH_NONLIN>(ins, ln, kk); break;
  }
}

}  // namespace dsp
}  // namespace srk
