// Per-module DSP on REGISTERS, one voice per lane, for patch-specialised ("fused") voice kernels.
//
// fused_gen.cpp turns a planned patch into a short piece of CUDA C++ -- one object of the op
// templates below per module, one call per module and sample group, wires as local arrays -- that
// NVRTC (or nvcc, ahead of time) compiles for sm_100a together with this header.  All arithmetic is
// here, hand-written; the generated part is the wiring.  Against the interpreter (dsp.cuh +
// voice_kernel.cuh) a fused kernel has no wire tiles in shared memory, no per-chunk state traffic
// and no dispatch: module state lives in registers for the whole render, every wire is a register,
// and what limits it is instruction issue.
//
// Arithmetic contract as in dsp.cuh: explicit round-to-nearest intrinsics everywhere (nothing can
// be contracted into an FMA), IEEE division, denormals kept; bit-for-bit with the reference except
// where the reference itself goes through the platform libm (f64 sin, pow).
//
// No standard headers: NVRTC has none.  Everything is inside namespace fz.
#pragma once
#include "fused_args.h"
#include "libm_glibc.cuh"

namespace fz {

typedef unsigned int u32;
typedef int i32;
typedef unsigned long long u64;

#define FZ_DEV __device__ __forceinline__

// (macros, not functions: with -lineinfo an inlined helper's instructions carry the HELPER's line, and the per-op
// attribution of an ncu capture -- scripts/ncu_fused_summary.py -- needs the line of the op body that used it)
#define fadd(a, b) __fadd_rn((a), (b))
#define fsub(a, b) __fsub_rn((a), (b))
#define fmul(a, b) __fmul_rn((a), (b))
#define dadd(a, b) __dadd_rn((a), (b))
#define dsub(a, b) __dsub_rn((a), (b))
#define dmul(a, b) __dmul_rn((a), (b))
// A value the compiler must treat as unknown.  Ops whose output is the same for every sample of a group (a Math
// module with no inputs is how a patch spells a constant) write it through this: nvcc 12.9 mis-folds nested selects
// over complementary predicates of ONE value at -O2 and above (select(x > 0, select(x <= 0, a, b), c) came out as
// select(x > 0, a, 0) in a transition detector fed by a constant: tests/test_fused.py keeps the case).  No instruction.
FZ_DEV float opaque(float x) {
  asm volatile("" : "+f"(x));
  return x;
}
FZ_DEV float asf(u32 x) { return __uint_as_float(x); }
FZ_DEV u32 asu(float x) { return __float_as_uint(x); }

// This lane's voice.
struct Ctx {
  const SrkFusedArgs* a;
  u32 v;         // voice inside this launch (idle lanes shadow the last voice and never store)
  u32 lane;
  u32 group;     // voice group of 32: this warp (single-stage kernels) or this warp and its S - 1 siblings
  u32 stage;     // pipeline stage this warp runs (0 in single-stage kernels)
  u32 gib;       // group index inside the thread block
  u32 n_active;  // voices of the group that exist
  bool active;
  FZ_DEV u32 ld_state(u32 w) const { return a->state[(size_t)w * a->V + v]; }
  FZ_DEV void st_state(u32 w, u32 x) const { if (active) a->state[(size_t)w * a->V + v] = x; }
  FZ_DEV u32 ld_param(u32 w) const { return a->params[(size_t)w * a->V + v]; }
};

// f64 `x % 1.0` (Rust) == fmod(x, 1.0): x - trunc(x) is exact for finite x, NaN for +-inf.
static __device__ __noinline__ double fmod1_exact(double x) { return dsub(x, trunc(x)); }
// fmod(x, 1.0) for x in [0, 2).  (x - (x >= 1 ? 1 : 0) is one select fewer but puts the compare IN the phase
// recurrence's dependent chain instead of beside the subtraction: cfg4 @ 16384 13.1 -> 14.7 ms, profiles/r06f_tune_all.txt)
FZ_DEV double wrap01(double x) { return x >= 1.0 ? dsub(x, 1.0) : x; }

// OscillatorModule::poly_blep, src/synth/oscillator.rs:50-67 (select-based: one division serves either arm)
FZ_DEV double blep_eval(double t, double dt, double one_minus_dt) {
  const bool lo = t < dt;
  const bool hi = !lo & (t > one_minus_dt);
  const double q = __ddiv_rn(lo ? t : dsub(t, 1.0), dt);
  const double qq = dmul(q, q);
  const double r_lo = dsub(dsub(dadd(q, q), qq), 1.0);
  const double r_hi = dadd(dadd(dadd(qq, q), q), 1.0);
  const double r = lo ? r_lo : (hi ? r_hi : 0.0);
  return dt == 0.0 ? 0.0 : r;
}

// ---- OscillatorModule::calc, src/synth/oscillator.rs:108-158 ------------------------------------
// HAS_CV: delta = 440 * 2^(cv + val) / sr per sample (:43-48, :132), else the voice's constant (computed on the
// host in f64 with glibc, bit-identical to the reference).  OUTS: bit 0 sine, 1 square, 2 saw are read by somebody.
// a / b correctly rounded, given y = RN(1 / b) (__drcp_rn), for normal operands whose quotient and residuals stay
// normal: one multiply, one refinement that makes the quotient faithful, then Markstein's final step (the residual
// a - b q is exact in an FMA; with a correctly rounded reciprocal and a faithful q, RN(q + r y) = RN(a / b)).
// Five instructions on the f64 pipe and no branch, against __ddiv_rn's reciprocal, Newton steps, range test and
// slow-path call.  Only used where b is constant over a render (the oscillator's delta); tests/c/ddiv_markstein.c
// checks it against IEEE division on 1e8 operand pairs of the ranges that occur.
FZ_DEV double ddiv_by_const(double a, double b, double y) {
  const double q0 = dmul(a, y);
  const double q1 = __fma_rn(__fma_rn(-b, q0, a), y, q0);
  return __fma_rn(__fma_rn(-b, q1, a), y, q1);
}
// poly_blep with that division (same selects as blep_eval; dt > 0 here)
FZ_DEV double blep_eval_const(double t, double dt, double one_minus_dt, double rdt) {
  const bool lo = t < dt;
  const bool hi = !lo & (t > one_minus_dt);
  const double q = ddiv_by_const(lo ? t : dsub(t, 1.0), dt, rdt);
  const double qq = dmul(q, q);
  const double r_lo = dsub(dsub(dadd(q, q), qq), 1.0);
  const double r_hi = dadd(dadd(dadd(qq, q), q), 1.0);
  return lo ? r_lo : (hi ? r_hi : 0.0);
}

// RATE = 1 ("audio rate", chosen by the host, which knows every voice's delta): constant delta with 2^-200 <= delta <
// 1/8 for EVERY voice and large enough that most groups have some lane next to a discontinuity anyway -- the group
// test is dropped and every group runs the branch-free path, so that the oscillator shares one basic block with
// the modules around it (what limits a fused kernel is dependent-instruction latency: profiles/r04b).
template <bool HAS_CV, bool HAS_SYNC, int OUTS, bool AA, int RATE>
struct Osc {
  double pos, d0, val, sr;
  double om0, h1, h2, rd0;  // fast-path bounds and RN(1 / d0) for a constant delta (see run)
  double rsr;               // RN(1 / sample_rate)
  bool last, small;

  FZ_DEV void load(const Ctx& c, u32 sw, u32 val_bits, u32 d_lo, u32 d_hi, float sample_rate) {
    pos = __hiloint2double((int)c.ld_state(sw + 1), (int)c.ld_state(sw));
    last = c.ld_state(sw + 2) != 0u;
    val = (double)asf(val_bits);
    d0 = __hiloint2double((int)d_hi, (int)d_lo);
    sr = (double)sample_rate;
    rsr = __drcp_rn(sr);
    om0 = dsub(1.0, d0);
    // square: (pos + 0.5) % 1.0 stays clear of both ends while pos <= h1 (< 0.5) or pos >= h2 (> 0.5).
    // h1 + 0.5 == om0 exactly, so fl(pos + 0.5) <= om0 for every pos <= h1 (rounding is monotonic);
    // above, fl(pos + 0.5) - 1 >= 2 d0 - 2^-52 >= d0 as long as d0 >= 2^-50 (else h2 is unreachable).
    h1 = dsub(om0, 0.5);
    h2 = d0 >= 0x1p-50 ? dadd(0.5, dmul(2.0, d0)) : 2.0;
    // the branch-free path needs at most one wrap per group and a division whose intermediates stay normal
    small = (d0 >= 0x1p-200) & (d0 < 0.125);
    rd0 = __drcp_rn(d0);
  }
  FZ_DEV void store(const Ctx& c, u32 sw) const {
    c.st_state(sw, (u32)__double2loint(pos));
    c.st_state(sw + 1, (u32)__double2hiint(pos));
    c.st_state(sw + 2, last ? 1u : 0u);
  }
  FZ_DEV bool needs_generic() const { return false; }

  template <int U>
  FZ_DEV void shape(const double (&ps)[U], const double (&dl)[U], float* sine, float* square, float* saw) {
    constexpr bool SINE = OUTS & 1, SQUARE = OUTS & 2, SAW = OUTS & 4;
    if (SINE) sine_port(ps, sine, 1);  // (float) of glibc's f64 sin, libm_glibc.cuh
    if (SQUARE || SAW) {
      double om[U], p2[U];
      bool near = false;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        om[j] = dsub(1.0, dl[j]);
        near |= (ps[j] < dl[j]) | (ps[j] > om[j]);
        if (SQUARE) {
          p2[j] = wrap01(dadd(ps[j], 0.5));  // (pos + 0.5) % 1.0, pos in [0, 1)
          near |= (p2[j] < dl[j]) | (p2[j] > om[j]);
        }
      }
      if (AA && __any_sync(0xFFFFFFFFu, near)) {
        double pb0[U];
#pragma unroll
        for (int j = 0; j < U; ++j) pb0[j] = blep_eval(ps[j], dl[j], om[j]);
        if (SQUARE) {
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const float base = ps[j] < 0.5 ? -1.0f : 1.0f;
            square[j] = fsub(base, __double2float_rn(dsub(pb0[j], blep_eval(p2[j], dl[j], om[j]))));
          }
        }
        if (SAW) {
#pragma unroll
          for (int j = 0; j < U; ++j)
            saw[j] = fsub(fsub(fmul(__double2float_rn(ps[j]), 2.0f), 1.0f), __double2float_rn(pb0[j]));
        }
      } else {  // every correction is 0.0, and x - 0.0f == x bit for bit
        if (SQUARE) {
#pragma unroll
          for (int j = 0; j < U; ++j) square[j] = ps[j] < 0.5 ? -1.0f : 1.0f;
        }
        if (SAW) {
#pragma unroll
          for (int j = 0; j < U; ++j) saw[j] = fsub(fmul(__double2float_rn(ps[j]), 2.0f), 1.0f);
        }
      }
    }
  }

  // Constant delta in [2^-200, 1/8), no sync: straight-line code, per-sample wrap (pos + d0 < 1.125), polyBLEP through
  // the constant-divisor division.  No vote, no call, no branch.
  template <int U>
  FZ_DEV void run_small(float* sine, float* square, float* saw) {
    constexpr bool SINE = OUTS & 1, SQUARE = OUTS & 2, SAW = OUTS & 4;
    double ps[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      ps[j] = pos;
      pos = wrap01(dadd(pos, d0));
    }
    if (SINE) sine_port(ps, sine, 1);  // (float) of glibc's f64 sin, libm_glibc.cuh
    if (SQUARE || SAW) {
      double pb0[U];
      if (AA) {
#pragma unroll
        for (int j = 0; j < U; ++j) pb0[j] = blep_eval_const(ps[j], d0, om0, rd0);
      }
      if (SQUARE) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const float base = ps[j] < 0.5 ? -1.0f : 1.0f;
          if (AA) {
            const double p2 = wrap01(dadd(ps[j], 0.5));
            square[j] = fsub(base, __double2float_rn(dsub(pb0[j], blep_eval_const(p2, d0, om0, rd0))));
          } else {
            square[j] = base;
          }
        }
      }
      if (SAW) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const float ramp = fsub(fmul(__double2float_rn(ps[j]), 2.0f), 1.0f);
          saw[j] = AA ? fsub(ramp, __double2float_rn(pb0[j])) : ramp;
        }
      }
    }
  }

  template <int U, bool FAST>
  FZ_DEV void run(const float* cv, const float* sync, float* sine, float* square, float* saw) {
    constexpr bool SINE = OUTS & 1, SQUARE = OUTS & 2, SAW = OUTS & 4;
    constexpr bool NEAR = AA && (SQUARE || SAW);
    double ps[U], dl[U];
    if (!HAS_CV && !HAS_SYNC && RATE == 1) {
      last = false;
      run_small<U>(sine, square, saw);
      return;
    }
    if (!HAS_CV && !HAS_SYNC) {
      // Constant delta, no sync: the phases of the group are pos, pos + d, ... as long as none of them wraps,
      // and when in addition no sample sits within d of a discontinuity every polyBLEP term is 0.0.  One
      // test per group (on the first and the last phase: they are monotonic) instead of per sample.
      double x = pos;
#pragma unroll
      for (int j = 0; j < U; ++j) { ps[j] = x; x = dadd(x, d0); }
      bool plain;
      if (NEAR) {
        plain = (ps[0] >= d0) & (ps[U - 1] <= om0);  // also: d0 <= ps[0] < 1, so x < 2
        if (SQUARE) plain &= (ps[U - 1] <= h1) | (ps[0] >= h2);
      } else {
        plain = (ps[U - 1] < 1.0) & (x < 2.0);
      }
      last = false;  // with no sync input the detector sees 0.0 every sample
      if (__all_sync(0xFFFFFFFFu, plain)) {
        pos = wrap01(x);
        if (SINE) sine_port(ps, sine, 1);
        if (SQUARE) {
          const float base = ps[0] < 0.5 ? -1.0f : 1.0f;  // the whole group is on one side of 0.5
#pragma unroll
          for (int j = 0; j < U; ++j) square[j] = NEAR ? base : (ps[j] < 0.5 ? -1.0f : 1.0f);
        }
        if (SAW) {
#pragma unroll
          for (int j = 0; j < U; ++j) saw[j] = fsub(fmul(__double2float_rn(ps[j]), 2.0f), 1.0f);
        }
        return;
      }
      if (__all_sync(0xFFFFFFFFu, small)) {
        run_small<U>(sine, square, saw);
        return;
      }
    }
    // general path: per-sample wrap, sync reset, per-sample delta
    bool edge[U];
#pragma unroll
    for (int j = 0; j < U; ++j) edge[j] = false;
    if (HAS_SYNC) {
      bool prev = last;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const bool above = sync[j] > 0.0f;
        edge[j] = above && !prev;
        prev = above;
      }
      last = prev;
    } else {
      last = false;
    }
    if (HAS_CV) {
      // delta = 440 * 2^(cv + val) / sample_rate (:43-48, :132).  The divisor is constant over the render: Markstein's
      // division by a constant (five f64 instructions, no branch) whenever every numerator of the group is a normal number
      // whose quotient stays normal, IEEE division otherwise (2^x under- or overflowed, or the CV was not finite)
      double num[U], oct[U], pw[U];
      bool tame = true;
#pragma unroll
      for (int j = 0; j < U; ++j) oct[j] = dadd((double)cv[j], val);
      exp2_glibc_group(oct, pw);
#pragma unroll
      for (int j = 0; j < U; ++j) {
        num[j] = dmul(440.0, pw[j]);
        const u32 ex = ((u32)__double2hiint(num[j]) >> 20) & 0x7ffu;
        tame &= (ex - 64u) < 1920u;  // biased exponent in [64, 1984)
      }
      if (__all_sync(0xFFFFFFFFu, tame)) {
#pragma unroll
        for (int j = 0; j < U; ++j) dl[j] = ddiv_by_const(num[j], sr, rsr);
      } else {
#pragma unroll
        for (int j = 0; j < U; ++j) dl[j] = __ddiv_rn(num[j], sr);
      }
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j) dl[j] = d0;
    }
    const double pos0 = pos;
    bool odd = false;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (HAS_SYNC) pos = edge[j] ? 0.0 : pos;
      ps[j] = pos;
      const double x = dadd(pos, dl[j]);
      odd |= !(x < 2.0);  // NaN, inf or a step of more than one period: exact path below
      pos = wrap01(x);
    }
    if (__any_sync(0xFFFFFFFFu, odd)) {
      pos = pos0;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (HAS_SYNC) pos = edge[j] ? 0.0 : pos;
        ps[j] = pos;
        pos = fmod1_exact(dadd(pos, dl[j]));
      }
    }
    shape<U>(ps, dl, sine, square, saw);
  }
};

// ---- NoiseModule::calc, src/synth/oscillator.rs:381-388 (seeded Philox4x32-10 stand-in) ----------
FZ_DEV void philox4x32_10(u32 (&c)[4], u32 k0, u32 k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const u32 hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const u32 hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const u32 n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

struct Noise {
  u64 n;
  u32 voice, module, k0, k1;
  FZ_DEV void load(const Ctx& c, u32 sw, u32 module_index) {
    n = ((u64)c.ld_state(sw + 1) << 32) | c.ld_state(sw);
    voice = c.a->voice_offset + c.v;
    module = module_index;
    k0 = c.a->seed_lo;
    k1 = c.a->seed_hi;
  }
  FZ_DEV void store(const Ctx& c, u32 sw) const {
    c.st_state(sw, (u32)n);
    c.st_state(sw + 1, (u32)(n >> 32));
  }
  FZ_DEV bool needs_generic() const { return (n & 3u) != 0u; }  // (the same for every voice)
  FZ_DEV static float shape(u32 r) {
    const float u = fmul((float)(r >> 8), 1.0f / 16777216.0f);  // rand 0.8.5 Standard f32
    return fmul(fsub(u, 0.5f), 2.0f);
  }
  FZ_DEV void draw(u64 blk, u32 (&c)[4]) const {
    c[0] = (u32)blk; c[1] = (u32)(blk >> 32); c[2] = voice; c[3] = module;
    philox4x32_10(c, k0, k1);
  }
  template <int U, bool FAST>
  FZ_DEV void run(float* out) {
    if (out) {
      u32 c[4];
      if (FAST && U % 4 == 0) {  // one Philox block per 4 samples: the FAST body runs only while the counter is a multiple of 4
#pragma unroll
        for (int j = 0; j < U; j += 4) {
          draw((n + j) >> 2, c);
          out[j] = shape(c[0]); out[j + 1] = shape(c[1]); out[j + 2] = shape(c[2]); out[j + 3] = shape(c[3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          draw((n + j) >> 2, c);
          const u32 q = (u32)((n + j) & 3u);
          out[j] = shape(q == 0 ? c[0] : q == 1 ? c[1] : q == 2 ? c[2] : c[3]);
        }
      }
    }
    n += U;
  }
};

// ---- MoogFilterModule::calc, src/synth/filter.rs:182-221, :60-91 ---------------------------------
FZ_DEV float clamp1(float x) { return fmaxf(fminf(x, 1.0f), -1.0f); }
FZ_DEV void moog_coef(float fc, float r, float& f, float& p, float& q) {  // :61-68
  q = fsub(1.0f, fc);
  p = fadd(fc, fmul(fmul(0.8f, fc), q));
  f = fsub(fmul(p, 2.0f), 1.0f);
  q = fmul(r, fadd(1.0f, fmul(fmul(0.5f, q), fadd(fsub(1.0f, q), fmul(fmul(5.6f, q), q)))));
}

// The cached coefficients always equal moog_coef(current fc, r), except while the state is still the all-zero
// Default and (fc, r) == (0, 0) hits that zero cache (`virgin`), where they stay 0 (see dsp.cuh).
template <bool HAS_AUDIO, bool HAS_CV, int OUTS>
struct Moog {
  float f, p, q, b0, b1, b2, b3, b4, c_freq, c_res;
  float freq, r, exp_amt;
  bool virgin;

  FZ_DEV void load(const Ctx& c, u32 sw, u32 freq_bits, u32 res_bits, u32 exp_bits) {
    f = asf(c.ld_state(sw)); p = asf(c.ld_state(sw + 1)); q = asf(c.ld_state(sw + 2));
    b0 = asf(c.ld_state(sw + 3)); b1 = asf(c.ld_state(sw + 4)); b2 = asf(c.ld_state(sw + 5));
    b3 = asf(c.ld_state(sw + 6)); b4 = asf(c.ld_state(sw + 7));
    c_freq = asf(c.ld_state(sw + 8)); c_res = asf(c.ld_state(sw + 9));
    freq = asf(freq_bits);
    r = fminf(fmaxf(asf(res_bits), 0.0f), 1.0f);  // :214
    exp_amt = asf(exp_bits);
    if (!HAS_CV && c.a->n_samples > 0) {  // constant cutoff: one cache check for the whole render (:61)
      const float fc = fminf(fmaxf(fadd(freq, fmul(0.0f, exp_amt)), 0.0f), 0.9f);  // :213 with cv = 0.0
      if (fc != c_freq || r != c_res) {
        c_freq = fc;
        c_res = r;
        moog_coef(fc, r, f, p, q);
      }
    }
    virgin = (c_freq == 0.0f) & (c_res == 0.0f) & (f == 0.0f);
  }
  FZ_DEV void store(const Ctx& c, u32 sw) const {
    c.st_state(sw, asu(f)); c.st_state(sw + 1, asu(p)); c.st_state(sw + 2, asu(q));
    c.st_state(sw + 3, asu(b0)); c.st_state(sw + 4, asu(b1)); c.st_state(sw + 5, asu(b2));
    c.st_state(sw + 6, asu(b3)); c.st_state(sw + 7, asu(b4));
    c.st_state(sw + 8, asu(c_freq)); c.st_state(sw + 9, asu(c_res));
  }

  FZ_DEV bool needs_generic() const { return HAS_CV && __any_sync(0xFFFFFFFFu, virgin); }
  template <int U, bool FAST>
  FZ_DEV void run(const float* audio, const float* cv, float* lowpass, float* bandpass, float* highpass) {
    float fj[U], pj[U], qj[U], in_[U], o3[U], o4[U];
    if (HAS_CV) {
      float fc[U];
#pragma unroll
      for (int j = 0; j < U; ++j) fc[j] = fminf(fmaxf(fadd(freq, fmul(cv[j], exp_amt)), 0.0f), 0.9f);  // :213
#pragma unroll
      for (int j = 0; j < U; ++j) moog_coef(fc[j], r, fj[j], pj[j], qj[j]);
      if (!FAST && __any_sync(0xFFFFFFFFu, virgin)) {  // only until (fc, r) first leaves (0, 0); FAST tiles have no virgin lane
#pragma unroll
        for (int j = 0; j < U; ++j) {
          virgin = virgin & (fc[j] == 0.0f) & (r == 0.0f);
          if (virgin) { fj[j] = 0.0f; pj[j] = 0.0f; qj[j] = 0.0f; }
        }
      }
      if (FAST || !virgin) { c_freq = fc[U - 1]; c_res = r; }
      f = fj[U - 1]; p = pj[U - 1]; q = qj[U - 1];
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j) { fj[j] = f; pj[j] = p; qj[j] = q; }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {  // :69-82
      const float in = fsub(HAS_AUDIO ? audio[j] : 0.0f, fmul(qj[j], b4));
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
    if (OUTS & 1) {
#pragma unroll
      for (int j = 0; j < U; ++j) lowpass[j] = o4[j];
    }
    if (OUTS & 4) {
#pragma unroll
      for (int j = 0; j < U; ++j) highpass[j] = fsub(in_[j], o4[j]);
    }
    if (OUTS & 2) {
#pragma unroll
      for (int j = 0; j < U; ++j) bandpass[j] = fmul(3.0f, fsub(o3[j], o4[j]));
    }
  }
};

// The ladder filter as two ops, for kernels cut into pipeline stages: the coefficient block (:61-68; a pure function of
// the CV, 14 instructions per sample, none of them on the ladder's dependent chain) can then run in ANOTHER stage -- the
// ladder's own stage is the slowest of every BASELINE patch (cfg2 @ 4096 voices: 2.80 ms with the coefficients inside
// it, 2.42 ms without them, profiles/r05i) -- and travels as three wires (f, p, q).  Same arithmetic, same state words:
// MoogCoef owns f, p, q and the cache key (c_freq, c_res), MoogCore the delay line b0..b4.
struct MoogCoef {
  float f, p, q, c_freq, c_res;
  float freq, r, exp_amt;
  bool virgin;
  FZ_DEV void load(const Ctx& c, u32 sw, u32 freq_bits, u32 res_bits, u32 exp_bits) {
    f = asf(c.ld_state(sw)); p = asf(c.ld_state(sw + 1)); q = asf(c.ld_state(sw + 2));
    c_freq = asf(c.ld_state(sw + 8)); c_res = asf(c.ld_state(sw + 9));
    freq = asf(freq_bits);
    r = fminf(fmaxf(asf(res_bits), 0.0f), 1.0f);  // :214
    exp_amt = asf(exp_bits);
    virgin = (c_freq == 0.0f) & (c_res == 0.0f) & (f == 0.0f);
  }
  FZ_DEV void store(const Ctx& c, u32 sw) const {
    c.st_state(sw, asu(f)); c.st_state(sw + 1, asu(p)); c.st_state(sw + 2, asu(q));
    c.st_state(sw + 8, asu(c_freq)); c.st_state(sw + 9, asu(c_res));
  }
  FZ_DEV bool needs_generic() const { return __any_sync(0xFFFFFFFFu, virgin); }
  template <int U, bool FAST>
  FZ_DEV void run(const float* cv, float* fo, float* po, float* qo) {
    float fc[U];
#pragma unroll
    for (int j = 0; j < U; ++j) fc[j] = fminf(fmaxf(fadd(freq, fmul(cv[j], exp_amt)), 0.0f), 0.9f);  // :213
#pragma unroll
    for (int j = 0; j < U; ++j) moog_coef(fc[j], r, fo[j], po[j], qo[j]);
    if (!FAST && __any_sync(0xFFFFFFFFu, virgin)) {
#pragma unroll
      for (int j = 0; j < U; ++j) {
        virgin = virgin & (fc[j] == 0.0f) & (r == 0.0f);
        if (virgin) { fo[j] = 0.0f; po[j] = 0.0f; qo[j] = 0.0f; }
      }
    }
    if (FAST || !virgin) { c_freq = fc[U - 1]; c_res = r; }
    f = fo[U - 1]; p = po[U - 1]; q = qo[U - 1];
  }
};

template <bool HAS_AUDIO, int OUTS>
struct MoogCore {
  float b0, b1, b2, b3, b4;
  FZ_DEV void load(const Ctx& c, u32 sw) {
    b0 = asf(c.ld_state(sw + 3)); b1 = asf(c.ld_state(sw + 4)); b2 = asf(c.ld_state(sw + 5));
    b3 = asf(c.ld_state(sw + 6)); b4 = asf(c.ld_state(sw + 7));
  }
  FZ_DEV void store(const Ctx& c, u32 sw) const {
    c.st_state(sw + 3, asu(b0)); c.st_state(sw + 4, asu(b1)); c.st_state(sw + 5, asu(b2));
    c.st_state(sw + 6, asu(b3)); c.st_state(sw + 7, asu(b4));
  }
  template <int U, bool FAST>
  FZ_DEV void run(const float* audio, const float* fj, const float* pj, const float* qj, float* lowpass, float* bandpass, float* highpass) {
    float in_[U], o3[U], o4[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {  // :69-82
      const float in = fsub(HAS_AUDIO ? audio[j] : 0.0f, fmul(qj[j], b4));
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
    if (OUTS & 1) {
#pragma unroll
      for (int j = 0; j < U; ++j) lowpass[j] = o4[j];
    }
    if (OUTS & 4) {
#pragma unroll
      for (int j = 0; j < U; ++j) highpass[j] = fsub(in_[j], o4[j]);
    }
    if (OUTS & 2) {
#pragma unroll
      for (int j = 0; j < U; ++j) bandpass[j] = fmul(3.0f, fsub(o3[j], o4[j]));
    }
  }
};

// ---- ADSRModule::calc, src/synth/adsr.rs:134-217 --------------------------------------------------
enum : u32 { ADSR_ATTACK = 0, ADSR_DECAY = 1, ADSR_SUSTAIN = 2, ADSR_RELEASE = 3, ADSR_NONE = 4 };

template <bool HAS_GATE>
struct Adsr {
  float phase, r_val, from_a_val;
  u32 mode;
  bool last;
  float s_val, inc_a, inc_d, inc_r, one_minus_s;

  FZ_DEV void load(const Ctx& c, u32 sw, u32 a_bits, u32 d_bits, u32 s_bits, u32 r_bits, float sr) {
    phase = asf(c.ld_state(sw)); r_val = asf(c.ld_state(sw + 1)); from_a_val = asf(c.ld_state(sw + 2));
    const u32 m = c.ld_state(sw + 3);
    mode = m & 0xFFu;
    last = (m >> 8) & 1u;
    s_val = asf(s_bits);
    // `1.0 / (self.sample_rate * self.x_sec)` is loop invariant: same IEEE value every sample
    inc_a = __fdiv_rn(1.0f, fmul(sr, asf(a_bits)));
    inc_d = __fdiv_rn(1.0f, fmul(sr, asf(d_bits)));
    inc_r = __fdiv_rn(1.0f, fmul(sr, asf(r_bits)));
    one_minus_s = fsub(1.0f, s_val);
  }
  FZ_DEV void store(const Ctx& c, u32 sw) const {
    c.st_state(sw, asu(phase)); c.st_state(sw + 1, asu(r_val)); c.st_state(sw + 2, asu(from_a_val));
    c.st_state(sw + 3, mode | (last ? 1u << 8 : 0u));
  }

  // One sample of the five-arm `match self.mode` (:144-200) as selects.
  FZ_DEV float step(float g) {
    const bool above = g > 0.0f;
    const bool high = HAS_GATE & above;         // `gate.is_some() && gate[i] > 0.0`
    const bool low = !HAS_GATE | (g <= 0.0f);   // `gate.is_none() || gate[i] <= 0.0` (NaN is neither)
    const bool tr = above & !last;              // TransitionDetector, None => sees 0.0
    last = above;
    const bool m_none = mode == ADSR_NONE, m_att = mode == ADSR_ATTACK, m_dec = mode == ADSR_DECAY;
    const bool m_sus = mode == ADSR_SUSTAIN, m_rel = mode == ADSR_RELEASE;
    const bool rel_retrig = m_rel & high;  // Release restarts the attack first, then still advances (:188-199)
    const float ph0 = rel_retrig ? 0.0f : phase;
    const float inc = m_att ? inc_a : m_dec ? inc_d : inc_r;
    const float ph1 = fadd(ph0, inc);
    const bool ge = ph1 >= 1.0f;
    u32 nm = mode;
    nm = (m_none & high) ? ADSR_ATTACK : nm;
    nm = (m_att & ge) ? ADSR_DECAY : nm;
    nm = m_dec ? (tr ? ADSR_ATTACK : ge ? ADSR_SUSTAIN : ADSR_DECAY) : nm;
    nm = m_sus ? (tr ? ADSR_ATTACK : low ? ADSR_RELEASE : ADSR_SUSTAIN) : nm;
    nm = m_rel ? (ge ? ADSR_NONE : rel_retrig ? ADSR_ATTACK : ADSR_RELEASE) : nm;
    const bool zero = (m_none & high) | ((m_att | m_dec) & (ge | tr)) | (m_sus & (low | tr)) | (m_rel & ge);
    const bool advance = m_att | m_dec | m_rel;
    const float np = zero ? 0.0f : advance ? ph1 : phase;
    r_val = (m_att & !ge & tr) ? from_a_val : r_val;
    r_val = (m_rel & ge) ? 0.0f : r_val;
    mode = nm;
    phase = np;
    const bool n_att = mode == ADSR_ATTACK;
    const float omp = fsub(1.0f, phase);
    const float lin = fadd(n_att ? r_val : s_val, fmul(n_att ? fsub(1.0f, r_val) : one_minus_s, n_att ? phase : omp));
    float v = lin;                              // Attack / Decay (:203-204)
    v = mode == ADSR_RELEASE ? fmul(s_val, omp) : v;
    v = mode == ADSR_SUSTAIN ? s_val : v;
    v = mode == ADSR_NONE ? 0.0f : v;
    r_val = n_att ? r_val : v;
    from_a_val = n_att ? v : from_a_val;
    return v;
  }

  // A group in which nothing happens to the envelope's mode (no gate edge it reacts to, no phase wrap): two or
  // three flops per sample, the same ones in the same order as step().  The voices of a group normally share
  // their gate, so the mode is tested once per WARP and each mode has its own few lines; anything else (modes
  // differ between lanes, an edge, a wrap) goes through step().  The tests are conservative: `sh` / `na` recognise
  // a gate that stays high after being high / never rises in this group.
  template <int U>
  FZ_DEV bool quiet(const float (&g)[U], float* o) {
    const u32 full = 0xFFFFFFFFu;
    const u32 m0 = __shfl_sync(full, mode, 0);
    if (!__all_sync(full, mode == m0)) return false;
    bool all_above = true, none_above = true;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const bool above = g[j] > 0.0f;
      all_above &= above;
      none_above &= !above;
    }
    const bool sh = HAS_GATE & all_above & last;  // high throughout, no rising edge, never low
    const bool na = none_above;                   // no sample above 0: no rising edge, no `high`
    if (m0 == ADSR_SUSTAIN) {
      if (!__all_sync(full, sh)) return false;
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = s_val;
      r_val = s_val;
      last = true;
      return true;
    }
    if (m0 == ADSR_NONE) {
      if (!__all_sync(full, na)) return false;
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = 0.0f;
      r_val = 0.0f;
      last = false;
      return true;
    }
    float ph[U];
    bool ge = false;
    const float inc = m0 == ADSR_ATTACK ? inc_a : m0 == ADSR_DECAY ? inc_d : inc_r;
    float acc = phase;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      acc = fadd(acc, inc);
      ph[j] = acc;
      ge |= acc >= 1.0f;
    }
    if (m0 == ADSR_RELEASE) {
      if (!__all_sync(full, !ge & na)) return false;
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = fmul(s_val, fsub(1.0f, ph[j]));
      r_val = o[U - 1];
      phase = acc;
      last = false;
      return true;
    }
    if (!__all_sync(full, !ge & (sh | na))) return false;
    if (m0 == ADSR_ATTACK) {
      const float span = fsub(1.0f, r_val);
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = fadd(r_val, fmul(span, ph[j]));
      from_a_val = o[U - 1];
    } else {  // Decay
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = fadd(s_val, fmul(one_minus_s, fsub(1.0f, ph[j])));
      r_val = o[U - 1];
    }
    phase = acc;
    last = all_above;  // (sh: stays high; na: the last sample is not above)
    return true;
  }

  template <int U, bool FAST>
  FZ_DEV void run(const float* gate, float* out) {
    float g[U], o[U];
#pragma unroll
    for (int j = 0; j < U; ++j) g[j] = HAS_GATE ? gate[j] : 0.0f;
    if (!quiet<U>(g, o)) {
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = step(g[j]);
    }
    if (out) {
#pragma unroll
      for (int j = 0; j < U; ++j) out[j] = o[j];
    }
  }
};

// ---- VCAModule::calc, src/synth/vca.rs:117-148 ---------------------------------------------------
template <bool BOTH, bool NEGATIVE>
struct Vca {
  template <int U, bool FAST>
  FZ_DEV void run(const float* audio, const float* cv, float* out) const {
    if (!out) return;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (BOTH) out[j] = (NEGATIVE | (cv[j] > 0.0f)) ? fmul(audio[j], cv[j]) : 0.0f;
      else out[j] = opaque(0.0f);  // :143 `_ => output.fill(0.0)`
    }
  }
};

// ---- MonoMixerModule::calc, src/synth/mixer.rs:101-122 -------------------------------------------
template <int CONNECTED>  // bit k: input k is connected
struct Mixer {
  float gain[4];
  FZ_DEV void load(u32 g0, u32 g1, u32 g2, u32 g3) { gain[0] = asf(g0); gain[1] = asf(g1); gain[2] = asf(g2); gain[3] = asf(g3); }
  template <int U, bool FAST>
  FZ_DEV void run(const float* in0, const float* in1, const float* in2, const float* in3, float* out) const {
    if (!out) return;
    float o[U];  // output.fill(0.0) then `*dst += src * gain` per connected input, in order
#pragma unroll
    for (int j = 0; j < U; ++j) o[j] = 0.0f;
    if (CONNECTED & 1) {
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = fadd(o[j], fmul(in0[j], gain[0]));
    }
    if (CONNECTED & 2) {
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = fadd(o[j], fmul(in1[j], gain[1]));
    }
    if (CONNECTED & 4) {
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = fadd(o[j], fmul(in2[j], gain[2]));
    }
    if (CONNECTED & 8) {
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = fadd(o[j], fmul(in3[j], gain[3]));
    }
#pragma unroll
    for (int j = 0; j < U; ++j) out[j] = CONNECTED ? o[j] : opaque(o[j]);
  }
};

// ---- MathModule / NonLinearModule::calc, src/synth/math.rs:139-160, :203-205, :292-313 ----------
static __device__ __noinline__ float nonlinear(float a, float b) {
  return a > 0.0f ? powf_glibc(a, b) : -powf_glibc(-a, b);  // glibc's powf, bit for bit (libm_glibc.cuh)
}
template <int WHICH /*0 add 1 sub 2 mul 3 non-linear*/, bool HAS_A, bool HAS_B>
struct Math {
  float constant;
  FZ_DEV void load(u32 bits) { constant = asf(bits); }
  template <int U, bool FAST>
  FZ_DEV void run(const float* i1, const float* i2, float* out) const {
    if (!out) return;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const float a = HAS_A ? i1[j] : 0.0f;      // (None, _) => 0.0
      const float b = HAS_B ? i2[j] : constant;  // (_, None) => constant
      const float r = WHICH == 0 ? fadd(a, b) : WHICH == 1 ? fsub(a, b) : WHICH == 2 ? fmul(a, b) : nonlinear(a, b);
      out[j] = (HAS_A || HAS_B) ? r : opaque(r);
    }
  }
};

// ---- GridSequencerModule::calc, src/synth/sequencer.rs:190-246 -----------------------------------
struct GridSeq {
  u32 step, n_steps;
  bool last_step, last_sync;
  float last_cv, inv_steps;
  const int* table;
  FZ_DEV void load(const Ctx& c, u32 sw, u32 table_off, u32 steps, float inv) {
    const u32 s0 = c.ld_state(sw);
    step = s0 & 0xFFFFu;
    last_step = (s0 >> 16) & 1u;
    last_sync = (s0 >> 17) & 1u;
    last_cv = asf(c.ld_state(sw + 1));
    inv_steps = inv;
    table = c.a->tables + table_off;
    n_steps = steps;
  }
  FZ_DEV void store(const Ctx& c, u32 sw) const {
    c.st_state(sw, step | (last_step ? 1u << 16 : 0u) | (last_sync ? 1u << 17 : 0u));
    c.st_state(sw + 1, asu(last_cv));
  }
  template <int U, bool FAST>
  FZ_DEV void run(const float* step_in, const float* sync_in, float* cv, float* gate, float* sync_out) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const float st = step_in ? step_in[j] : 0.0f;
      const float sy = sync_in ? sync_in[j] : 0.0f;
      const bool a_step = st > 0.0f, a_sync = sy > 0.0f;
      step += (a_step & !last_step) ? 1u : 0u;     // :220-223
      step = (a_sync & !last_sync) ? 0u : step;    // :224-226
      last_step = a_step;
      last_sync = a_sync;
      step = step >= n_steps ? 0u : step;          // :227-230
      const int cell = __ldg(table + step);
      const bool some = cell >= 0;
      const float c = some ? fmul((float)(cell & 0xFFFF), inv_steps) : last_cv;  // :231-239
      const float g = some ? ((cell >> 16) & 1 ? 1.0f : st) : 0.0f;
      last_cv = c;
      if (cv) cv[j] = c;
      if (gate) gate[j] = g;
      if (sync_out) sync_out[j] = step == 0u ? 1.0f : 0.0f;
    }
  }
};

// ---- PatternSequencerModule::calc, src/synth/sequencer.rs:482-533 (ports first .. first + 2) ------
struct PatSeq {
  u32 step, n_steps, first;
  bool last_step, last_sync;
  const int* table;
  FZ_DEV void load(const Ctx& c, u32 sw, u32 table_off, u32 steps, u32 first_port) {
    const u32 s0 = c.ld_state(sw);
    step = s0 & 0xFFFFu;
    last_step = (s0 >> 16) & 1u;
    last_sync = (s0 >> 17) & 1u;
    table = c.a->tables + table_off;
    n_steps = steps;
    first = first_port;
  }
  FZ_DEV void store(const Ctx& c, u32 sw) const {
    c.st_state(sw, step | (last_step ? 1u << 16 : 0u) | (last_sync ? 1u << 17 : 0u));
  }
  template <int U, bool FAST>
  FZ_DEV void run(const float* step_in, const float* sync_in, float* o0, float* o1, float* o2) {
    float* out[3] = {o0, o1, o2};
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const float st = step_in ? step_in[j] : 0.0f;
      const float sy = sync_in ? sync_in[j] : 0.0f;
      const bool a_step = st > 0.0f, a_sync = sy > 0.0f;
      step += (a_step & !last_step) ? 1u : 0u;
      step = (a_sync & !last_sync) ? 0u : step;
      last_step = a_step;
      last_sync = a_sync;
      step = step >= n_steps ? 0u : step;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (!out[k]) continue;
        const u32 portno = first + k;
        float v;
        if (portno == 8u) {
          v = step == 0u ? 1.0f : 0.0f;              // sync_out (:526)
        } else {
          const int cell = __ldg(table + portno * n_steps + step);
          v = cell < 0 ? 0.0f : (cell ? 1.0f : st);  // :515-524
        }
        out[k][j] = v;
      }
    }
  }
};

// ---- glibc 2.39 exp2f restated operation by operation (see dsp.cuh: the result steers an index) ---
static __device__ const unsigned long long kExp2fTab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};
FZ_DEV float exp2f_glibc(float x) {
  const double xd = (double)x;
  const double shift = 0x1.8p+52 / 32.0;
  double kd = dadd(xd, shift);
  const u64 ki = (u64)__double_as_longlong(kd);
  kd = dsub(kd, shift);
  const double r = dsub(xd, kd);
  const u64 t = __ldg(&kExp2fTab[ki & 31u]) + (ki << 47);
  const double s = __longlong_as_double((long long)t);
  const double z = dadd(dmul(0x1.c6af84b912394p-5, r), 0x1.ebfce50fac4f3p-3);
  const double r2 = dmul(r, r);
  double y = dadd(dmul(0x1.62e42ff0c52d6p-1, r), 1.0);
  y = dadd(dmul(z, r2), y);
  y = dmul(y, s);
  float out = __double2float_rn(y);
  out = x >= 128.0f ? __int_as_float(0x7f800000) : out;
  out = x <= -150.0f ? 0.0f : out;
  return x != x ? fadd(x, x) : out;
}

// ---- SampleModule::calc, src/synth/sample.rs:192-240 ---------------------------------------------
struct Sample {
  float pos, ratio;
  bool playing, last;
  const float* wave;
  u32 len;
  FZ_DEV void load(const Ctx& c, u32 sw, u32 desc_off) {
    pos = asf(c.ld_state(sw));
    const u32 s1 = c.ld_state(sw + 1);
    playing = s1 & 1u;
    last = (s1 >> 1) & 1u;
    const int* d = c.a->tables + desc_off;  // WaveDesc: offset, len, ratio, is_new
    wave = c.a->waves + (u32)d[0];
    len = (u32)d[1];
    ratio = __int_as_float(d[2]);
    if (d[3]) { pos = 0.0f; playing = false; }  // :212-216, first block after a load
  }
  FZ_DEV void store(const Ctx& c, u32 sw) const {
    c.st_state(sw, asu(pos));
    c.st_state(sw + 1, (playing ? 1u : 0u) | (last ? 2u : 0u));
  }
  template <int U, bool FAST>
  FZ_DEV void run(const float* gate, const float* cv, float* out) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const float g = gate ? gate[j] : 0.0f;
      const bool above = g > 0.0f;
      const bool trigger = above & !last;              // :219-221
      last = above;
      pos = trigger ? 0.0f : pos;                      // :222-225
      playing |= trigger;
      u32 idx = __float2uint_rz(pos);                  // `pos as usize`: saturating, NaN -> 0
      const bool past = idx >= len;                    // :226-229
      pos = past ? 0.0f : pos;
      playing &= !past;
      idx = past ? 0u : idx;
      const float x = len ? __ldg(wave + idx) : 0.0f;  // :230-234
      if (out) out[j] = x;
      const float e = cv ? exp2f_glibc(cv[j]) : 1.0f;
      pos = playing ? fadd(pos, fmul(ratio, e)) : pos; // :235-238
    }
  }
};

// ---- delayed (feedback) wires: rings f32 [R][B][V] in HBM, synth.rs:168-192 + :32 ----------------
struct Rings {
  u32 idx;  // (ring_phase + n) % B of the group being worked on
  FZ_DEV void init(const Ctx& c) { idx = c.a->ring_phase; }
  template <int U>
  FZ_DEV void load(const Ctx& c, u32 ring, float* out) const {
    const SrkFusedArgs& a = *c.a;
    const float* base = a.rings + (size_t)ring * a.B * a.V + c.v;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      u32 i = idx + j;
      i = i >= a.B ? i - a.B : i;
      out[j] = base[(size_t)i * a.V];
    }
  }
  template <int U>
  FZ_DEV void store(const Ctx& c, u32 ring, const float* in) const {
    const SrkFusedArgs& a = *c.a;
    float* base = a.rings + (size_t)ring * a.B * a.V + c.v;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      u32 i = idx + j;
      i = i >= a.B ? i - a.B : i;
      if (c.active) base[(size_t)i * a.V] = in[j];
    }
  }
  template <int U>
  FZ_DEV void advance(const Ctx& c) {
    idx += U;
    idx = idx >= c.a->B ? idx - c.a->B : idx;
  }
};

// ---- OutputModule::calc (output.rs:46-60) + the group's share of the mixdown ----------------------
// Every distinct wire feeding the Output module goes through a [32 samples][32 voices] f32 tile in shared
// memory (one STS per sample).  When a tile is full: (a) stems -- lane 0 issues one TMA bulk tensor store per
// channel, box {32 voices, 32 samples, 1 channel} of the [C][N][V] tensor (the box is clipped at N and V, so
// ragged tails need no special case); tiles are double-buffered and the buffer is reused only after
// cp.async.bulk.wait_group.read; (b) mix -- lane r sums sample row r over the 32 voices (see flush).
FZ_DEV u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
FZ_DEV void tma_store_3d(const SrkTensorMap* map, const float* tile, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :: "l"(map), "r"(smem_u32(tile)), "r"(x), "r"(y), "r"(z) : "memory");
}
FZ_DEV void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
FZ_DEV void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
FZ_DEV void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
FZ_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef SRK_TILE
#define SRK_TILE SRK_FUSED_TILE
#endif
constexpr int kTile = SRK_TILE;  // samples per output / cross-stage tile: 32, or 16 when shared memory is short
constexpr int kTileElems = kTile * 32;

// D distinct wires, C channels; chan_wire[c] = index of the distinct wire feeding channel c, or -1 (None: zeros).
template <int D, int C>
struct Out {
  float* tiles;  // this warp's [2][D][32][32]
  float* cur;    // this lane's column of the buffer being filled
  u32 buf;
  bool want_tile;

  FZ_DEV void init(const Ctx& c, float* group_smem) {
    tiles = group_smem;
    cur = tiles + c.lane;
    buf = 0;
    want_tile = (c.a->stems != nullptr) | (c.a->partial != nullptr);
  }
  // start of a tile: the buffer about to be written was read by the TMA store issued two tiles ago
  FZ_DEV void begin_tile(const Ctx& c) {
    if (c.a->use_tma && c.a->stems) {
      if (c.lane == 0) tma_wait_read<1>();
      __syncwarp();
    }
  }
  template <int U>
  FZ_DEV void put(const Ctx&, int d, u32 row, const float* w) {
    if (!want_tile) return;
    float* t = cur + d * kTileElems + row * 32;
#pragma unroll
    for (int j = 0; j < U; ++j) t[j * 32] = w[j];
  }
  FZ_DEV void flush(const Ctx& c, const SrkTensorMap* map, const int (&chan_wire)[C], u32 n0, u32 rows) {
    const SrkFusedArgs& a = *c.a;
    if (!want_tile) return;
    const float* t0 = tiles + (size_t)buf * D * kTileElems;
    if (a.stems) {
      if (a.use_tma) {
        fence_async_smem();  // this lane's generic-proxy tile writes -> visible to the async proxy
        __syncwarp();
        if (c.lane == 0) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch)
            if (chan_wire[ch] >= 0) tma_store_3d(map, t0 + chan_wire[ch] * kTileElems, (int)(c.group * 32), (int)n0, ch);
          tma_commit();
        }
      } else {
        __syncwarp();
        if (c.active) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            if (chan_wire[ch] < 0) continue;
            const float* t = t0 + chan_wire[ch] * kTileElems + c.lane;
            float* dst = a.stems + ((size_t)ch * a.n_samples + n0) * a.V + c.v;
            for (u32 r = 0; r < rows; ++r) __stcs(dst + (size_t)r * a.V, t[r * 32]);
          }
        }
      }
      if (c.active) {  // channels nobody feeds: bufs[c] stays zeros (output.rs:51-57)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          if (chan_wire[ch] >= 0) continue;
          float* dst = a.stems + ((size_t)ch * a.n_samples + n0) * a.V + c.v;
          for (u32 r = 0; r < rows; ++r) __stcs(dst + (size_t)r * a.V, 0.0f);
        }
      }
    }
    if (a.partial) {
      // lane r sums sample row r over the 32 voices, four voices per LDS.128, starting at the chunk (absolute sample
      // index) mod 8: the eight lanes of a quarter warp read eight different chunks (no bank conflict), and the order
      // of the additions is a function of the absolute sample index alone (chunked renders are bit-identical)
      __syncwarp();
      const u32 r = c.lane & (kTile - 1);  // (kTile = 16: the upper half warp repeats the lower one and stores nothing)
      const u32 s8 = (a.n_abs + n0 + r) & 7u;
      float sums[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float4* row4 = reinterpret_cast<const float4*>(t0 + d * kTileElems + r * 32);
        float acc = 0.0f;
        if (c.n_active == 32u) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 x = row4[(s8 + k) & 7u];
            acc = fadd(fadd(fadd(fadd(acc, x.x), x.y), x.z), x.w);
          }
        } else {
#pragma unroll 1
          for (int k = 0; k < 8; ++k) {
            const u32 ch4 = (s8 + k) & 7u;
            const float4 x = row4[ch4];
            const u32 v0 = ch4 * 4u;
            acc = fadd(acc, v0 < c.n_active ? x.x : 0.0f);
            acc = fadd(acc, v0 + 1u < c.n_active ? x.y : 0.0f);
            acc = fadd(acc, v0 + 2u < c.n_active ? x.z : 0.0f);
            acc = fadd(acc, v0 + 3u < c.n_active ? x.w : 0.0f);
          }
        }
        sums[d] = acc;
      }
      if (c.lane < rows) {  // (rows <= kTile)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          float s = 0.0f;
#pragma unroll
          for (int d = 0; d < D; ++d) s = chan_wire[ch] == d ? sums[d] : s;
          a.partial[((size_t)c.group * C + ch) * a.n_samples + n0 + r] = s;
        }
      }
      __syncwarp();
    }
    if (a.use_tma && a.stems) {
      buf ^= 1u;
      cur = tiles + (size_t)buf * D * kTileElems + c.lane;
    }
  }
  FZ_DEV void finish(const Ctx& c) {
    if (c.a->use_tma && c.a->stems && c.lane == 0) tma_wait_all();
  }
};

// ---- pipelined fused kernels: the S warps of a voice group run consecutive slices ("stages") of the patch, one tile
//      apart; a wire that crosses stages is a ring of XD [kTile][32] tiles in shared memory; done[s] counts the tiles
//      stage s has finished.  A stage waits until its producers have finished the tile it is about to read and until the
//      consumers of its own wires have finished the tile whose ring slot it is about to overwrite.
struct Pipe {
  volatile u32* done;
  float* xbase;  // this lane's column of cross tile 0
  u32 lane;
  FZ_DEV void init(const Ctx& c, float* group_smem, u32 out_floats, u32 cross_floats) {
    xbase = group_smem + out_floats + c.lane;
    done = reinterpret_cast<volatile u32*>(group_smem + out_floats + cross_floats);
    lane = c.lane;
  }
  // Flag protocol: the producer's tile stores, then (after __syncwarp, which orders the warp's lanes) lane 0's release
  // store of the count; the consumer's lane 0 spins with acquire loads, then __syncwarp.  CTA scope: every stage of a
  // group lives in the same thread block.  (No __threadfence_block(): it compiles to MEMBAR.SC.CTA, which waits for
  // the warp's outstanding global stores -- the stems -- on every tile.)
  FZ_DEV void wait_ge(u32 stage, int need) const {
    // EVERY lane polls (one broadcast LDS): a lane-0-only spin loop leaves the warp split into two divergent halves for
    // the rest of the tile, and every instruction of the tile then issues twice (profiles/r04f: 8.7 ms instead of 7.0)
    const u32 addr = smem_u32(const_cast<u32*>(done + stage));
    u32 v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    while ((int)v < need) {
      __nanosleep(40);  // (a polling warp shares its scheduler with working warps when an SM holds several groups)
      asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    }
    __syncwarp();
  }
  // a delayed (feedback) wire stored by a later stage: every sample up to (t + 1) * kTile - 1 - B must be in HBM
  FZ_DEV void wait_ring(u32 stage, u32 t, u32 B) const {
    const int last = (int)((t + 1u) * kTile) - 1 - (int)B;  // last delayed sample index this tile reads
    if (last >= 0) wait_ge(stage, last / kTile + 1);
  }
  FZ_DEV void publish(u32 stage, u32 tiles_done) const {
    __syncwarp();
    if (lane == 0) {
      const u32 addr = smem_u32(const_cast<u32*>(done + stage));
      asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(addr), "r"(tiles_done) : "memory");
    }
  }
  template <int U>
  FZ_DEV void put(u32 tile_index, u32 row, const float* w) const {  // tile_index = x * XD + slot
    float* t = xbase + tile_index * kTileElems + row * 32;
#pragma unroll
    for (int j = 0; j < U; ++j) t[j * 32] = w[j];
  }
  template <int U>
  FZ_DEV void get(u32 tile_index, u32 row, float* w) const {
    const float* t = xbase + tile_index * kTileElems + row * 32;
#pragma unroll
    for (int j = 0; j < U; ++j) w[j] = t[j * 32];
  }
};

// Sets up this lane's voice; false when the whole warp has no voice (it may then simply return: warps of a
// fused kernel never synchronise with each other).
FZ_DEV bool ctx_init(Ctx& c, const SrkFusedArgs* a, u32 n_stages) {
  c.a = a;
  c.lane = threadIdx.x & 31u;
  const u32 warp = threadIdx.x >> 5;
  c.stage = warp % n_stages;
  c.gib = warp / n_stages;
  c.group = blockIdx.x * ((blockDim.x >> 5) / n_stages) + c.gib;
  const u32 v0 = c.group * 32u;
  if (v0 >= a->V) return false;
  c.n_active = min(32u, a->V - v0);
  c.active = c.lane < c.n_active;
  c.v = c.active ? v0 + c.lane : a->V - 1u;
  return true;
}

}  // namespace fz
