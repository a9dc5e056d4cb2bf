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
// Every op is a small struct: load() brings the voice's state and parameters from the
// group's shared-memory tables into registers, run() advances one chunk of a wire tile,
// store() writes the state back.  A warp that owns a single instruction (pipelined
// schedule) calls load() once, run() per chunk and store() once -- the recurrence state
// lives in registers for the whole render; the one-warp schedule calls all three per
// chunk.
//
// Shape of run(): the chunk is walked in groups of kGroup samples (then single samples
// for a ragged tail).  Inside a group the code is straight-line: inputs are loaded
// first, the *stateless* per-sample work (coefficients, 2^x, sin, polyBLEP) is
// independent across the samples, and only the true recurrence (phase, ladder stages,
// envelope state machine) forms a dependent chain -- ptxas overlaps the two.  "Is this
// port connected" is decided once per op or per group, never per sample.
// Two measured facts shape the code (profiles/r01c_*, r01d_*): (1) with one or two warps
// per SM sub-partition a taken branch or a reconvergence point costs tens of cycles of
// instruction fetch, so the per-sample paths are select-based with bitwise (not
// short-circuit) predicates, and rare cases are tested once per group; (2) the
// instruction caches are small (L0 ~6 KB per sub-partition, 32 KB per SM) and the warps
// of a voice group run *different* op bodies concurrently, so every hot body has to stay
// a few KB -- hence groups of 4, and rare paths behind __noinline__ calls.
// Wire tiles are [K samples][32 voices] f32 in shared memory: sample k of this lane is
// p[k * 32], so offsets inside a group are immediates.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

#include "libm_glibc.cuh"
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
// every finite x, NaN for +-inf like fmod.  Kept out of line: the phase accumulator only
// ever sees x in [0, 2), where the result is x or x - 1 (both exact), and tests once per
// group whether anything else turned up (wrap01 / the `odd` flag in OscOp).
static __device__ __noinline__ double fmod1_exact(double x) { return dsub(x, trunc(x)); }
// fmod(x, 1.0) for x in [0, 2); NaN stays NaN.
__device__ __forceinline__ double wrap01(double x) { return x >= 1.0 ? dsub(x, 1.0) : x; }

// OscillatorModule::poly_blep, src/synth/oscillator.rs:50-67, select-based: both arms
// divide by dt, so ONE IEEE division of the selected numerator serves whichever arm is
// live.  A sample needs it only when `t < dt || t > 1 - dt` (the other case returns 0.0;
// dt == 0 can satisfy neither for t in [0, 1)), which callers test once per group.
__device__ __forceinline__ double blep_eval(double t, double dt, double one_minus_dt) {
  const bool lo = t < dt;
  const bool hi = !lo & (t > one_minus_dt);
  const double q = __ddiv_rn(lo ? t : dsub(t, 1.0), dt);
  const double qq = dmul(q, q);
  const double r_lo = dsub(dsub(dadd(q, q), qq), 1.0);
  const double r_hi = dadd(dadd(dadd(qq, q), q), 1.0);
  const double r = lo ? r_lo : (hi ? r_hi : 0.0);
  return dt == 0.0 ? 0.0 : r;
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
  uint32_t voice;      // global voice index (noise key)
  uint32_t seed_lo, seed_hi;
  const int32_t* tables;  // sequencer step tables / WaveDescs (shared memory, uniform over voices)
  const float* waves;     // Sample modules' tables (HBM, uniform over voices)
};

__device__ __forceinline__ float* wire(const Lane& ln, int slot) {
  if (slot < 0) return nullptr;
  const WireDesc d = ln.wd[slot];
  return ln.tiles + (size_t)(d.base + (ln.chunk & d.mask)) * ln.tile_elems;
}

// A wire slot resolved once (at load()): first tile of its ring + ring mask; at() picks
// the tile of the chunk being worked on.  nullptr base = not connected / nobody reads it.
struct Port {
  float* base;
  uint32_t mask;
  __device__ __forceinline__ float* at(const Lane& ln) const {
    return base ? base + (ln.chunk & mask) * ln.tile_elems : nullptr;
  }
};
__device__ __forceinline__ Port port(const Lane& ln, int slot) {
  if (slot < 0) return Port{nullptr, 0};
  const WireDesc d = ln.wd[slot];
  return Port{ln.tiles + d.base * ln.tile_elems, d.mask};
}

template <int U>
using UC = std::integral_constant<int, U>;

// Samples per straight-line group.  A translation unit may set SRK_SAMPLE_GROUP before including this header
// (each kernel image is its own translation unit; nothing here is linked across them).
#ifndef SRK_SAMPLE_GROUP
#define SRK_SAMPLE_GROUP 4
#endif
constexpr int kGroup = SRK_SAMPLE_GROUP;

// Runs body(UC<kGroup>, k) over full groups of samples in [k0, k1), body(UC<1>, k) over the tail.
template <class Body>
__device__ __forceinline__ void for_groups(int k0, int k1, Body&& body) {
#pragma unroll 1
  for (; k0 + kGroup <= k1; k0 += kGroup) body(UC<kGroup>(), k0);
#pragma unroll 1
  for (; k0 < k1; ++k0) body(UC<1>(), k0);
}
template <class Body>
__device__ __forceinline__ void for_groups(int kk, Body&& body) {
  for_groups(0, kk, body);
}

// ---- OscillatorModule::calc, src/synth/oscillator.rs:108-158 ----------------
struct OscOp {
  uint32_t* s;
  double pos, val, delta_const, sr;
  bool last, aa;
  Port p_cv, p_sync, p_sine, p_square, p_saw, p_dlo, p_dhi;
  bool ext;  // delta arrives on a pair of wires from an OscDeltaOp

  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    s = ln.st + ins.state * L;
    pos = __hiloint2double((int)s[L], (int)s[0]);
    last = s[2 * L] != 0u;
    const uint32_t* p = ln.pr + ins.param * L;
    val = (double)__uint_as_float(p[0]);
    delta_const = __hiloint2double((int)p[2 * L], (int)p[L]);
    sr = (double)ins.imm;
    aa = !(ins.flags & F_OSC_NO_ANTIALIASING);  // uniform over voices: rides in the instruction, not in a per-voice word
    ext = ins.n_ch != 0;
    p_cv = port(ln, ins.in[0]); p_sync = port(ln, ins.in[1]);
    p_dlo = port(ln, ext ? ins.in[2] : -1); p_dhi = port(ln, ext ? ins.in[3] : -1);
    p_sine = port(ln, ins.out[0]); p_square = port(ln, ins.out[1]); p_saw = port(ln, ins.out[2]);
  }
  __device__ __forceinline__ void store() {
    s[0] = (uint32_t)__double2loint(pos);
    s[L] = (uint32_t)__double2hiint(pos);
    s[2 * L] = last ? 1u : 0u;
  }

  // OUTS: bit 0 sine, bit 1 square, bit 2 saw are read by somebody.  Compile-time, because a
  // per-group `if (port connected)` costs more than the arithmetic it guards (see header).
  // DM: where delta comes from -- 0 the voice's constant, 1 the CV input (converted here),
  // 2 an OscDeltaOp's wire pair.
  template <int DM, bool HAS_SYNC, int OUTS>
  __device__ __forceinline__ void run_t(const Lane& ln, int kb, int ke) {
    constexpr bool SINE = OUTS & 1, SQUARE = OUTS & 2, SAW = OUTS & 4;
    constexpr bool HAS_CV = DM == 1;
    const float* cv = p_cv.at(ln);
    const float* dlo = p_dlo.at(ln);
    const float* dhi = p_dhi.at(ln);
    const float* sync = p_sync.at(ln);
    constexpr bool has_sync = HAS_SYNC;
    float* sine = p_sine.at(ln);
    float* square = p_square.at(ln);
    float* saw = p_saw.at(ln);
    // Work on copies: when the op object itself ends up in local memory (it is captured by
    // reference in the resident loop), the recurrence must not go through it every sample
    // (r01x: 27 % long-scoreboard stalls on the phase add).
    double pos = this->pos;
    bool last = this->last;
    const double val = this->val, delta_const = this->delta_const, sr = this->sr;
    const bool aa = this->aa;
    for_groups(kb, ke, [&](auto u, int k0) {
      constexpr int U = decltype(u)::value;
      float cvv[U];
      bool edge[U];  // sync transition on this sample (:125-131)
      double ps[U], dl[U];
      if (HAS_CV) {
#pragma unroll
        for (int j = 0; j < U; ++j) cvv[j] = cv[(k0 + j) * L];
      }
#pragma unroll
      for (int j = 0; j < U; ++j) edge[j] = false;
      if (has_sync) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const bool above = sync[(k0 + j) * L] > 0.0f;
          edge[j] = above & !last;
          last = above;
        }
      }
      // get_freq_in_hz (:43-48) then / sample_rate (:132): stateless
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (DM == 2)
          dl[j] = __hiloint2double(__float_as_int(dhi[(k0 + j) * L]), __float_as_int(dlo[(k0 + j) * L]));
        else
          dl[j] = HAS_CV ? __ddiv_rn(dmul(440.0, exp2_glibc(dadd((double)cvv[j], val))), sr) : delta_const;
      }
      // the recurrence: sync reset, pos += delta; pos %= 1.0 (:151-152)
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
      if (odd) {
        pos = pos0;
#pragma unroll
        for (int j = 0; j < U; ++j) {
          if (HAS_SYNC) pos = edge[j] ? 0.0 : pos;
          ps[j] = pos;
          pos = fmod1_exact(dadd(pos, dl[j]));
        }
      }
      if (SINE) sine_port(ps, sine + k0 * L, L);  // (float) of glibc's f64 sin, libm_glibc.cuh
      if (SQUARE || SAW) {
        // polyBLEP corrections (:135-149) are zero unless a sample sits within dt of a
        // discontinuity: one test per group picks the plain or the corrected write-out.
        // (x - 0.0f == x bit for bit, so the plain path skips the subtraction.)
        double om[U], p2[U];
        bool near = false;
#pragma unroll
        for (int j = 0; j < U; ++j) {
          om[j] = dsub(1.0, dl[j]);
          near |= (ps[j] < dl[j]) | (ps[j] > om[j]);
          if (SQUARE) {
            // (pos + 0.5) % 1.0: pos is in [0, 1) or NaN (it is itself the result of `% 1.0`
            // of a non-negative sum), so the sum is in [0.5, 1.5)
            p2[j] = wrap01(dadd(ps[j], 0.5));
            near |= (p2[j] < dl[j]) | (p2[j] > om[j]);
          }
        }
        if (aa & near) {
          double pb0[U];
#pragma unroll
          for (int j = 0; j < U; ++j) pb0[j] = blep_eval(ps[j], dl[j], om[j]);
          if (SQUARE) {
#pragma unroll
            for (int j = 0; j < U; ++j) {
              const float base = ps[j] < 0.5 ? -1.0f : 1.0f;
              square[(k0 + j) * L] = fsub(base, __double2float_rn(dsub(pb0[j], blep_eval(p2[j], dl[j], om[j]))));
            }
          }
          if (SAW) {
#pragma unroll
            for (int j = 0; j < U; ++j)
              saw[(k0 + j) * L] = fsub(fsub(fmul(__double2float_rn(ps[j]), 2.0f), 1.0f), __double2float_rn(pb0[j]));
          }
        } else {
          if (SQUARE) {
#pragma unroll
            for (int j = 0; j < U; ++j) square[(k0 + j) * L] = ps[j] < 0.5 ? -1.0f : 1.0f;
          }
          if (SAW) {
#pragma unroll
            for (int j = 0; j < U; ++j) saw[(k0 + j) * L] = fsub(fmul(__double2float_rn(ps[j]), 2.0f), 1.0f);
          }
        }
      }
    });
    // with no sync input the detector sees 0.0 every sample: `last` just goes false
    if (!has_sync && ke > kb) last = false;
    this->pos = pos;
    this->last = last;
  }

  template <int DM, bool HAS_SYNC>
  __device__ __forceinline__ void run_outs(const Lane& ln, int kb, int ke) {
    const int outs = (p_sine.base ? 1 : 0) | (p_square.base ? 2 : 0) | (p_saw.base ? 4 : 0);
    switch (outs) {
      case 0: run_t<DM, HAS_SYNC, 0>(ln, kb, ke); break;
      case 1: run_t<DM, HAS_SYNC, 1>(ln, kb, ke); break;
      case 2: run_t<DM, HAS_SYNC, 2>(ln, kb, ke); break;
      case 3: run_t<DM, HAS_SYNC, 3>(ln, kb, ke); break;
      case 4: run_t<DM, HAS_SYNC, 4>(ln, kb, ke); break;
      case 5: run_t<DM, HAS_SYNC, 5>(ln, kb, ke); break;
      case 6: run_t<DM, HAS_SYNC, 6>(ln, kb, ke); break;
      default: run_t<DM, HAS_SYNC, 7>(ln, kb, ke); break;
    }
  }

  // The phase recurrence alone over [kb, ke) for an oscillator with neither CV nor sync
  // (what a time-split copy runs outside its own share): DADD, DADD, compare, select per
  // sample, 8 samples per loop trip, the odd-step test once per trip.
  __device__ __forceinline__ void advance(int kb, int ke) {
    const double dl = delta_const;
    double pos = this->pos;  // (a register copy, see run_t)
    int k = kb;
#pragma unroll 1
    for (; k + 8 <= ke; k += 8) {
      const double pos0 = pos;
      bool odd = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double x = dadd(pos, dl);
        odd |= !(x < 2.0);
        pos = wrap01(x);
      }
      if (odd) {
        pos = pos0;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) pos = fmod1_exact(dadd(pos, dl));
      }
    }
#pragma unroll 1
    for (; k < ke; ++k) {
      const double x = dadd(pos, dl);
      pos = x < 2.0 ? wrap01(x) : fmod1_exact(x);
    }
    this->pos = pos;
    if (ke > kb) last = false;
  }

  // flags = (n << 4) | i (| F_OSC_NO_ANTIALIASING) for a time-split copy (program.cpp): every copy advances the phase
  // through the whole chunk, copy i shapes the outputs of the i-th n-th of every chunk (so
  // within one barrier interval each copy does phase(K) + shape(K / n)).
  __device__ __forceinline__ void run(const Instr& ins, const Lane& ln, int kk) {
    const uint32_t n = ins.flags >> 4;
    const bool cv = p_cv.base != nullptr, sync = p_sync.base != nullptr;
    int lo = 0, hi = kk;  // the span this copy shapes
    if (n > 1) {
      const int span = (int)(ln.tile_elems / L / n);
      lo = min(kk, (int)(ins.flags & 7u) * span);
      hi = min(kk, lo + span);
    }
    if (ext) {  // delta from an OscDeltaOp: the phase runs over the whole chunk, shaping over [lo, hi)
      if (sync) {
        run_t<2, true, 0>(ln, 0, lo); run_outs<2, true>(ln, lo, hi); run_t<2, true, 0>(ln, hi, kk);
      } else {
        run_t<2, false, 0>(ln, 0, lo); run_outs<2, false>(ln, lo, hi); run_t<2, false, 0>(ln, hi, kk);
      }
    } else if (cv) {  // (never split: program.cpp takes the conversion out first when warps allow)
      if (sync) run_outs<1, true>(ln, 0, kk);
      else run_outs<1, false>(ln, 0, kk);
    } else if (sync) {
      run_t<0, true, 0>(ln, 0, lo); run_outs<0, true>(ln, lo, hi); run_t<0, true, 0>(ln, hi, kk);
    } else {
      advance(0, lo); run_outs<0, false>(ln, lo, hi); advance(hi, kk);
    }
  }
  __device__ __forceinline__ bool owns_state(const Instr& ins) const { return (ins.flags & 7u) == 0; }
};

// The V/oct conversion of a CV-driven oscillator on its own warp(s) (program.cpp splits it off
// when warps are spare): delta = 440 * 2^(cv + val) / sample_rate, get_freq_in_hz (:43-48) then
// `/ sample_rate` (:132), the same f64 operations as OscOp's DM == 1 path; the result travels as
// the two 32-bit halves of the f64 on a pair of wires.  Stateless, so a time-split copy
// (flags = (n << 4) | i) simply converts its n-th of every chunk.
struct OscDeltaOp {
  double val, sr;
  Port p_cv, p_lo, p_hi;
  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    val = (double)__uint_as_float(ln.pr[ins.param * L]);
    sr = (double)ins.imm;
    p_cv = port(ln, ins.in[0]);
    p_lo = port(ln, ins.out[0]); p_hi = port(ln, ins.out[1]);
  }
  __device__ __forceinline__ void store() {}
  __device__ __forceinline__ void run(const Instr& ins, const Lane& ln, int kk) {
    const float* cv = p_cv.at(ln);
    float* lo = p_lo.at(ln);
    float* hi = p_hi.at(ln);
    const uint32_t n = ins.flags >> 4;
    int kb = 0, ke = kk;
    if (n > 1) {
      const int span = (int)(ln.tile_elems / L / n);
      kb = min(kk, (int)(ins.flags & 7u) * span);
      ke = min(kk, kb + span);
    }
    for_groups(kb, ke, [&](auto u, int k0) {
      constexpr int U = decltype(u)::value;
      float c[U];
      double d[U];
#pragma unroll
      for (int j = 0; j < U; ++j) c[j] = cv[(k0 + j) * L];
#pragma unroll
      for (int j = 0; j < U; ++j) d[j] = __ddiv_rn(dmul(440.0, exp2_glibc(dadd((double)c[j], val))), sr);
#pragma unroll
      for (int j = 0; j < U; ++j) {
        lo[(k0 + j) * L] = __int_as_float(__double2loint(d[j]));
        hi[(k0 + j) * L] = __int_as_float(__double2hiint(d[j]));
      }
    });
  }
};

// ---- The oscillator split in two (program.cpp, SRK_OSC_PHASE_SPLIT=1, pipelined schedule only) ----
// OscPhaseOp runs what is sequential -- sync reset, `pos += delta; pos %= 1.0` (:125-131, :151-152) --
// once per oscillator and publishes the phase every sample is shaped at as the two halves of the f64
// on a wire pair; OscShapeOp (time-split like OscDeltaOp) turns phases into sine / square / saw with
// exactly OscOp's arithmetic.  Without this every time-split OscOp copy repeats the whole recurrence.
struct OscPhaseOp {
  uint32_t* s;
  double pos, delta_const;
  bool last, ext;
  Port p_sync, p_dlo, p_dhi, p_lo, p_hi;
  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    s = ln.st + ins.state * L;
    pos = __hiloint2double((int)s[L], (int)s[0]);
    last = s[2 * L] != 0u;
    const uint32_t* p = ln.pr + ins.param * L;
    delta_const = __hiloint2double((int)p[2 * L], (int)p[L]);
    ext = ins.n_ch != 0;
    p_sync = port(ln, ins.in[1]);
    p_dlo = port(ln, ext ? ins.in[2] : -1); p_dhi = port(ln, ext ? ins.in[3] : -1);
    p_lo = port(ln, ins.out[0]); p_hi = port(ln, ins.out[1]);
  }
  __device__ __forceinline__ void store() {
    s[0] = (uint32_t)__double2loint(pos);
    s[L] = (uint32_t)__double2hiint(pos);
    s[2 * L] = last ? 1u : 0u;
  }
  template <bool EXT, bool HAS_SYNC>
  __device__ __forceinline__ void run_t(const Lane& ln, int kk) {
    const float* dlo = p_dlo.at(ln);
    const float* dhi = p_dhi.at(ln);
    const float* sync = p_sync.at(ln);
    float* lo = p_lo.at(ln);
    float* hi = p_hi.at(ln);
    double pos = this->pos;
    bool last = this->last;
    const double delta_const = this->delta_const;
    for_groups(kk, [&](auto u, int k0) {
      constexpr int U = decltype(u)::value;
      bool edge[U];
      double ps[U], dl[U];
#pragma unroll
      for (int j = 0; j < U; ++j) edge[j] = false;
      if (HAS_SYNC) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const bool above = sync[(k0 + j) * L] > 0.0f;
          edge[j] = above & !last;
          last = above;
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j)
        dl[j] = EXT ? __hiloint2double(__float_as_int(dhi[(k0 + j) * L]), __float_as_int(dlo[(k0 + j) * L])) : delta_const;
      const double pos0 = pos;
      bool odd = false;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (HAS_SYNC) pos = edge[j] ? 0.0 : pos;
        ps[j] = pos;
        const double x = dadd(pos, dl[j]);
        odd |= !(x < 2.0);
        pos = wrap01(x);
      }
      if (odd) {
        pos = pos0;
#pragma unroll
        for (int j = 0; j < U; ++j) {
          if (HAS_SYNC) pos = edge[j] ? 0.0 : pos;
          ps[j] = pos;
          pos = fmod1_exact(dadd(pos, dl[j]));
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        lo[(k0 + j) * L] = __int_as_float(__double2loint(ps[j]));
        hi[(k0 + j) * L] = __int_as_float(__double2hiint(ps[j]));
      }
    });
    if (!HAS_SYNC && kk > 0) last = false;
    this->pos = pos;
    this->last = last;
  }
  __device__ __forceinline__ void run(const Instr&, const Lane& ln, int kk) {
    const bool sync = p_sync.base != nullptr;
    if (ext) { if (sync) run_t<true, true>(ln, kk); else run_t<true, false>(ln, kk); }
    else { if (sync) run_t<false, true>(ln, kk); else run_t<false, false>(ln, kk); }
  }
};

struct OscShapeOp {
  double delta_const;
  bool aa, ext;
  Port p_plo, p_phi, p_dlo, p_dhi, p_sine, p_square, p_saw;
  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    const uint32_t* p = ln.pr + ins.param * L;
    delta_const = __hiloint2double((int)p[2 * L], (int)p[L]);
    aa = !(ins.flags & F_OSC_NO_ANTIALIASING);
    ext = ins.n_ch != 0;
    p_plo = port(ln, ins.in[0]); p_phi = port(ln, ins.in[1]);
    p_dlo = port(ln, ext ? ins.in[2] : -1); p_dhi = port(ln, ext ? ins.in[3] : -1);
    p_sine = port(ln, ins.out[0]); p_square = port(ln, ins.out[1]); p_saw = port(ln, ins.out[2]);
  }
  __device__ __forceinline__ void store() {}
  template <bool EXT, int OUTS>
  __device__ __forceinline__ void run_t(const Lane& ln, int kb, int ke) {
    constexpr bool SINE = OUTS & 1, SQUARE = OUTS & 2, SAW = OUTS & 4;
    const float* plo = p_plo.at(ln);
    const float* phi = p_phi.at(ln);
    const float* dlo = p_dlo.at(ln);
    const float* dhi = p_dhi.at(ln);
    float* sine = p_sine.at(ln);
    float* square = p_square.at(ln);
    float* saw = p_saw.at(ln);
    const double delta_const = this->delta_const;
    const bool aa = this->aa;
    for_groups(kb, ke, [&](auto u, int k0) {
      constexpr int U = decltype(u)::value;
      double ps[U], dl[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        ps[j] = __hiloint2double(__float_as_int(phi[(k0 + j) * L]), __float_as_int(plo[(k0 + j) * L]));
        dl[j] = EXT ? __hiloint2double(__float_as_int(dhi[(k0 + j) * L]), __float_as_int(dlo[(k0 + j) * L])) : delta_const;
      }
      if (SINE) sine_port(ps, sine + k0 * L, L);  // (:133)
      if (SQUARE || SAW) {  // (:135-149), as OscOp::run_t
        double om[U], p2[U];
        bool near = false;
#pragma unroll
        for (int j = 0; j < U; ++j) {
          om[j] = dsub(1.0, dl[j]);
          near |= (ps[j] < dl[j]) | (ps[j] > om[j]);
          if (SQUARE) {
            p2[j] = wrap01(dadd(ps[j], 0.5));
            near |= (p2[j] < dl[j]) | (p2[j] > om[j]);
          }
        }
        if (aa & near) {
          double pb0[U];
#pragma unroll
          for (int j = 0; j < U; ++j) pb0[j] = blep_eval(ps[j], dl[j], om[j]);
          if (SQUARE) {
#pragma unroll
            for (int j = 0; j < U; ++j) {
              const float base = ps[j] < 0.5 ? -1.0f : 1.0f;
              square[(k0 + j) * L] = fsub(base, __double2float_rn(dsub(pb0[j], blep_eval(p2[j], dl[j], om[j]))));
            }
          }
          if (SAW) {
#pragma unroll
            for (int j = 0; j < U; ++j)
              saw[(k0 + j) * L] = fsub(fsub(fmul(__double2float_rn(ps[j]), 2.0f), 1.0f), __double2float_rn(pb0[j]));
          }
        } else {
          if (SQUARE) {
#pragma unroll
            for (int j = 0; j < U; ++j) square[(k0 + j) * L] = ps[j] < 0.5 ? -1.0f : 1.0f;
          }
          if (SAW) {
#pragma unroll
            for (int j = 0; j < U; ++j) saw[(k0 + j) * L] = fsub(fmul(__double2float_rn(ps[j]), 2.0f), 1.0f);
          }
        }
      }
    });
  }
  template <bool EXT>
  __device__ __forceinline__ void run_outs(const Lane& ln, int kb, int ke) {
    const int outs = (p_sine.base ? 1 : 0) | (p_square.base ? 2 : 0) | (p_saw.base ? 4 : 0);
    switch (outs) {
      case 1: run_t<EXT, 1>(ln, kb, ke); break;
      case 2: run_t<EXT, 2>(ln, kb, ke); break;
      case 3: run_t<EXT, 3>(ln, kb, ke); break;
      case 4: run_t<EXT, 4>(ln, kb, ke); break;
      case 5: run_t<EXT, 5>(ln, kb, ke); break;
      case 6: run_t<EXT, 6>(ln, kb, ke); break;
      case 7: run_t<EXT, 7>(ln, kb, ke); break;
      default: break;
    }
  }
  __device__ __forceinline__ void run(const Instr& ins, const Lane& ln, int kk) {
    const uint32_t n = ins.flags >> 4;
    int kb = 0, ke = kk;
    if (n > 1) {
      const int span = (int)(ln.tile_elems / L / n);
      kb = min(kk, (int)(ins.flags & 7u) * span);
      ke = min(kk, kb + span);
    }
    if (ext) run_outs<true>(ln, kb, ke); else run_outs<false>(ln, kb, ke);
  }
};

// ---- NoiseModule::calc, src/synth/oscillator.rs:381-388 (seeded generator) ----
struct NoiseOp {
  uint32_t* s;
  uint64_t n;
  Port p_out;

  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    s = ln.st + ins.state * L;
    n = ((uint64_t)s[L] << 32) | s[0];
    p_out = port(ln, ins.out[0]);
  }
  __device__ __forceinline__ void store() {
    s[0] = (uint32_t)n;
    s[L] = (uint32_t)(n >> 32);
  }
  __device__ __forceinline__ void run(const Instr& ins, const Lane& ln, int kk) {
    float* out = p_out.at(ln);
    if (out) {
      auto draw = [&](uint64_t blk, uint32_t (&c)[4]) {
        c[0] = (uint32_t)blk; c[1] = (uint32_t)(blk >> 32); c[2] = ln.voice; c[3] = ins.aux;
        philox4x32_10(c, ln.seed_lo, ln.seed_hi);
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
#pragma unroll 1
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
  }
};

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
struct MoogOp {
  uint32_t* s;
  float f, p, q, b0, b1, b2, b3, b4, c_freq, c_res;
  float freq, r, exp_amt;
  bool ext;  // coefficients arrive on three wires from a MoogCoefOp, which then owns f, p, q and the cache
  Port p_audio, p_cv, p_lp, p_bp, p_hp, p_f, p_p, p_q;

  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    s = ln.st + ins.state * L;
    f = __uint_as_float(s[0]); p = __uint_as_float(s[L]); q = __uint_as_float(s[2 * L]);
    b0 = __uint_as_float(s[3 * L]); b1 = __uint_as_float(s[4 * L]); b2 = __uint_as_float(s[5 * L]);
    b3 = __uint_as_float(s[6 * L]); b4 = __uint_as_float(s[7 * L]);
    c_freq = __uint_as_float(s[8 * L]); c_res = __uint_as_float(s[9 * L]);
    const uint32_t* pp = ln.pr + ins.param * L;
    freq = __uint_as_float(pp[0]);
    r = fminf(fmaxf(__uint_as_float(pp[L]), 0.0f), 1.0f);  // :214
    exp_amt = __uint_as_float(pp[2 * L]);
    ext = ins.flags & F_MOOG_EXT_COEF;
    p_audio = port(ln, ins.in[0]);
    p_cv = ext ? Port{nullptr, 0} : port(ln, ins.in[1]);
    p_f = port(ln, ext ? ins.in[1] : -1); p_p = port(ln, ext ? ins.in[2] : -1); p_q = port(ln, ext ? ins.in[3] : -1);
    p_lp = port(ln, ins.out[0]); p_bp = port(ln, ins.out[1]); p_hp = port(ln, ins.out[2]);
  }
  __device__ __forceinline__ void store() {
    s[3 * L] = __float_as_uint(b0); s[4 * L] = __float_as_uint(b1); s[5 * L] = __float_as_uint(b2);
    s[6 * L] = __float_as_uint(b3); s[7 * L] = __float_as_uint(b4);
    if (ext) return;
    s[0] = __float_as_uint(f); s[L] = __float_as_uint(p); s[2 * L] = __float_as_uint(q);
    s[8 * L] = __float_as_uint(c_freq); s[9 * L] = __float_as_uint(c_res);
  }

  struct Taps {  // this chunk's tiles
    const float *audio, *cv, *wf, *wp, *wq;
    float *lowpass, *bandpass, *highpass;
  };
  __device__ __forceinline__ Taps taps(const Lane& ln) const {
    return Taps{p_audio.at(ln), p_cv.at(ln), p_f.at(ln), p_p.at(ln), p_q.at(ln), p_lp.at(ln), p_bp.at(ln), p_hp.at(ln)};
  }

  // U samples starting at k0.  COEF: 0 = no CV (constant cutoff), 1 = CV, coefficients computed
  // here, 2 = coefficients from a MoogCoefOp.  OUTS: bit 0 lowpass, bit 1 bandpass, bit 2 highpass
  // are read by somebody (compile-time, like OscOp's).  First everything that does not depend on
  // the ladder state (input loads, the coefficient block, in sample order: the cache key and the
  // `virgin` flag are sequential but cheap), then the dependent chain (:69-82) and the write-out.
  template <int COEF, int OUTS, int U>
  __device__ __forceinline__ void group(const Taps& t, int k0, bool& virgin) {
    float a[U], fj[U], pj[U], qj[U], in_[U], o3[U], o4[U];
#pragma unroll
    for (int j = 0; j < U; ++j) a[j] = 0.0f;
    if (t.audio) {
#pragma unroll
      for (int j = 0; j < U; ++j) a[j] = t.audio[(k0 + j) * L];
    }
    if (COEF == 1) {
      float fc[U];
#pragma unroll
      for (int j = 0; j < U; ++j) fc[j] = fminf(fmaxf(fadd(freq, fmul(t.cv[(k0 + j) * L], exp_amt)), 0.0f), 0.9f);  // :213
#pragma unroll
      for (int j = 0; j < U; ++j) moog_coef(fc[j], r, fj[j], pj[j], qj[j]);
      if (virgin) {  // only until (fc, r) first leaves (0, 0)
#pragma unroll
        for (int j = 0; j < U; ++j) {
          virgin = virgin & (fc[j] == 0.0f) & (r == 0.0f);
          if (virgin) { fj[j] = 0.0f; pj[j] = 0.0f; qj[j] = 0.0f; }
        }
      }
      if (!virgin) { c_freq = fc[U - 1]; c_res = r; }
      f = fj[U - 1]; p = pj[U - 1]; q = qj[U - 1];
    } else if (COEF == 2) {
#pragma unroll
      for (int j = 0; j < U; ++j) { fj[j] = t.wf[(k0 + j) * L]; pj[j] = t.wp[(k0 + j) * L]; qj[j] = t.wq[(k0 + j) * L]; }
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j) { fj[j] = f; pj[j] = p; qj[j] = q; }
    }
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
    if (OUTS & 1) {
#pragma unroll
      for (int j = 0; j < U; ++j) t.lowpass[(k0 + j) * L] = o4[j];
    }
    if (OUTS & 4) {
#pragma unroll
      for (int j = 0; j < U; ++j) t.highpass[(k0 + j) * L] = fsub(in_[j], o4[j]);
    }
    if (OUTS & 2) {
#pragma unroll
      for (int j = 0; j < U; ++j) t.bandpass[(k0 + j) * L] = fmul(3.0f, fsub(o3[j], o4[j]));
    }
  }

  template <int COEF>
  __device__ __forceinline__ bool begin(int kk) {  // -> the `virgin` flag for this stretch
    if (COEF == 0 && kk > 0) {  // cutoff is constant: one cache check (:61)
      const float fc = fminf(fmaxf(fadd(freq, fmul(0.0f, exp_amt)), 0.0f), 0.9f);  // :213 with cv = 0.0
      if (fc != c_freq || r != c_res) {
        c_freq = fc;
        c_res = r;
        moog_coef(fc, r, f, p, q);
      }
    }
    return (c_freq == 0.0f) & (c_res == 0.0f) & (f == 0.0f);
  }

  template <int COEF, int OUTS>
  __device__ __forceinline__ void run_t(const Lane& ln, int kk) {
    bool virgin = begin<COEF>(kk);
    const Taps t = taps(ln);
    int k0 = 0;
#pragma unroll 1
    for (; k0 + kGroup <= kk; k0 += kGroup) group<COEF, OUTS, kGroup>(t, k0, virgin);
#pragma unroll 1
    for (; k0 < kk; ++k0) group<COEF, OUTS, 1>(t, k0, virgin);
  }

  // The resident form for a whole render: ONE loop over every group of every full chunk, the
  // block barrier and the four tile pointers inline at the chunk boundaries.  The filter is the
  // pipeline's critical stage, and the chunk-at-a-time form spent ~1.5k cycles per barrier
  // interval refetching its cold prologue / epilogue code (profiles/r01i_k8, r01t).
  template <int COEF, int OUTS, class Sync>
  __device__ __forceinline__ void run_all_t(Lane& ln, uint32_t n_samples, Sync&& sync) {
    const uint32_t K = ln.tile_elems / L;
    const uint32_t full_chunks = n_samples / K, groups_per_chunk = K / kGroup;
    bool virgin = begin<COEF>((int)n_samples);
    Taps t = taps(ln);
    uint32_t in_chunk = 0;
    int k0 = 0;
    ln.chunk = 0;
    t = taps(ln);
#pragma unroll 1
    for (uint32_t g = 0; g < full_chunks * groups_per_chunk; ++g) {
      if (in_chunk == groups_per_chunk) {
        sync();
        in_chunk = 0;
        k0 = 0;
        ++ln.chunk;
        t = taps(ln);
      }
      group<COEF, OUTS, kGroup>(t, k0, virgin);
      k0 += kGroup;
      ++in_chunk;
    }
    if (full_chunks) sync();
    const int rest = (int)(n_samples - full_chunks * K);
    if (rest) {  // ragged last chunk
      ln.chunk = full_chunks;
      t = taps(ln);
      int k = 0;
#pragma unroll 1
      for (; k + kGroup <= rest; k += kGroup) group<COEF, OUTS, kGroup>(t, k, virgin);
#pragma unroll 1
      for (; k < rest; ++k) group<COEF, OUTS, 1>(t, k, virgin);
      sync();
    }
  }

  template <int COEF, class Sync>
  __device__ __forceinline__ void run_all_outs(Lane& ln, uint32_t n_samples, Sync&& sync) {
    const int outs = (p_lp.base ? 1 : 0) | (p_bp.base ? 2 : 0) | (p_hp.base ? 4 : 0);
    switch (outs) {
      case 0: run_all_t<COEF, 0>(ln, n_samples, sync); break;
      case 1: run_all_t<COEF, 1>(ln, n_samples, sync); break;
      case 2: run_all_t<COEF, 2>(ln, n_samples, sync); break;
      case 3: run_all_t<COEF, 3>(ln, n_samples, sync); break;
      case 4: run_all_t<COEF, 4>(ln, n_samples, sync); break;
      case 5: run_all_t<COEF, 5>(ln, n_samples, sync); break;
      case 6: run_all_t<COEF, 6>(ln, n_samples, sync); break;
      default: run_all_t<COEF, 7>(ln, n_samples, sync); break;
    }
  }
  // One sync() after every chunk, like the chunk-at-a-time loop.  K must be a multiple of kGroup.
  template <class Sync>
  __device__ __forceinline__ void run_all(Lane& ln, uint32_t n_samples, Sync&& sync) {
    if (ext) run_all_outs<2>(ln, n_samples, sync);
    else if (p_cv.base) run_all_outs<1>(ln, n_samples, sync);
    else run_all_outs<0>(ln, n_samples, sync);
  }

  template <int COEF>
  __device__ __forceinline__ void run_outs(const Lane& ln, int kk) {
    const int outs = (p_lp.base ? 1 : 0) | (p_bp.base ? 2 : 0) | (p_hp.base ? 4 : 0);
    switch (outs) {
      case 0: run_t<COEF, 0>(ln, kk); break;
      case 1: run_t<COEF, 1>(ln, kk); break;
      case 2: run_t<COEF, 2>(ln, kk); break;
      case 3: run_t<COEF, 3>(ln, kk); break;
      case 4: run_t<COEF, 4>(ln, kk); break;
      case 5: run_t<COEF, 5>(ln, kk); break;
      case 6: run_t<COEF, 6>(ln, kk); break;
      default: run_t<COEF, 7>(ln, kk); break;
    }
  }

  __device__ __forceinline__ void run(const Instr&, const Lane& ln, int kk) {
    if (ext) run_outs<2>(ln, kk);
    else if (p_cv.base) run_outs<1>(ln, kk);
    else run_outs<0>(ln, kk);
  }
};

// The coefficient half of a CV-driven filter on its own warp (program.cpp splits it off when
// warps are spare): same arithmetic and cache semantics as MoogOp's COEF == 1 path; owns the
// state words f, p, q, freq, res and writes (f, p, q) per sample to three wires.
struct MoogCoefOp {
  uint32_t* s;
  float f, p, q, c_freq, c_res, freq, r, exp_amt;
  Port p_cv, p_f, p_p, p_q;

  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    s = ln.st + ins.state * L;
    f = __uint_as_float(s[0]); p = __uint_as_float(s[L]); q = __uint_as_float(s[2 * L]);
    c_freq = __uint_as_float(s[8 * L]); c_res = __uint_as_float(s[9 * L]);
    const uint32_t* pp = ln.pr + ins.param * L;
    freq = __uint_as_float(pp[0]);
    r = fminf(fmaxf(__uint_as_float(pp[L]), 0.0f), 1.0f);  // :214
    exp_amt = __uint_as_float(pp[2 * L]);
    p_cv = port(ln, ins.in[0]);
    p_f = port(ln, ins.out[0]); p_p = port(ln, ins.out[1]); p_q = port(ln, ins.out[2]);
  }
  __device__ __forceinline__ void store() {
    s[0] = __float_as_uint(f); s[L] = __float_as_uint(p); s[2 * L] = __float_as_uint(q);
    s[8 * L] = __float_as_uint(c_freq); s[9 * L] = __float_as_uint(c_res);
  }
  __device__ __forceinline__ void run(const Instr&, const Lane& ln, int kk) {
    const float* cv = p_cv.at(ln);
    float* wf = p_f.at(ln);
    float* wp = p_p.at(ln);
    float* wq = p_q.at(ln);
    bool virgin = (c_freq == 0.0f) & (c_res == 0.0f) & (f == 0.0f);
    for_groups(kk, [&](auto u, int k0) {
      constexpr int U = decltype(u)::value;
      float fc[U], fj[U], pj[U], qj[U];
#pragma unroll
      for (int j = 0; j < U; ++j) fc[j] = fminf(fmaxf(fadd(freq, fmul(cv[(k0 + j) * L], exp_amt)), 0.0f), 0.9f);  // :213
#pragma unroll
      for (int j = 0; j < U; ++j) moog_coef(fc[j], r, fj[j], pj[j], qj[j]);
      if (virgin) {  // only until (fc, r) first leaves (0, 0)
#pragma unroll
        for (int j = 0; j < U; ++j) {
          virgin = virgin & (fc[j] == 0.0f) & (r == 0.0f);
          if (virgin) { fj[j] = 0.0f; pj[j] = 0.0f; qj[j] = 0.0f; }
        }
      }
      if (!virgin) { c_freq = fc[U - 1]; c_res = r; }
      f = fj[U - 1]; p = pj[U - 1]; q = qj[U - 1];
#pragma unroll
      for (int j = 0; j < U; ++j) { wf[(k0 + j) * L] = fj[j]; wp[(k0 + j) * L] = pj[j]; wq[(k0 + j) * L] = qj[j]; }
    });
  }
};

// ---- ADSRModule::calc, src/synth/adsr.rs:134-217 ------------------------------
struct AdsrOp {
  uint32_t* s;
  float phase, r_val, from_a_val;
  uint32_t mode;
  bool last;
  float s_val, inc_a, inc_d, inc_r, one_minus_s;
  Port p_gate, p_out;

  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    s = ln.st + ins.state * L;
    phase = __uint_as_float(s[0]); r_val = __uint_as_float(s[L]); from_a_val = __uint_as_float(s[2 * L]);
    mode = s[3 * L] & 0xFFu;
    last = (s[3 * L] >> 8) & 1u;
    const uint32_t* pp = ln.pr + ins.param * L;
    const float a_sec = __uint_as_float(pp[0]), d_sec = __uint_as_float(pp[L]), r_sec = __uint_as_float(pp[3 * L]);
    s_val = __uint_as_float(pp[2 * L]);
    const float sr = ins.imm;
    // `1.0 / (self.sample_rate * self.x_sec)` is loop invariant: same IEEE value every sample
    inc_a = __fdiv_rn(1.0f, fmul(sr, a_sec));
    inc_d = __fdiv_rn(1.0f, fmul(sr, d_sec));
    inc_r = __fdiv_rn(1.0f, fmul(sr, r_sec));
    one_minus_s = fsub(1.0f, s_val);
    p_gate = port(ln, ins.in[0]);
    p_out = port(ln, ins.out[0]);
  }
  __device__ __forceinline__ void store() {
    s[0] = __float_as_uint(phase); s[L] = __float_as_uint(r_val); s[2 * L] = __float_as_uint(from_a_val);
    s[3 * L] = mode | (last ? 1u << 8 : 0u);
  }

  // One sample of the five-arm `match self.mode` (:144-200) as selects: each arm's next
  // (mode, phase, r_val) is a few compares, and lanes in different modes cost nothing extra.
  __device__ __forceinline__ float step(float g, bool has_gate) {
    const bool above = g > 0.0f;
    const bool high = has_gate & above;         // `gate.is_some() && gate[i] > 0.0`
    const bool low = !has_gate | (g <= 0.0f);   // `gate.is_none() || gate[i] <= 0.0` (NaN is neither)
    const bool tr = above & !last;              // TransitionDetector, None => sees 0.0
    last = above;
    const bool m_none = mode == ADSR_NONE, m_att = mode == ADSR_ATTACK, m_dec = mode == ADSR_DECAY;
    const bool m_sus = mode == ADSR_SUSTAIN, m_rel = mode == ADSR_RELEASE;
    // Release restarts the attack first, then still advances by the release increment (:188-199)
    const bool rel_retrig = m_rel & high;
    const float ph0 = rel_retrig ? 0.0f : phase;
    const float inc = m_att ? inc_a : m_dec ? inc_d : inc_r;
    const float ph1 = fadd(ph0, inc);
    const bool ge = ph1 >= 1.0f;
    uint32_t nm = mode;
    nm = (m_none & high) ? ADSR_ATTACK : nm;
    nm = (m_att & ge) ? ADSR_DECAY : nm;
    nm = m_dec ? (tr ? ADSR_ATTACK : ge ? ADSR_SUSTAIN : ADSR_DECAY) : nm;
    nm = m_sus ? (tr ? ADSR_ATTACK : low ? ADSR_RELEASE : ADSR_SUSTAIN) : nm;
    nm = m_rel ? (ge ? ADSR_NONE : rel_retrig ? ADSR_ATTACK : ADSR_RELEASE) : nm;
    // next phase (None without a gate and Sustain keep theirs)
    const bool zero = (m_none & high) | ((m_att | m_dec) & (ge | tr)) | (m_sus & (low | tr)) | (m_rel & ge);
    const bool advance = m_att | m_dec | m_rel;
    const float np = zero ? 0.0f : advance ? ph1 : phase;
    // r_val: a retrigger during attack restarts from where the attack began; release end clears it
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

  // A group of samples in which nothing happens to the envelope's mode: no gate edge the
  // current mode reacts to and no phase wrap.  Then each arm is two or three flops per
  // sample (the same ones, in the same order, as step()).  Returns false -- with the
  // state untouched -- when something does happen; the caller then runs step().
  template <int U>
  __device__ __forceinline__ bool quiet(const float (&g)[U], bool has_gate, float (&o)[U]) {
    bool prev = last, any_tr = false, any_high = false, any_low = false;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const bool above = g[j] > 0.0f;
      any_tr |= above & !prev;
      any_high |= has_gate & above;
      any_low |= !has_gate | (g[j] <= 0.0f);
      prev = above;
    }
    float ph[U];
    bool ge = false;
    const float inc = mode == ADSR_ATTACK ? inc_a : mode == ADSR_DECAY ? inc_d : inc_r;
    float acc = phase;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      acc = fadd(acc, inc);
      ph[j] = acc;
      ge |= acc >= 1.0f;
    }
    if (mode == ADSR_SUSTAIN) {
      if (any_low | any_tr) return false;
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = s_val;
      r_val = s_val;
    } else if (mode == ADSR_NONE) {
      if (any_high) return false;
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = 0.0f;
      r_val = 0.0f;
    } else if (mode == ADSR_ATTACK) {
      if (ge | any_tr) return false;
      const float span = fsub(1.0f, r_val);
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = fadd(r_val, fmul(span, ph[j]));
      from_a_val = o[U - 1];
      phase = acc;
    } else if (mode == ADSR_DECAY) {
      if (ge | any_tr) return false;
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = fadd(s_val, fmul(one_minus_s, fsub(1.0f, ph[j])));
      r_val = o[U - 1];
      phase = acc;
    } else {  // Release
      if (ge | any_high) return false;
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = fmul(s_val, fsub(1.0f, ph[j]));
      r_val = o[U - 1];
      phase = acc;
    }
    last = prev;
    return true;
  }

  __device__ __forceinline__ void run(const Instr& ins, const Lane& ln, int kk) {
    const float* gate = p_gate.at(ln);
    float* out = p_out.at(ln);
    const bool has_gate = gate != nullptr;
    for_groups(kk, [&](auto u, int k0) {
      constexpr int U = decltype(u)::value;
      float g[U], o[U];
#pragma unroll
      for (int j = 0; j < U; ++j) g[j] = has_gate ? gate[(k0 + j) * L] : 0.0f;
      if (!quiet<U>(g, has_gate, o)) {
#pragma unroll
        for (int j = 0; j < U; ++j) o[j] = step(g[j], has_gate);
      }
      if (out) {
#pragma unroll
        for (int j = 0; j < U; ++j) out[(k0 + j) * L] = o[j];
      }
    });
  }
};

// ---- VCAModule::calc, src/synth/vca.rs:117-148 ---------------------------------
struct VcaOp {
  bool negative;
  Port p_audio, p_cv, p_out;
  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    negative = ins.flags & F_VCA_NEGATIVE;  // uniform over voices (vca.rs:14 has no per-voice meaning)
    p_audio = port(ln, ins.in[0]); p_cv = port(ln, ins.in[1]); p_out = port(ln, ins.out[0]);
  }
  __device__ __forceinline__ void store() {}
  __device__ __forceinline__ void run(const Instr& ins, const Lane& ln, int kk) {
    const float* audio = p_audio.at(ln);
    const float* cv = p_cv.at(ln);
    float* out = p_out.at(ln);
    if (!out) return;
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
      for (int j = 0; j < U; ++j) out[(k0 + j) * L] = (negative | (c[j] > 0.0f)) ? fmul(a[j], c[j]) : 0.0f;
    });
  }
};

// ---- MonoMixerModule::calc, src/synth/mixer.rs:101-122 -------------------------
struct MixerOp {
  float gain[4];
  Port p_in[4], p_out;
  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    const uint32_t* pp = ln.pr + ins.param * L;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gain[j] = __uint_as_float(pp[j * L]);
      p_in[j] = port(ln, ins.in[j]);
    }
    p_out = port(ln, ins.out[0]);
  }
  __device__ __forceinline__ void store() {}
  __device__ __forceinline__ void run(const Instr& ins, const Lane& ln, int kk) {
    float* out = p_out.at(ln);
    if (!out) return;
    const float* in0 = p_in[0].at(ln);
    const float* in1 = p_in[1].at(ln);
    const float* in2 = p_in[2].at(ln);
    const float* in3 = p_in[3].at(ln);
    for_groups(kk, [&](auto u, int k0) {
      constexpr int U = decltype(u)::value;
      float o[U];  // output.fill(0.0) then `*dst += src * gain` per connected input, in order
#pragma unroll
      for (int j = 0; j < U; ++j) o[j] = 0.0f;
      if (in0) {
#pragma unroll
        for (int j = 0; j < U; ++j) o[j] = fadd(o[j], fmul(in0[(k0 + j) * L], gain[0]));
      }
      if (in1) {
#pragma unroll
        for (int j = 0; j < U; ++j) o[j] = fadd(o[j], fmul(in1[(k0 + j) * L], gain[1]));
      }
      if (in2) {
#pragma unroll
        for (int j = 0; j < U; ++j) o[j] = fadd(o[j], fmul(in2[(k0 + j) * L], gain[2]));
      }
      if (in3) {
#pragma unroll
        for (int j = 0; j < U; ++j) o[j] = fadd(o[j], fmul(in3[(k0 + j) * L], gain[3]));
      }
#pragma unroll
      for (int j = 0; j < U; ++j) out[(k0 + j) * L] = o[j];
    });
  }
};

// ---- MathModule / NonLinearModule::calc, src/synth/math.rs:139-160, :292-313 ----
// math.rs:203-205 `if a > 0.0 { a.powf(b) } else { -(-a).powf(b) }` in f32: the platform's powf, restated
// operation by operation in libm_glibc.cuh (bit-identical to glibc 2.39).  Out of line: it is large.
static __device__ __noinline__ float nonlinear(float a, float b) {
  return a > 0.0f ? powf_glibc(a, b) : -powf_glibc(-a, b);
}

template <int WHICH>
__device__ __forceinline__ float math_op(float a, float b) {
  if (WHICH == F_MATH_ADD) return fadd(a, b);
  if (WHICH == F_MATH_SUB) return fsub(a, b);
  if (WHICH == F_MATH_MUL) return fmul(a, b);
  return nonlinear(a, b);
}

struct MathOp {
  float constant;
  Port p_a, p_b, p_out;
  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    constant = __uint_as_float(ln.pr[ins.param * L]);
    p_a = port(ln, ins.in[0]); p_b = port(ln, ins.in[1]); p_out = port(ln, ins.out[0]);
  }
  __device__ __forceinline__ void store() {}

  template <int WHICH>
  __device__ __forceinline__ void run_t(const Instr& ins, const Lane& ln, int kk) {
    float* out = p_out.at(ln);
    if (!out) return;
    const float* i1 = p_a.at(ln);
    const float* i2 = p_b.at(ln);
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
  __device__ __forceinline__ void run(const Instr& ins, const Lane& ln, int kk) {
    switch (ins.flags) {
      case F_MATH_ADD: run_t<F_MATH_ADD>(ins, ln, kk); break;
      case F_MATH_SUB: run_t<F_MATH_SUB>(ins, ln, kk); break;
      case F_MATH_MUL: run_t<F_MATH_MUL>(ins, ln, kk); break;
      default: run_t<F_MATH_NONLIN>(ins, ln, kk); break;
    }
  }
};

// ---- GridSequencerModule::calc, src/synth/sequencer.rs:190-246 --------------------------
// Inputs step (clock), sync; outputs cv, gate, sync.  The step table (Option<(u16, bool)> per
// step: -1 = None, else val | hold << 16) is the same for every voice and sits in shared memory;
// each voice has its own step counter, detectors and held CV.
struct GridSeqOp {
  uint32_t* s;
  uint32_t step;
  bool last_step, last_sync;
  float last_cv, inv_steps;
  const int32_t* table;
  uint32_t n_steps;
  Port p_step, p_sync, p_cv, p_gate, p_syncout;

  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    s = ln.st + ins.state * L;
    step = s[0] & 0xFFFFu;
    last_step = (s[0] >> 16) & 1u;
    last_sync = (s[0] >> 17) & 1u;
    last_cv = __uint_as_float(s[L]);
    inv_steps = ins.imm;
    table = ln.tables + ins.aux;
    n_steps = ins.n_ch;
    p_step = port(ln, ins.in[0]); p_sync = port(ln, ins.in[1]);
    p_cv = port(ln, ins.out[0]); p_gate = port(ln, ins.out[1]); p_syncout = port(ln, ins.out[2]);
  }
  __device__ __forceinline__ void store() {
    s[0] = step | (last_step ? 1u << 16 : 0u) | (last_sync ? 1u << 17 : 0u);
    s[L] = __float_as_uint(last_cv);
  }
  __device__ __forceinline__ void run(const Instr&, const Lane& ln, int kk) {
    const float* step_in = p_step.at(ln);
    const float* sync_in = p_sync.at(ln);
    float* cv = p_cv.at(ln);
    float* gate = p_gate.at(ln);
    float* sync_out = p_syncout.at(ln);
#pragma unroll 2
    for (int k = 0; k < kk; ++k) {
      const float st = step_in ? step_in[k * L] : 0.0f;
      const float sy = sync_in ? sync_in[k * L] : 0.0f;
      const bool a_step = st > 0.0f, a_sync = sy > 0.0f;
      step += (a_step & !last_step) ? 1u : 0u;     // :220-223
      step = (a_sync & !last_sync) ? 0u : step;    // :224-226
      last_step = a_step;
      last_sync = a_sync;
      step = step >= n_steps ? 0u : step;          // :227-230
      const int32_t cell = table[step];
      const bool some = cell >= 0;
      const float c = some ? fmul((float)(cell & 0xFFFF), inv_steps) : last_cv;  // :231-239
      const float g = some ? ((cell >> 16) & 1 ? 1.0f : st) : 0.0f;
      last_cv = c;
      if (cv) cv[k * L] = c;
      if (gate) gate[k * L] = g;
      if (sync_out) sync_out[k * L] = step == 0u ? 1.0f : 0.0f;
    }
  }
};

// ---- PatternSequencerModule::calc, src/synth/sequencer.rs:482-533 -----------------------
// 8 gate rows + sync = 9 output ports; one instruction covers ports first..first+2
// (ins.flags = first) and keeps its own copy of the step counter.  Table: rows x steps,
// -1 = None, 0 = Some(false) (pass the clock), 1 = Some(true) (hold).
struct PatSeqOp {
  uint32_t* s;
  uint32_t step;
  bool last_step, last_sync;
  const int32_t* table;
  uint32_t n_steps, first;
  Port p_step, p_sync, p_out[3];

  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    s = ln.st + ins.state * L;
    step = s[0] & 0xFFFFu;
    last_step = (s[0] >> 16) & 1u;
    last_sync = (s[0] >> 17) & 1u;
    table = ln.tables + ins.aux;
    n_steps = ins.n_ch;
    first = ins.flags;
    p_step = port(ln, ins.in[0]); p_sync = port(ln, ins.in[1]);
#pragma unroll
    for (int j = 0; j < 3; ++j) p_out[j] = port(ln, ins.out[j]);
  }
  __device__ __forceinline__ void store() {
    s[0] = step | (last_step ? 1u << 16 : 0u) | (last_sync ? 1u << 17 : 0u);
  }
  __device__ __forceinline__ void run(const Instr&, const Lane& ln, int kk) {
    const float* step_in = p_step.at(ln);
    const float* sync_in = p_sync.at(ln);
    float* out[3] = {p_out[0].at(ln), p_out[1].at(ln), p_out[2].at(ln)};
#pragma unroll 2
    for (int k = 0; k < kk; ++k) {
      const float st = step_in ? step_in[k * L] : 0.0f;
      const float sy = sync_in ? sync_in[k * L] : 0.0f;
      const bool a_step = st > 0.0f, a_sync = sy > 0.0f;
      step += (a_step & !last_step) ? 1u : 0u;
      step = (a_sync & !last_sync) ? 0u : step;
      last_step = a_step;
      last_sync = a_sync;
      step = step >= n_steps ? 0u : step;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (!out[j]) continue;
        const uint32_t portno = first + j;
        float v;
        if (portno == 8u) {
          v = step == 0u ? 1.0f : 0.0f;              // sync_out (:526)
        } else {
          const int32_t cell = table[portno * n_steps + step];
          v = cell < 0 ? 0.0f : (cell ? 1.0f : st);  // :515-524
        }
        out[j][k * L] = v;
      }
    }
  }
};

// ---- glibc 2.39 exp2f (sysdeps/ieee754/flt-32/e_exp2f.c), restated operation by operation ----
// The Sample module's playback rate is `ratio * 2.0_f32.powf(cv)` (sample.rs:234-235) -> exp2f of
// the platform libm, and its result steers an INDEX: a 1-ulp difference moves the play position
// and eventually picks another table entry, so "a few ulp" is not good enough here.  glibc's
// algorithm is short and all in f64: x = k/32 + r, 2^x = 2^(k/32) * (1 + C2 r + r^2 (C1 + C0 r)),
// rounded to f32 once at the end.  Checked on the CPU against glibc for every one of the 2^32
// inputs, with and without FMA contraction (both agree; tests/test_sample.py samples it, and the
// GPU test compares the device result with the oracle's glibc call).
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
__device__ __forceinline__ float exp2f_glibc(float x) {
  const double xd = (double)x;
  const double shift = 0x1.8p+52 / 32.0;
  double kd = dadd(xd, shift);
  const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
  kd = dsub(kd, shift);
  const double r = dsub(xd, kd);
  const unsigned long long t = __ldg(&kExp2fTab[ki & 31u]) + (ki << 47);
  const double s = __longlong_as_double((long long)t);
  const double z = dadd(dmul(0x1.c6af84b912394p-5, r), 0x1.ebfce50fac4f3p-3);
  const double r2 = dmul(r, r);
  double y = dadd(dmul(0x1.62e42ff0c52d6p-1, r), 1.0);
  y = dadd(dmul(z, r2), y);
  y = dmul(y, s);
  float out = __double2float_rn(y);
  // |x| >= 128 (and NaN): overflow, underflow, x + x  (e_exp2f.c special cases)
  out = x >= 128.0f ? __int_as_float(0x7f800000) : out;
  out = x <= -150.0f ? 0.0f : out;
  return x != x ? fadd(x, x) : out;
}

// ---- SampleModule::calc, src/synth/sample.rs:192-240 --------------------------------------
// Inputs gate, cv; one output.  The table (WaveBox.samples) is uniform over voices and stays in
// HBM -- a per-lane gather through the read-only path; per voice: play position (f32), playing,
// gate detector.  `pos as usize` is Rust's saturating cast (NaN / negative -> 0) == cvt.rzi.u32.
struct SampleOp {
  uint32_t* s;
  float pos, ratio;
  bool playing, last;
  const float* wave;
  uint32_t len;
  Port p_gate, p_cv, p_out;

  __device__ __forceinline__ void load(const Instr& ins, const Lane& ln) {
    s = ln.st + ins.state * L;
    pos = __uint_as_float(s[0]);
    playing = s[L] & 1u;
    last = (s[L] >> 1) & 1u;
    const WaveDesc* d = reinterpret_cast<const WaveDesc*>(ln.tables + ins.aux);
    wave = ln.waves + d->offset;
    len = d->len;
    ratio = d->ratio;
    if (d->is_new && ln.chunk == 0) { pos = 0.0f; playing = false; }  // :212-216, first block after a load
    p_gate = port(ln, ins.in[0]); p_cv = port(ln, ins.in[1]); p_out = port(ln, ins.out[0]);
  }
  __device__ __forceinline__ void store() {
    s[0] = __float_as_uint(pos);
    s[L] = (playing ? 1u : 0u) | (last ? 2u : 0u);
  }
  __device__ __forceinline__ void run(const Instr&, const Lane& ln, int kk) {
    const float* gate = p_gate.at(ln);
    const float* cv = p_cv.at(ln);
    float* out = p_out.at(ln);
#pragma unroll 2
    for (int k = 0; k < kk; ++k) {
      const float g = gate ? gate[k * L] : 0.0f;
      const bool above = g > 0.0f;
      const bool trigger = above & !last;              // :219-221
      last = above;
      pos = trigger ? 0.0f : pos;                      // :222-225
      playing |= trigger;
      uint32_t idx = __float2uint_rz(pos);
      const bool past = idx >= len;                    // :226-229
      pos = past ? 0.0f : pos;
      playing &= !past;
      idx = past ? 0u : idx;
      const float x = len ? __ldg(wave + idx) : 0.0f;  // :230-234
      if (out) out[k * L] = x;
      const float e = cv ? exp2f_glibc(cv[k * L]) : 1.0f;
      pos = playing ? fadd(pos, fmul(ratio, e)) : pos; // :235-238
    }
  }
};

}  // namespace dsp
}  // namespace srk
