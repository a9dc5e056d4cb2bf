// The voice kernel: one lane per voice, the whole sample loop in-kernel (see engine.cu for the
// HBM / shared-memory layout and DESIGN.md §4 for the two schedules).  A header because the two
// instantiations are compiled as separate translation units (voice_kernel_solo.cu,
// voice_kernel_pipelined.cu): each is ~2 minutes of ptxas on its own, and keeping them apart also
// keeps each schedule's hot code contiguous (the instruction caches are the scarce resource).
#pragma once
#include <cuda_runtime.h>

#include "dsp.cuh"
#include "voice_args.hpp"

namespace srk {

// What the ring and output ops need beyond dsp::Lane.
struct GroupCtx {
  const RenderArgs& a;
  uint32_t v;         // this lane's voice (idle lanes shadow the last voice)
  uint32_t n_active;  // voices of this group that exist
  bool active;
  int lane;
  bool solo;          // one warp runs the whole program in plan order
  uint32_t group;     // voice group (of 32) this warp / block renders
};

// OP_RING_LOAD / OP_RING_STORE: the delayed (feedback) wires, rings f32 [R][B][V] in HBM.
__device__ __forceinline__ void run_ring_load(const Instr& ins, const dsp::Lane& ln, const GroupCtx& g, int kk) {
  const uint32_t n0 = ln.chunk * g.a.K;
  const RenderArgs& a = g.a;
  const float* ring = a.rings + (size_t)ins.aux * a.B * a.V + g.v;
  float* out = dsp::wire(ln, ins.out[0]);
  uint32_t idx = (a.ring_phase + n0) % a.B;
  for (int k = 0; k < kk; ++k) {
    out[k * 32] = ring[(size_t)idx * a.V];
    idx = idx + 1 == a.B ? 0 : idx + 1;
  }
}

__device__ __forceinline__ void run_ring_store(const Instr& ins, const dsp::Lane& ln, const GroupCtx& g, int kk) {
  const uint32_t n0 = ln.chunk * g.a.K;
  const RenderArgs& a = g.a;
  float* ring = a.rings + (size_t)ins.aux * a.B * a.V + g.v;
  const float* in = dsp::wire(ln, ins.in[0]);
  uint32_t idx = (a.ring_phase + n0) % a.B;
  for (int k = 0; k < kk; ++k) {
    if (g.active) ring[(size_t)idx * a.V] = in[k * 32];
    idx = idx + 1 == a.B ? 0 : idx + 1;
  }
}

// OP_OUTPUT: OutputModule::calc, src/synth/output.rs:46-60 -- bufs[c] = input c or zeros, for
// up to 4 channels; `bufs` here is the stems array [C][N][V] in HBM (one 128-byte line per
// warp, sample and channel; streaming stores).  Channels fed by the same wire load it once.
__device__ __forceinline__ void run_output(const Instr& ins, const dsp::Lane& ln, const GroupCtx& g, int kk) {
  const RenderArgs& a = g.a;
  const uint32_t n0 = ln.chunk * a.K;
  if (!a.stems || !g.active) return;
  const float* src[kOutputChannelsPerInstr];
  float* dst[kOutputChannelsPerInstr];
#pragma unroll
  for (int j = 0; j < kOutputChannelsPerInstr; ++j) {
    src[j] = j < ins.n_ch ? dsp::wire(ln, ins.in[j]) : nullptr;
    dst[j] = a.stems + ((size_t)(ins.aux + j) * a.n_samples + n0) * a.V + g.v;
  }
  const uint32_t V = a.V;  // offsets inside one chunk fit 32 bits as long as K * V < 2^32 (checked at launch)
  uint32_t row = 0;        // k0 * V, carried instead of recomputed
  dsp::for_groups(kk, [&](auto u, int k0) {
    constexpr int U = decltype(u)::value;
    float x[U];
    uint32_t off[U];
#pragma unroll
    for (int q = 0; q < U; ++q) off[q] = row + q * V;
    row += U * V;
#pragma unroll
    for (int j = 0; j < kOutputChannelsPerInstr; ++j) {
      if (j >= ins.n_ch) break;
      if (j == 0 || ins.in[j] != ins.in[j - 1]) {
#pragma unroll
        for (int q = 0; q < U; ++q) x[q] = src[j] ? src[j][(k0 + q) * 32] : 0.0f;
      }
#pragma unroll
      for (int q = 0; q < U; ++q) __stcs(dst[j] + off[q], x[q]);
    }
  });
}

// OP_MIX: this group's share of the mixdown, partial[group][c][n] = sum over the group's voices.
// Transposed read of the [K][32] tile, 32 sample rows at a time: lane r adds the 32 voices of sample row r, four voices
// per LDS.128, starting at the 16-byte chunk (absolute sample index) mod 8 -- the eight lanes of a quarter warp read eight
// different chunks (no bank conflict), and the order of the additions is a function of the absolute sample index alone:
// the mix has the same bits whatever the chunk length, however a render is cut into calls, and whichever schedule runs
// (the fused kernels add in exactly this order, fused_ops.cuh Out::flush).
__device__ __forceinline__ void run_mix(const Instr& ins, const dsp::Lane& ln, const GroupCtx& g, int kk) {
  const RenderArgs& a = g.a;
  const uint32_t n0 = ln.chunk * a.K;
  if (!a.partial || g.n_active == 0) return;
  const int lane = g.lane;
  if (g.solo) __syncwarp();  // the tile was written by this warp a moment ago
  for (int kb = 0; kb < kk; kb += 32) {
    const int r = kb + lane;  // this lane's sample row
    const uint32_t s8 = (a.n_abs + n0 + (uint32_t)r) & 7u;
    float sum = 0.0f;
    for (int j = 0; j < ins.n_ch; ++j) {
      const float* src = dsp::wire(ln, ins.in[j]);
      if (!src) {
        sum = 0.0f;
      } else if (j == 0 || ins.in[j] != ins.in[j - 1]) {
        float acc = 0.0f;
        if (r < kk) {
          const float4* row4 = reinterpret_cast<const float4*>(src - lane + r * 32);
          if (g.n_active == 32) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 x = row4[(s8 + k) & 7u];
              acc = dsp::fadd(dsp::fadd(dsp::fadd(dsp::fadd(acc, x.x), x.y), x.z), x.w);
            }
          } else {
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
              const uint32_t ch4 = (s8 + k) & 7u;
              const float4 x = row4[ch4];
              const uint32_t v0 = ch4 * 4u;
              acc = dsp::fadd(acc, v0 < g.n_active ? x.x : 0.0f);
              acc = dsp::fadd(acc, v0 + 1u < g.n_active ? x.y : 0.0f);
              acc = dsp::fadd(acc, v0 + 2u < g.n_active ? x.z : 0.0f);
              acc = dsp::fadd(acc, v0 + 3u < g.n_active ? x.w : 0.0f);
            }
          }
        }
        sum = acc;
      }  // else: same wire as the previous channel, same sums
      if (r < kk) a.partial[((size_t)g.group * a.C + ins.aux + j) * a.n_samples + n0 + r] = sum;
    }
  }
  if (g.solo) __syncwarp();
}

// A warp that owns ONE instruction never goes back to the interpreter: it loads the module's
// state into registers once, then loops run() + one block barrier per iteration, and stores
// the state at the end.  (Besides keeping state in registers this keeps each warp inside one
// contiguous piece of code: the interpreter's dispatch hops across all inlined op bodies and
// measured ~100 cycles of instruction fetch per hop, profiles/r01g_k8.)
template <class Body>
__device__ __forceinline__ void resident_loop(const Instr& ins, dsp::Lane& ln, const RenderArgs& a, uint32_t n_chunks,
                                              uint32_t n_iter, Body&& body) {
  for (uint32_t it = 0; it < n_iter; ++it) {
    const uint32_t chunk = it - ins.stage;
    if (chunk < n_chunks) {  // also false while it < stage (wraps)
      ln.chunk = chunk;
      body((int)min(a.K, a.n_samples - chunk * a.K));
    }
    __syncthreads();
  }
}

template <class Op>
__device__ __forceinline__ void run_resident(const Instr& ins, dsp::Lane& ln, const RenderArgs& a, uint32_t n_chunks,
                                             uint32_t n_iter) {
  Op op;
  op.load(ins, ln);
  resident_loop(ins, ln, a, n_chunks, n_iter, [&](int kk) { op.run(ins, ln, kk); });
  op.store();
}

template <class Op>
__device__ __forceinline__ void run_once(const Instr& ins, const dsp::Lane& ln, int kk) {
  Op op;
  op.load(ins, ln);
  op.run(ins, ln, kk);
  op.store();
}

// Two instantiations, two separate pieces of code: SOLO (one warp per voice group runs the whole
// program chunk by chunk; the throughput shape) carries only the interpreter, PIPELINED adds the
// resident single-instruction loops.  Keeping them apart keeps each one's hot code close together
// (the one-warp schedule lost 14 % when the resident variants grew the shared kernel, r01s).
// FULL adds the modules outside the BASELINE patches (sequencers, sample player) to the interpreter.
// The one-warp kernel is compiled twice, with and without them: its speed follows the size and
// layout of the whole switch (every added op body cost it 7 % even for programs that never
// execute it, profiles/r03a), so the common patches run the smaller image.  Instructions only a
// pipelined program contains (OP_OSC_DELTA, OP_MOOG_COEF) are left out of both one-warp images.
template <bool SOLO, bool FULL>
__global__ void __launch_bounds__(kMaxThreads, 1) render_voices_kernel(const RenderArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const Instr* prog = reinterpret_cast<const Instr*>(smem_raw);
  const WireDesc* wd = reinterpret_cast<const WireDesc*>(prog + a.n_instr);
  const uint16_t* warp_begin = reinterpret_cast<const uint16_t*>(wd + a.n_wires);
  // PIPELINED: the block's warps share one voice group's tables.  SOLO: every warp of the block is a
  // voice group of its own (own tables, same program), kept in step by the per-chunk barrier so that
  // the warps of an SM walk the same op bodies together -- the one-warp schedule is bound by
  // instruction fetch (33 % of its stall samples were "no instruction", profiles/r03c), and aligned
  // warps share the fetched lines.
  const uint32_t group = SOLO ? blockIdx.x * (blockDim.x >> 5) + wid : blockIdx.x;
  const uint32_t group_words = (a.S + a.P + a.n_tiles * a.K) * 32;
  uint32_t* st = reinterpret_cast<uint32_t*>(smem_raw + (size_t)a.blob_vec * 16) + (SOLO ? wid * group_words : 0u);
  uint32_t* pr = st + a.S * 32;
  float* tiles = reinterpret_cast<float*>(pr + a.P * 32);

  // stage the patch program (port/wire table) once per block
  for (uint32_t i = tid; i < a.blob_vec; i += blockDim.x) reinterpret_cast<uint4*>(smem_raw)[i] = a.blob[i];
  const uint32_t v0 = group * 32;
  const uint32_t n_active = v0 < a.V ? min(32u, a.V - v0) : 0u;  // 0: a spare warp of the last SOLO block
  const bool active = (uint32_t)lane < n_active;
  const uint32_t v = active ? v0 + lane : a.V - 1;  // idle lanes shadow the last voice, never store
  const uint32_t w0 = SOLO ? 0u : (uint32_t)wid, dw = SOLO ? 1u : a.n_warps;
  for (uint32_t w = w0; w < a.S; w += dw) st[w * 32 + lane] = a.state[(size_t)w * a.V + v];
  for (uint32_t w = w0; w < a.P; w += dw) pr[w * 32 + lane] = a.params[(size_t)w * a.V + v];
  __syncthreads();

  dsp::Lane ln{st + lane, pr + lane, tiles + lane, wd, a.K * 32, 0, a.voice_offset + v, a.seed_lo, a.seed_hi,
               reinterpret_cast<const int32_t*>(smem_raw + a.table_off), a.waves};
  const uint32_t pc0 = warp_begin[SOLO ? 0 : wid], pc1 = warp_begin[SOLO ? 1 : wid + 1];
  const uint32_t K = a.K;
  const uint32_t n_chunks = (a.n_samples + K - 1) / K;
  const uint32_t n_iter = n_chunks + a.n_stages - 1;
  const GroupCtx g{a, v, n_active, active, lane, SOLO, group};

  if (!SOLO && pc1 == pc0 + 1) {
    const Instr ins = prog[pc0];
    switch (ins.op) {
      case OP_OSC: {
        dsp::OscOp op;
        op.load(ins, ln);
        resident_loop(ins, ln, a, n_chunks, n_iter, [&](int kk) { op.run(ins, ln, kk); });
        if (op.owns_state(ins)) op.store();
        break;
      }
      case OP_MOOG: {  // the critical stage: one flat loop over the whole render (MoogOp::run_all)
        dsp::MoogOp op;
        op.load(ins, ln);
        if (K % dsp::kGroup == 0) {
          for (uint32_t it = 0; it < ins.stage; ++it) __syncthreads();
          op.run_all(ln, a.n_samples, [] { __syncthreads(); });
          for (uint32_t it = ins.stage + n_chunks; it < n_iter; ++it) __syncthreads();
        } else {
          resident_loop(ins, ln, a, n_chunks, n_iter, [&](int kk) { op.run(ins, ln, kk); });
        }
        op.store();
        break;
      }
      case OP_MOOG_COEF: run_resident<dsp::MoogCoefOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_GRIDSEQ: run_resident<dsp::GridSeqOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_OSC_DELTA: run_resident<dsp::OscDeltaOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_OSC_PHASE: run_resident<dsp::OscPhaseOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_OSC_SHAPE: run_resident<dsp::OscShapeOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_PATSEQ: run_resident<dsp::PatSeqOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_SAMPLE: run_resident<dsp::SampleOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_ADSR: run_resident<dsp::AdsrOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_NOISE: run_resident<dsp::NoiseOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_VCA: run_resident<dsp::VcaOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_MIXER: run_resident<dsp::MixerOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_MATH: run_resident<dsp::MathOp>(ins, ln, a, n_chunks, n_iter); break;
      case OP_OUTPUT: resident_loop(ins, ln, a, n_chunks, n_iter, [&](int kk) { run_output(ins, ln, g, kk); }); break;
      case OP_MIX: resident_loop(ins, ln, a, n_chunks, n_iter, [&](int kk) { run_mix(ins, ln, g, kk); }); break;
      case OP_RING_LOAD: resident_loop(ins, ln, a, n_chunks, n_iter, [&](int kk) { run_ring_load(ins, ln, g, kk); }); break;
      case OP_RING_STORE: resident_loop(ins, ln, a, n_chunks, n_iter, [&](int kk) { run_ring_store(ins, ln, g, kk); }); break;
      default: resident_loop(ins, ln, a, n_chunks, n_iter, [](int) {}); break;
    }
  } else {
    for (uint32_t it = 0; it < n_iter; ++it) {
      for (uint32_t pc = pc0; pc < pc1; ++pc) {
        const Instr& ins = prog[pc];  // stays in shared memory: fields are read where they are used
        const uint32_t chunk = it - ins.stage;
        if (chunk >= n_chunks) continue;  // also catches it < stage (wraps)
        const int kk = (int)min(K, a.n_samples - chunk * K);
        ln.chunk = chunk;
        switch (ins.op) {
          case OP_OSC: run_once<dsp::OscOp>(ins, ln, kk); break;
          case OP_MOOG: run_once<dsp::MoogOp>(ins, ln, kk); break;
          case OP_MOOG_COEF: if constexpr (!SOLO) run_once<dsp::MoogCoefOp>(ins, ln, kk); break;
          case OP_OSC_DELTA: if constexpr (!SOLO) run_once<dsp::OscDeltaOp>(ins, ln, kk); break;
          case OP_OSC_PHASE: if constexpr (!SOLO) run_once<dsp::OscPhaseOp>(ins, ln, kk); break;
          case OP_OSC_SHAPE: if constexpr (!SOLO) run_once<dsp::OscShapeOp>(ins, ln, kk); break;
          case OP_GRIDSEQ: if constexpr (FULL) run_once<dsp::GridSeqOp>(ins, ln, kk); break;
          case OP_PATSEQ: if constexpr (FULL) run_once<dsp::PatSeqOp>(ins, ln, kk); break;
          case OP_SAMPLE: if constexpr (FULL) run_once<dsp::SampleOp>(ins, ln, kk); break;
          case OP_ADSR: run_once<dsp::AdsrOp>(ins, ln, kk); break;
          case OP_NOISE: run_once<dsp::NoiseOp>(ins, ln, kk); break;
          case OP_VCA: run_once<dsp::VcaOp>(ins, ln, kk); break;
          case OP_MIXER: run_once<dsp::MixerOp>(ins, ln, kk); break;
          case OP_MATH: run_once<dsp::MathOp>(ins, ln, kk); break;
          case OP_RING_LOAD: run_ring_load(ins, ln, g, kk); break;
          case OP_RING_STORE: run_ring_store(ins, ln, g, kk); break;
          case OP_OUTPUT: run_output(ins, ln, g, kk); break;
          case OP_MIX: run_mix(ins, ln, g, kk); break;
          default: break;
        }
        if (SOLO && a.solo_op_barrier) __syncthreads();  // (every SOLO instruction has stage 0: uniform)
      }
      if (SOLO && blockDim.x == 32) __syncwarp(); else __syncthreads();
    }
  }
  __syncthreads();
  if (active)
    for (uint32_t w = w0; w < a.S; w += dw) a.state[(size_t)w * a.V + v] = st[w * 32 + lane];
}

}  // namespace srk
