// Launch arguments of the voice kernel and the host-side launchers (each in its own translation unit:
// voice_kernel_solo.cu, voice_kernel_solo_full.cu, voice_kernel_pipelined.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "program.hpp"

namespace srk {

struct RenderArgs {
  const uint4* blob;      // [Instr x n_instr][WireDesc x n_wires][u16 warp_begin x (n_warps + 1)][pad][i32 tables]
  uint32_t* state;
  const uint32_t* params;
  float* rings;
  float* stems;
  float* partial;
  const float* waves;     // Sample modules' tables back to back (WaveDesc::offset indexes it)
  uint32_t blob_vec;      // blob size in uint4
  uint32_t table_off;     // byte offset of the sequencer tables inside the blob
  uint32_t n_instr, n_wires, n_warps, n_stages, n_tiles;
  uint32_t V;             // voices rendered by this launch
  uint32_t voice_offset;  // global index of voice 0 (noise key)
  uint32_t n_samples;
  uint32_t S, P, C, B;
  uint32_t K;             // samples per chunk: power of two <= 128
  uint32_t log2K;
  uint32_t ring_phase;    // absolute sample index of sample 0, mod B
  uint32_t n_abs;         // ... and its low 32 bits: fixes the mixdown's summation order (run_mix)
  uint32_t seed_lo, seed_hi;
  uint32_t solo_op_barrier;  // one-warp schedule, several groups per block: barrier after every instruction, not only per chunk
};

constexpr int kMaxThreads = kMaxWarps * 32;

cudaError_t launch_voices_solo(const RenderArgs& a, unsigned grid, unsigned threads, size_t smem, cudaStream_t stream);       // BASELINE modules only
cudaError_t launch_voices_solo_full(const RenderArgs& a, unsigned grid, unsigned threads, size_t smem, cudaStream_t stream);  // every module kind
cudaError_t launch_voices_pipelined(const RenderArgs& a, unsigned grid, unsigned threads, size_t smem, cudaStream_t stream);

}  // namespace srk
