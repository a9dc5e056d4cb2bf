// Launch arguments of a fused voice kernel (fused_ops.cuh / fused_gen.cpp).  Plain C types only: this
// header is compiled by the host (engine.cu), by nvcc and by NVRTC (which has no <stdint.h>).
#pragma once

#define SRK_FUSED_MAX_UNIFORM 512  /* parameter words (uniform over voices) that ride in the kernel arguments */
#define SRK_FUSED_TILE 32          /* samples per output tile: [32 samples][32 voices] f32 = one TMA box */

struct SrkFusedArgs {
  unsigned* state;         /* u32 [S][V], the interpreter's layout (program.hpp) */
  const unsigned* params;  /* u32 [P][V]; only the words marked per-voice are read */
  float* rings;            /* f32 [R][B][V] */
  float* stems;            /* f32 [C][N][V] or null */
  float* partial;          /* f32 [G][C][N] or null */
  const float* waves;      /* Sample tables back to back */
  const int* tables;       /* sequencer step tables / WaveDescs (Program::tables) */
  unsigned V, voice_offset, n_samples, C, B, ring_phase;
  unsigned n_abs;          /* absolute index of sample 0 (low 32 bits): fixes the mixdown's summation order */
  unsigned seed_lo, seed_hi;
  unsigned use_tma;        /* stems leave through the tensor map (V % 4 == 0), else per-lane stores */
  unsigned pad0, pad1;
  unsigned u[SRK_FUSED_MAX_UNIFORM];  /* parameter word w when it is uniform over voices */
};

/* CUtensorMap, opaque: f32 [C][N][V] stems, box {32 voices, 32 samples, 1 channel} */
struct alignas(64) SrkTensorMap {
  unsigned long long opaque[16];
};
