// SOLO instantiation of the voice kernel: one warp per 32-voice group (several groups per block) runs the whole program chunk
// by chunk (the throughput schedule), for programs made of the BASELINE modules only.  sm_100a only.
#ifdef SRK_SOLO_SAMPLE_GROUP
#define SRK_SAMPLE_GROUP SRK_SOLO_SAMPLE_GROUP
#endif
#include "voice_kernel.cuh"

namespace srk {

cudaError_t launch_voices_solo(const RenderArgs& a, unsigned grid, unsigned threads, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(render_voices_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  render_voices_kernel<true, false><<<grid, threads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace srk
