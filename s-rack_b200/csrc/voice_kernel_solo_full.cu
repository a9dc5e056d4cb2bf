// SOLO instantiation of the voice kernel with every module kind in the interpreter (sequencers,
// sample player): one warp per 32-voice group, plan order.  sm_100a only.
#ifdef SRK_SOLO_SAMPLE_GROUP
#define SRK_SAMPLE_GROUP SRK_SOLO_SAMPLE_GROUP
#endif
#include "voice_kernel.cuh"

namespace srk {

cudaError_t launch_voices_solo_full(const RenderArgs& a, unsigned grid, unsigned threads, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(render_voices_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  render_voices_kernel<true, true><<<grid, threads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace srk
