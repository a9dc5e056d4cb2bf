// SOLO instantiation of the voice kernel: one warp per 32-voice group runs the whole program chunk
// by chunk (the throughput schedule).  sm_100a only.
#include "voice_kernel.cuh"

namespace srk {

cudaError_t launch_voices_solo(const RenderArgs& a, unsigned grid, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(render_voices_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  render_voices_kernel<true><<<grid, 32, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace srk
