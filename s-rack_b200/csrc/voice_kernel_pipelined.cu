// PIPELINED instantiation of the voice kernel: up to 16 warps per 32-voice group, one instruction
// per warp, resident module state (the latency schedule).  sm_100a only.
#include "voice_kernel.cuh"

namespace srk {

cudaError_t launch_voices_pipelined(const RenderArgs& a, unsigned grid, unsigned threads, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(render_voices_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  render_voices_kernel<false, true><<<grid, threads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace srk
