// k_fixed.cu — the bench kernel: four authors, default capacities, the reference's own delay model; layout is a compile-time
// constant (sim_core.cuh FIXED).  BASELINE configs[2].
#include "kernels.cuh"
namespace lbft {
cudaError_t launch_fixed(const KernelSel& k, const Params& P, cudaStream_t stream) {
  if (k.wide || !k.fixed || k.qmode != 2) return cudaErrorInvalidValue;
  constexpr int T = LaunchShape<2, true>::kThreads;
  const size_t dyn = (size_t)(T / 32) * q2_tile_words(64, true) * sizeof(uint32_t);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(lbft_event_loop_kernel<16, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  lbft_event_loop_kernel<16, 2, true><<<(P.num_instances + T - 1) / T, T, dyn, stream>>>(P);
  return cudaGetLastError();
}
}  // namespace lbft
