// k_wide.cu — warp-per-instance kernels (sim_core.cuh G = 32): small batches and large committees.
#include "kernels.cuh"
namespace lbft {
cudaError_t launch_wide(const KernelSel& k, const Params& P, cudaStream_t stream) {
  if (!k.wide || (k.fixed != FX_NONE && k.fixed != FX_COMMITTEE64) || k.rec || k.res) return cudaErrorInvalidValue;
  switch (k.qmode) {
    case 2: return k.smem ? launch_wide_groups<16, 2, true>(k, P, stream) : launch_wide_groups<16, 2, false>(k, P, stream);
    case 1: return launch_wide_groups<16, 1, false>(k, P, stream);
    case 3:
      if (k.nmax == 16) return launch_wide_groups<16, 3, false>(k, P, stream);
      if (k.nmax == 32) return launch_wide_groups<32, 3, false>(k, P, stream);
      return launch_wide_groups<64, 3, false>(k, P, stream);
    default:
      if (k.nmax == 16) return launch_wide_groups<16, 0, false>(k, P, stream);
      if (k.nmax == 32) return launch_wide_groups<32, 0, false>(k, P, stream);
      return launch_wide_groups<64, 0, false>(k, P, stream);
  }
}
}  // namespace lbft
