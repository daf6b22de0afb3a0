// k_wide.cu — warp-per-instance kernels (sim_core.cuh G = 32): small batches and large committees.
#include "kernels.cuh"
namespace lbft {
cudaError_t launch_wide(const KernelSel& k, const Params& P, cudaStream_t stream) {
  if (!k.wide || k.fixed || k.rec || k.res) return cudaErrorInvalidValue;
  switch (k.qmode) {
    case 2: return launch_wide_variant<16, 2>(P, stream);
    case 1: return launch_wide_variant<16, 1>(P, stream);
    case 3:
      if (k.nmax == 16) return launch_wide_variant<16, 3>(P, stream);
      if (k.nmax == 32) return launch_wide_variant<32, 3>(P, stream);
      return launch_wide_variant<64, 3>(P, stream);
    default:
      if (k.nmax == 16) return launch_wide_variant<16, 0>(P, stream);
      if (k.nmax == 32) return launch_wide_variant<32, 0>(P, stream);
      return launch_wide_variant<64, 0>(P, stream);
  }
}
}  // namespace lbft
