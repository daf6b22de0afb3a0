// k_scan.cu — thread-per-instance kernels over the scan queues (QMODE 1: 64-bit keys in HBM, QMODE 2: 32-bit keys in shared
// memory), committees of <= 5, with their recording / resumable variants.
#include "kernels.cuh"
namespace lbft {
cudaError_t launch_scan(const KernelSel& k, const Params& P, cudaStream_t stream) {
  if (k.wide || k.fixed) return cudaErrorInvalidValue;
  if (k.qmode == 2 && k.tile == 8 && !k.rec && !k.res && !k.epochs && !k.tds) return launch_sparse_tiles<16, 2, 8>(P, stream);
  if (k.tile != 32) return cudaErrorInvalidValue;
  if (k.qmode == 2) return launch_thread_variants<16, 2>(k, P, stream);
  if (k.qmode == 1) return launch_thread_variants<16, 1>(k, P, stream);
  return cudaErrorInvalidValue;
}
}  // namespace lbft
