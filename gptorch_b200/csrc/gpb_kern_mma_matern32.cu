// gpb_kern_mma_matern32.cu -- the DMMA covariance backward kernels of one family (see gpb_kern_mma.cuh).
#include "gpb_kern_mma.cuh"

namespace gpb {
int kbwd_mma_launch_matern32(const KbwdParams& p, int ncb, bool gpr, bool g2, cudaStream_t stream) {
  return kbwd_mma_launch_family<KERN_MATERN32>(p, ncb, gpr, g2, stream);
}
}  // namespace gpb
