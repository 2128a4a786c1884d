// gpb_kern_mma_rbf.cu -- the DMMA covariance backward kernels of one family (see gpb_kern_mma.cuh).
#include "gpb_kern_mma.cuh"

namespace gpb {
int kbwd_mma_launch_rbf(const KbwdParams& p, int ncb, bool gpr, bool g2, cudaStream_t stream) {
  return kbwd_mma_launch_family<KERN_RBF>(p, ncb, gpr, g2, stream);
}
}  // namespace gpb
