// gpb_kernbwd.cuh -- parameter block shared by the covariance backward kernels (gpb_kern.cu, gpb_kern_mma.cuh).
#pragma once
#include "gpb_kernfn.cuh"

namespace gpb {

struct KbwdParams {
  int kind;
  const double* X1; int n1; long ldx1;
  const double* X2; int n2; long ldx2;
  int D;
  const double* ell; int ell_len;
  const double* sigma2;
  const double* G; long ldg; int g_trans;      // dense upstream gradient
  const double* Mul; long ldm;                 // optional element-wise multiplier of G (same layout as G)
  int symmetric;                               // X2 is X (only KERN_WHITE looks at it)
  const double* Kinv; long ldk; const double* kd; const double* a; int dy; long lda;  // GPR form
  int strips, nrc;
  double* part_h;    // [ncta][D + 2]: S_d ..., sum G*K/sigma2, tr W
  double* part_g2;   // [strips][n2][D] or nullptr
};

constexpr int BM_TM = 64, BM_TN = 128, BM_THREADS = 256, BM_HLD = 136, BM_DY = 8;   // DMMA path: tile, threads, a-chunk

}  // namespace gpb
