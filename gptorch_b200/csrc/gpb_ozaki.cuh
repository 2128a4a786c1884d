// gpb_ozaki.cuh -- internal interface of the EXPERIMENTAL int8-sliced GEMM (gpb_ozaki.cu).
#pragma once
#include "gpb_common.cuh"

namespace gpb {

// Extended arguments (internal callers: the triangular products of gpb_potri_lower).
struct OzEx {
  int m = 0, n = 0, k = 0;
  double alpha = 1.0, beta = 0.0;
  // operands: logical A is m x k, logical B is n x k.  *_trans = 0: stored like that (row-major, K-major);
  // *_trans = 1: stored as the k x m (k x n) row-major array whose transpose is meant (the split transposes on the fly).
  // *_tri / *_row0 / *_col0 refer to the STORED array: element (r, c) counts as zero when col0 + c lies left of the
  // 128-block of row0 + r (upper triangular by blocks inside a buffer whose other blocks hold unrelated data).
  const double* A = nullptr; long lda = 0; int a_trans = 0, a_tri = 0, a_row0 = 0, a_col0 = 0;
  const double* B = nullptr; long ldb = 0; int b_trans = 0, b_tri = 0, b_row0 = 0, b_col0 = 0;
  double* C = nullptr; long ldc = 0;
  int lower = 0;            // skip 128-blocks right of the diagonal; rows of C are global rows gi0 + i, columns global j
  int gi0 = 0;
  double* Cdiag = nullptr;  // diagonal 128-blocks go here ([row][NB] layout of gpb_potri_lower) instead of into C
  int slices = 8;
  cudaStream_t stream = nullptr;
};

// C = beta C + alpha A B^T on the INT8 tensor path; GPB_ERR_UNSUPPORTED for shapes it does not take (callers fall back to
// the FP64 DMMA engine).
int gemm_ozaki_nt_ex(const OzEx& x);
int gemm_ozaki_nt(int m, int n, int k, double alpha, const double* A, long lda, const double* B, long ldb, double beta,
                  double* C, long ldc, int lower, int slices, cudaStream_t stream);
// slices configured for the blocked factorisations (0 = off, the default; GPB_OZAKI / gpb_ozaki_config)
int ozaki_slices();
void ozaki_set_slices(int s);
// thresholds of the hooks in gpb_chol.cu
constexpr int OZ_MIN_ROWS = 4096, OZ_MIN_COLS = 1024, OZ_MIN_K = 512, OZ_TRI_STRIP = 2048, OZ_TRTRI_MIN_S = 4096;

}  // namespace gpb
