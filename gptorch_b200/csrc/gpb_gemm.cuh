// gpb_gemm.cuh -- internal interface of the FP64 DMMA GEMM engine (see gpb_gemm.cu).
#pragma once
#include "gpb_common.cuh"

namespace gpb {

// Operand layouts (all matrices row-major):
//   GEMM_NT:  C[m][n] (+)= sum_k A[m][k] * B[n][k]     A is M x K, B is N x K   (SYRK / right-TRSM shape)
//   GEMM_TN:  C[m][n] (+)= sum_k A[k][m] * B[k][n]     A is K x M, B is K x N   (L^T-products, Gram sums)
//   GEMM_NN:  C[m][n] (+)= sum_k A[m][k] * B[k][n]     A is M x K, B is K x N
enum GemmMode { GEMM_NT = 0, GEMM_TN = 1, GEMM_NN = 2 };

enum : unsigned {
  GF_LOWER_TILES = 1u,   // only tiles that intersect the lower triangle (tile_n <= tile_m); M == N
  GF_KLO_M = 2u,         // k range starts at the tile's first row     (operand zero for k <  m0)
  GF_KHI_M = 4u,         // k range ends   at the tile's last row + 1  (operand zero for k >= m0 + BM)
  GF_KLO_N = 8u,         // same, keyed on the tile's column range
  GF_KHI_N = 16u,
  GF_DIAG_TO_WS = 32u,   // tiles in a diagonal 128-block are written to Cdiag (row m, col n - block start)
  GF_ROWS_INPLACE = 64u, // C overwrites the rows of A it is computed from (N <= 128): never split a row block over CTAs
};

struct GemmArgs {
  int M = 0, N = 0, K = 0;
  double alpha = 1.0, beta = 0.0;
  double* C = nullptr;
  long ldc = 0;
  long c_batch = 0;        // element stride of C between batches
  double* Cdiag = nullptr; // GF_DIAG_TO_WS target, N x 128 (ldd)
  long ldd = 0;
  int ax = 0, ay = 0;      // (col, row) element coordinates of A's view inside mapA
  int bx = 0, by = 0;      // same for B inside mapB
  int dax = 0, day = 0, dbx = 0, dby = 0;  // per-batch coordinate increments
  unsigned flags = 0;
  int batch = 1;
};

// Tensor maps: K-contiguous operands (A of NT/NN, B of NT) use boxes of 32 rows x 16 columns, MN-contiguous
// operands (A of TN, B of TN/NN) boxes of 16 x 16.  Launches with at most GEMM_SMALL_TILE_THRESHOLD 128x128
// tiles run with 64x64 tiles instead (4x the CTAs).
constexpr int GEMM_SMALL_TILE_THRESHOLD = 64;
constexpr int GEMM_SMALL_TILE_THRESHOLD_LOWER = 300;
constexpr int GEMM_HALF_TILE_THRESHOLD = 300;   // from this many 128x128 tiles on: 128x64 half tiles, two CTAs per SM
int gemm_launch(GemmMode mode, const CUtensorMap& mapA, const CUtensorMap& mapB, const GemmArgs& args,
                cudaStream_t stream);

inline int gemm_box_rows_a(GemmMode m) { return m == GEMM_TN ? 16 : 32; }
inline int gemm_box_rows_b(GemmMode m) { return m == GEMM_NT ? 32 : 16; }

}  // namespace gpb
