// gpb_solve.cu -- HBM-bound pieces of the hot path: triangular solves with few right-hand sides,
// log-determinant / sum-of-squares reduction, and small matrix utilities.
//
// Reference call sites: functions.trtrs (gptorch/functions.py:71-76) as used for alpha = L^-1 (y - m)
// (gptorch/models/gpr.py:62, :106), functions.lt_log_determinant (gptorch/functions.py:61-68), the jitter
// add of functions.jit_op (:36) and the zero upper triangle torch.cholesky returns (:47).
//
// trsv: ONE launch per (<=4 right-hand sides).  CTA `i` owns the 128-row block i of the solution; it streams
// the tiles L[i][j] (forward) or L[j][i] (transposed) exactly once -- 4 n^2 bytes in total, the roofline of
// the operation -- and waits on a per-block ready flag published by the CTA that owns block j.  Row blocks are
// handed out by an atomic ticket so a waiting CTA only ever depends on CTAs that are already running.
#include "gpb_common.cuh"
#include <algorithm>

namespace gpb {

constexpr int TRSV_THREADS = 256;

__device__ __forceinline__ int ld_acquire_i32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_i32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// A 128 x 128 row-major tile T (ld) is held in registers as 16 rows per warp x 4 values per lane (lanes along the
// columns: coalesced).  Loading is separated from the product so that the tile of the NEXT column block is already
// in flight while the CTA still waits for that block's solution: the loads do not depend on it.
struct TileRegs { double v[16][4]; };

__device__ __forceinline__ void tile_load(TileRegs& tr, const double* __restrict__ T, long ld, int rows_valid,
                                          int cols_valid) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int rr = 0; rr < 16; ++rr) {
    const int r = w * 16 + rr;
    const double* trow = T + static_cast<long>(r) * ld;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int c = lane + 32 * m;
      tr.v[rr][m] = (r < rows_valid && c < cols_valid) ? __ldcs(trow + c) : 0.0;
    }
  }
}

// y[r] = sum_c T[r][c] * x[c][q]; subtract: accs[r][q] -= y, else outs[r][q] = y.
template <int KR>
__device__ __forceinline__ void tile_rowdot(const TileRegs& tr, const double (*xs)[KR], double (*accs)[KR],
                                            bool subtract, double (*outs)[KR]) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double xr[4][KR];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int q = 0; q < KR; ++q) xr[m][q] = xs[lane + 32 * m][q];
#pragma unroll
  for (int rr = 0; rr < 16; ++rr) {
    const int r = w * 16 + rr;
    double part[KR];
#pragma unroll
    for (int q = 0; q < KR; ++q) {
      part[q] = 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m) part[q] += tr.v[rr][m] * xr[m][q];
      part[q] = warp_sum(part[q]);
    }
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < KR; ++q) {
        if (subtract) accs[r][q] -= part[q];
        else outs[r][q] = part[q];
      }
    }
  }
}

template <int KR>
__global__ void __launch_bounds__(TRSV_THREADS) trsv_fwd_kernel(const double* __restrict__ L, long ldl, int n,
                                                                const double* __restrict__ dinv, double* B, long ldb,
                                                                int col0, int kr, int* flags, int* ticket) {
  __shared__ int s_id;
  __shared__ double xs[NB][KR];
  __shared__ double accs[NB][KR];
  __shared__ double outs[NB][KR];
  const int t = threadIdx.x;
  if (t == 0) s_id = atomicAdd(ticket, 1);
  __syncthreads();
  const int i = s_id;
  const int r0 = i * NB;
  const int nbi = min(NB, n - r0);
  for (int idx = t; idx < NB * KR; idx += TRSV_THREADS) {
    const int r = idx / KR, q = idx % KR;
    accs[r][q] = (r < nbi && q < kr) ? B[static_cast<long>(r0 + r) * ldb + col0 + q] : 0.0;
  }
  TileRegs tr;
  for (int j = 0; j < i; ++j) {
    tile_load(tr, L + static_cast<long>(r0) * ldl + j * NB, ldl, nbi, NB);   // in flight while waiting below
    if (t == 0) {
      while (ld_acquire_i32(&flags[j]) == 0) { __nanosleep(20); }
    }
    __syncthreads();
    for (int idx = t; idx < NB * KR; idx += TRSV_THREADS) {
      const int c = idx / KR, q = idx % KR;
      xs[c][q] = (q < kr) ? __ldcg(&B[static_cast<long>(j * NB + c) * ldb + col0 + q]) : 0.0;
    }
    __syncthreads();
    tile_rowdot<KR>(tr, xs, accs, true, outs);
  }
  __syncthreads();
  // x_i = Inv_ii * acc   (Inv padded with identity; zeros above the diagonal)
  tile_load(tr, dinv + static_cast<long>(r0) * NB, NB, NB, NB);
  for (int idx = t; idx < NB * KR; idx += TRSV_THREADS) {
    const int c = idx / KR, q = idx % KR;
    xs[c][q] = accs[c][q];
  }
  __syncthreads();
  tile_rowdot<KR>(tr, xs, accs, false, outs);
  __syncthreads();
  for (int idx = t; idx < NB * KR; idx += TRSV_THREADS) {
    const int r = idx / KR, q = idx % KR;
    if (r < nbi && q < kr) B[static_cast<long>(r0 + r) * ldb + col0 + q] = outs[r][q];
  }
  __threadfence();
  __syncthreads();
  if (t == 0) st_release_i32(&flags[i], 1);
}

// Transposed solve: x_i = Inv_ii^T (b_i - sum_{j>i} L[j][i]^T x_j), blocks processed from the last one.
template <int KR>
__global__ void __launch_bounds__(TRSV_THREADS) trsv_bwd_kernel(const double* __restrict__ L, long ldl, int n,
                                                                const double* __restrict__ dinv, double* B, long ldb,
                                                                int col0, int kr, int* flags, int* ticket) {
  __shared__ int s_id;
  __shared__ double xs[NB][KR];
  __shared__ double accs[NB][KR];
  __shared__ double part2[2][NB][KR];
  const int t = threadIdx.x;
  const int nblk = (n + NB - 1) / NB;
  if (t == 0) s_id = atomicAdd(ticket, 1);
  __syncthreads();
  const int i = nblk - 1 - s_id;
  const int c0 = i * NB;
  const int nbi = min(NB, n - c0);
  const int c = t & 127, half = t >> 7;
  for (int idx = t; idx < NB * KR; idx += TRSV_THREADS) {
    const int r = idx / KR, q = idx % KR;
    accs[r][q] = (r < nbi && q < kr) ? B[static_cast<long>(c0 + r) * ldb + col0 + q] : 0.0;
  }
  for (int j = nblk - 1; j > i; --j) {
    const int nbj = min(NB, n - j * NB);
    // this thread's 64 rows of the tile L[j][i] (column c): issued before the wait, they do not depend on x_j
    double lv[64];
    {
      const double* tile = L + static_cast<long>(j) * NB * ldl + c0 + c;
      const int rbeg = half * 64;
#pragma unroll
      for (int u = 0; u < 64; ++u) lv[u] = (c < nbi && rbeg + u < nbj) ? __ldcs(tile + static_cast<long>(rbeg + u) * ldl) : 0.0;
    }
    if (t == 0) {
      while (ld_acquire_i32(&flags[j]) == 0) { __nanosleep(20); }
    }
    __syncthreads();
    for (int idx = t; idx < NB * KR; idx += TRSV_THREADS) {
      const int r = idx / KR, q = idx % KR;
      xs[r][q] = (r < nbj && q < kr) ? __ldcg(&B[static_cast<long>(j * NB + r) * ldb + col0 + q]) : 0.0;
    }
    __syncthreads();
    double p[KR];
#pragma unroll
    for (int q = 0; q < KR; ++q) p[q] = 0.0;
#pragma unroll
    for (int u = 0; u < 64; ++u) {
#pragma unroll
      for (int q = 0; q < KR; ++q) p[q] += lv[u] * xs[half * 64 + u][q];
    }
#pragma unroll
    for (int q = 0; q < KR; ++q) part2[half][c][q] = p[q];
    __syncthreads();
    if (half == 0) {
#pragma unroll
      for (int q = 0; q < KR; ++q) accs[c][q] -= part2[0][c][q] + part2[1][c][q];
    }
    // next iteration's first __syncthreads orders these writes before xs/part2 are reused
  }
  __syncthreads();
  // x_i[c] = sum_{r >= c} Inv_ii[r][c] * acc[r]
  {
    double p[KR];
#pragma unroll
    for (int q = 0; q < KR; ++q) p[q] = 0.0;
    const double* tile = dinv + static_cast<long>(c0) * NB + c;
    const int rbeg = half * 64, rend = rbeg + 64;
#pragma unroll 8
    for (int r = rbeg; r < rend; ++r) {
      const double v = tile[r * NB];
#pragma unroll
      for (int q = 0; q < KR; ++q) p[q] += v * accs[r][q];
    }
#pragma unroll
    for (int q = 0; q < KR; ++q) part2[half][c][q] = p[q];
  }
  __syncthreads();
  if (half == 0 && c < nbi) {
    for (int q = 0; q < kr; ++q) B[static_cast<long>(c0 + c) * ldb + col0 + q] = part2[0][c][q] + part2[1][c][q];
  }
  __threadfence();
  __syncthreads();
  if (t == 0) st_release_i32(&flags[i], 1);
}

size_t trsv_workspace_bytes(int n) { return (static_cast<size_t>((n + NB - 1) / NB) + 1) * sizeof(int); }

int trsv_lower(const double* L, int n, long ldl, const double* dinv, double* B, int k, long ldb, int trans,
               void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n <= 0 || k <= 0) return GPB_OK;
  if (!L || !dinv || !B || ldl < n || ldb < k || !workspace || workspace_bytes < trsv_workspace_bytes(n))
    return GPB_ERR_BADARG;
  const int nblk = (n + NB - 1) / NB;
  int* flags = static_cast<int*>(workspace);
  int* ticket = flags + nblk;
  for (int col0 = 0; col0 < k; col0 += 4) {
    const int kr = std::min(4, k - col0);
    GPB_CUDA_CHECK(cudaMemsetAsync(workspace, 0, trsv_workspace_bytes(n), stream));
    if (trans == 0) {
      if (kr == 1) trsv_fwd_kernel<1><<<nblk, TRSV_THREADS, 0, stream>>>(L, ldl, n, dinv, B, ldb, col0, kr, flags, ticket);
      else if (kr == 2) trsv_fwd_kernel<2><<<nblk, TRSV_THREADS, 0, stream>>>(L, ldl, n, dinv, B, ldb, col0, kr, flags, ticket);
      else trsv_fwd_kernel<4><<<nblk, TRSV_THREADS, 0, stream>>>(L, ldl, n, dinv, B, ldb, col0, kr, flags, ticket);
    } else {
      if (kr == 1) trsv_bwd_kernel<1><<<nblk, TRSV_THREADS, 0, stream>>>(L, ldl, n, dinv, B, ldb, col0, kr, flags, ticket);
      else if (kr == 2) trsv_bwd_kernel<2><<<nblk, TRSV_THREADS, 0, stream>>>(L, ldl, n, dinv, B, ldb, col0, kr, flags, ticket);
      else trsv_bwd_kernel<4><<<nblk, TRSV_THREADS, 0, stream>>>(L, ldl, n, dinv, B, ldb, col0, kr, flags, ticket);
    }
    count_launch();
    GPB_CUDA_CHECK(cudaGetLastError());
  }
  return GPB_OK;
}

// ------------------------------------------------------------------------------------------------
// out[0] = sum log L_ii ; out[1] = sum V^2.  One CTA, fixed reduction tree => deterministic.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) logdet_sumsq_kernel(const double* __restrict__ L, int n, long ldl,
                                                            const double* __restrict__ V, int vrows, int k, long ldv,
                                                            double* __restrict__ out) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += log(L[static_cast<long>(i) * (ldl + 1)]);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[0] = s;
  double q = 0.0;
  if (V != nullptr) {
    const long total = static_cast<long>(vrows) * k;
    for (long idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const long r = idx / k, c = idx - r * k;
      const double v = V[r * ldv + c];
      q += v * v;
    }
  }
  q = block_sum(q, scratch);
  if (threadIdx.x == 0) out[1] = q;
}

int logdet_sumsq(const double* L, int n, long ldl, const double* V, int vrows, int k, long ldv, double* out,
                 cudaStream_t stream) {
  if (!out || n < 0 || (n > 0 && !L)) return GPB_ERR_BADARG;
  logdet_sumsq_kernel<<<1, 1024, 0, stream>>>(L, n, ldl, V, vrows, k, ldv, out);
  count_launch();
  GPB_CUDA_CHECK(cudaGetLastError());
  return GPB_OK;
}

// ------------------------------------------------------------------------------------------------
// out[c][o] (+)= sum_r A[r][c] * Y[r][o]   (A: rows x cols, cols contiguous; Y: rows x dy, dy small)
// The "A Y" statistic of the sparse models for a row panel A^T (gptorch/models/sparse_gpr.py:137): HBM-bound, A is
// read exactly once.  Thread = one column, two row halves per CTA; per-strip partials are summed in fixed order.
// ------------------------------------------------------------------------------------------------
constexpr int GT_THREADS = 256;
constexpr int GT_DY = 4;

__global__ void __launch_bounds__(GT_THREADS) gemv_t_kernel(const double* __restrict__ A, long rows, int cols, long lda,
                                                            const double* __restrict__ Y, int dy0, int dy, long ldy,
                                                            int strips, double* __restrict__ part) {
  __shared__ double comb[128][GT_DY];
  const int t = threadIdx.x, c = blockIdx.x * 128 + (t & 127), half = t >> 7;
  const int strip = blockIdx.y;
  const long per = (rows + strips - 1) / strips;
  const long r_begin = strip * per, r_end = min(rows, r_begin + per);
  const long mid = r_begin + (r_end - r_begin + 1) / 2;
  const long lo = half == 0 ? r_begin : mid, hi = half == 0 ? mid : r_end;
  double acc[GT_DY];
#pragma unroll
  for (int o = 0; o < GT_DY; ++o) acc[o] = 0.0;
  if (c < cols) {
    const double* ap = A + c;
    long r = lo;
    for (; r + 8 <= hi; r += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(ap + (r + u) * lda);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int o = 0; o < GT_DY; ++o)
          if (o < dy) acc[o] += v[u] * __ldg(Y + (r + u) * ldy + dy0 + o);
      }
    }
    for (; r < hi; ++r) {
      const double v = __ldcs(ap + r * lda);
#pragma unroll
      for (int o = 0; o < GT_DY; ++o)
        if (o < dy) acc[o] += v * __ldg(Y + r * ldy + dy0 + o);
    }
  }
  if (half == 1) {
#pragma unroll
    for (int o = 0; o < GT_DY; ++o) comb[t & 127][o] = acc[o];
  }
  __syncthreads();
  if (half == 0 && c < cols) {
#pragma unroll
    for (int o = 0; o < GT_DY; ++o)
      if (o < dy) part[(static_cast<long>(strip) * cols + c) * GT_DY + o] = acc[o] + comb[t & 127][o];
  }
}

__global__ void gemv_t_finalize_kernel(const double* __restrict__ part, int strips, int cols, int dy0, int dy,
                                       double beta, double* __restrict__ out, long ldo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cols * dy) return;
  const int c = idx / dy, o = idx - c * dy;
  double s = 0.0;
  for (int st = 0; st < strips; ++st) s += part[(static_cast<long>(st) * cols + c) * GT_DY + o];
  double* dst = out + static_cast<long>(c) * ldo + dy0 + o;
  *dst = (beta != 0.0 ? beta * *dst : 0.0) + s;
}

static inline int gemv_t_strips(long rows, int cols) {
  const int ncb = (cols + 127) / 128;
  long s = (4 * 148 + ncb - 1) / ncb;
  s = std::max<long>(1, std::min<long>(s, (rows + 255) / 256));
  return static_cast<int>(s);
}

size_t gemv_t_workspace_bytes(long rows, int cols) {
  return static_cast<size_t>(gemv_t_strips(rows, cols)) * cols * GT_DY * sizeof(double) + 64;
}

int gemv_t(const double* A, long rows, int cols, long lda, const double* Y, int dy, long ldy, double beta, double* out,
           long ldo, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (rows < 0 || cols <= 0 || dy <= 0) return GPB_ERR_BADARG;
  if (!A || !Y || !out || lda < cols || ldy < dy || ldo < dy) return GPB_ERR_BADARG;
  if (!workspace || workspace_bytes < gemv_t_workspace_bytes(rows, cols)) return GPB_ERR_BADARG;
  const int strips = gemv_t_strips(rows, cols);
  double* part = static_cast<double*>(workspace);
  for (int dy0 = 0; dy0 < dy; dy0 += GT_DY) {
    const int d = std::min(GT_DY, dy - dy0);
    dim3 grid((cols + 127) / 128, strips);
    gemv_t_kernel<<<grid, GT_THREADS, 0, stream>>>(A, rows, cols, lda, Y, dy0, d, ldy, strips, part);
    count_launch();
    GPB_CUDA_CHECK(cudaGetLastError());
    gemv_t_finalize_kernel<<<(cols * d + 255) / 256, 256, 0, stream>>>(part, strips, cols, dy0, d, beta, out, ldo);
    count_launch();
    GPB_CUDA_CHECK(cudaGetLastError());
  }
  return GPB_OK;
}

// ================================================================================================
// row-wise dot products: out[i] = sum_j A[i][j] B[i][j]  (variance epilogues of the sparse models)
// ================================================================================================
// One CTA per row, fixed-order reduction (deterministic).  HBM-bound: reads 16 * cols bytes per row once.
__global__ void __launch_bounds__(256) rowdot_kernel(const double* __restrict__ A, long lda, const double* __restrict__ B,
                                                     long ldb, int cols, double alpha, double beta,
                                                     double* __restrict__ out) {
  __shared__ double red[32];
  const long i = blockIdx.x;
  const double* a = A + i * lda;
  const double* b = B + i * ldb;
  double s = 0.0;
  const bool vec = (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0);
  if (vec) {
    const int c2 = cols >> 1;
    for (int j = threadIdx.x; j < c2; j += blockDim.x) {
      const double2 x = __ldcs(reinterpret_cast<const double2*>(a) + j);
      const double2 y = __ldcs(reinterpret_cast<const double2*>(b) + j);
      s = fma(x.x, y.x, s);
      s = fma(x.y, y.y, s);
    }
    if ((cols & 1) && threadIdx.x == 0) s = fma(a[cols - 1], b[cols - 1], s);
  } else {
    for (int j = threadIdx.x; j < cols; j += blockDim.x) s = fma(__ldcs(a + j), __ldcs(b + j), s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[i] = alpha * s + (beta != 0.0 ? beta * out[i] : 0.0);
}

int rowdot(const double* A, long lda, const double* B, long ldb, long rows, int cols, double alpha, double beta,
           double* out, cudaStream_t stream) {
  if (rows < 0 || cols < 0) return GPB_ERR_BADARG;
  if (rows == 0) return GPB_OK;
  if (!A || !B || !out || lda < cols || ldb < cols) return GPB_ERR_BADARG;
  if (rows > 2147483647L) return GPB_ERR_UNSUPPORTED;
  rowdot_kernel<<<static_cast<unsigned>(rows), 256, 0, stream>>>(A, lda, B, ldb, cols, alpha, beta, out);
  count_launch();
  GPB_CUDA_CHECK(cudaGetLastError());
  return GPB_OK;
}

// ================================================================================================
// panel times a few vectors, and its adjoint update (mean / gradient epilogues of the sparse models)
// ================================================================================================
constexpr int GN_DY = 4;      // right-hand sides per pass
constexpr int GN_ROWS = 4;    // rows per CTA (V is re-read from L1/L2 once per CTA)

// out[i][o] = sum_j A[i][j] V[j][o]   (A: rows x cols panel, V: cols x dy, dy <= GN_DY per launch)
__global__ void __launch_bounds__(256) gemv_n_kernel(const double* __restrict__ A, long rows, int cols, long lda,
                                                     const double* __restrict__ V, int dy0, int dy, long ldv,
                                                     double* __restrict__ out, long ldo) {
  __shared__ double red[32];
  const long r0 = static_cast<long>(blockIdx.x) * GN_ROWS;
  double acc[GN_ROWS][GN_DY];
#pragma unroll
  for (int q = 0; q < GN_ROWS; ++q)
#pragma unroll
    for (int o = 0; o < GN_DY; ++o) acc[q][o] = 0.0;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    double v[GN_DY];
#pragma unroll
    for (int o = 0; o < GN_DY; ++o) v[o] = o < dy ? __ldg(V + static_cast<long>(j) * ldv + dy0 + o) : 0.0;
#pragma unroll
    for (int q = 0; q < GN_ROWS; ++q) {
      if (r0 + q < rows) {
        const double a = __ldcs(A + (r0 + q) * lda + j);
#pragma unroll
        for (int o = 0; o < GN_DY; ++o) acc[q][o] = fma(a, v[o], acc[q][o]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < GN_ROWS; ++q)
#pragma unroll
    for (int o = 0; o < GN_DY; ++o) {
      if (o >= dy) break;                       // uniform
      const double s = block_sum(acc[q][o], red);
      if (threadIdx.x == 0 && r0 + q < rows) out[(r0 + q) * ldo + dy0 + o] = s;
    }
}

int gemv_n(const double* A, long rows, int cols, long lda, const double* V, int dy, long ldv, double* out, long ldo,
           cudaStream_t stream) {
  if (rows < 0 || cols <= 0 || dy <= 0) return GPB_ERR_BADARG;
  if (rows == 0) return GPB_OK;
  if (!A || !V || !out || lda < cols || ldv < dy || ldo < dy) return GPB_ERR_BADARG;
  const long ctas = (rows + GN_ROWS - 1) / GN_ROWS;
  if (ctas > 2147483647L) return GPB_ERR_UNSUPPORTED;
  for (int dy0 = 0; dy0 < dy; dy0 += GN_DY) {
    gemv_n_kernel<<<static_cast<unsigned>(ctas), 256, 0, stream>>>(A, rows, cols, lda, V, dy0, std::min(GN_DY, dy - dy0), ldv,
                                                                  out, ldo);
    count_launch();
    GPB_CUDA_CHECK(cudaGetLastError());
  }
  return GPB_OK;
}

// A[i][j] <- scale * s[i] * A[i][j] + sum_o G[i][o] V[j][o]   (s may be NULL = 1; G / V may be NULL = no outer product)
__global__ void __launch_bounds__(256) rows_scale_add_outer_kernel(double* __restrict__ A, long rows, int cols, long lda,
                                                                   const double* __restrict__ s, double scale,
                                                                   const double* __restrict__ G, int dy, long ldg,
                                                                   const double* __restrict__ V, long ldv) {
  const long i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  double v = scale * (s ? s[i] : 1.0) * A[i * lda + j];
  if (G) {
    for (int o = 0; o < dy; ++o) v = fma(G[i * ldg + o], __ldg(V + static_cast<long>(j) * ldv + o), v);
  }
  A[i * lda + j] = v;
}

int rows_scale_add_outer(double* A, long rows, int cols, long lda, const double* s, double scale, const double* G, int dy,
                         long ldg, const double* V, long ldv, cudaStream_t stream) {
  if (rows < 0 || cols <= 0) return GPB_ERR_BADARG;
  if (rows == 0) return GPB_OK;
  if (!A || lda < cols || (G && (!V || dy <= 0 || ldg < dy || ldv < dy))) return GPB_ERR_BADARG;
  const long per = 65535;                       // gridDim.y limit
  for (long r0 = 0; r0 < rows; r0 += per) {
    const long nr = std::min(per, rows - r0);
    dim3 grid((cols + 255) / 256, static_cast<unsigned>(nr));
    rows_scale_add_outer_kernel<<<grid, 256, 0, stream>>>(A + r0 * lda, nr, cols, lda, s ? s + r0 : nullptr, scale,
                                                         G ? G + r0 * ldg : nullptr, dy, ldg, V, ldv);
    count_launch();
    GPB_CUDA_CHECK(cudaGetLastError());
  }
  return GPB_OK;
}

__global__ void tri_zero_upper_kernel(double* __restrict__ A, int n, long lda) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c < n && c > r) A[static_cast<long>(r) * lda + c] = 0.0;
}

int tri_zero_upper(double* A, int n, long lda, cudaStream_t stream) {
  if (n <= 0) return GPB_OK;
  if (!A || lda < n) return GPB_ERR_BADARG;
  dim3 grid((n + 255) / 256, n);
  tri_zero_upper_kernel<<<grid, 256, 0, stream>>>(A, n, lda);
  count_launch();
  GPB_CUDA_CHECK(cudaGetLastError());
  return GPB_OK;
}

__global__ void add_diag_kernel(double* __restrict__ A, int n, long lda, const double* __restrict__ value,
                                double host_value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[static_cast<long>(i) * (lda + 1)] += value ? *value : host_value;
}

int add_diag(double* A, int n, long lda, const double* value, double host_value, cudaStream_t stream) {
  if (n <= 0) return GPB_OK;
  if (!A || lda < n) return GPB_ERR_BADARG;
  add_diag_kernel<<<(n + 255) / 256, 256, 0, stream>>>(A, n, lda, value, host_value);
  count_launch();
  GPB_CUDA_CHECK(cudaGetLastError());
  return GPB_OK;
}

}  // namespace gpb
