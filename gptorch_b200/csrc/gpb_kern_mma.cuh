// gpb_kern_mma.cuh -- covariance backward on the FP64 tensor pipe; included by one translation unit per family
// (gpb_kern_mma_<family>.cu) so that the 9 instantiations of each family compile in parallel.
#pragma once
#include "gpb_kernbwd.cuh"
#include <atomic>

namespace gpb {

// ================================================================================================
// backward reductions on the FP64 tensor pipe (stationary families, D <= 32)
// ================================================================================================
// The same reduction as kern_bwd_kernel, restructured so that the O(D) work per matrix element runs as DMMA on
// register-resident fragments instead of scalar code on shared-memory operands.  With h_ij = g_ij * fac(r_ij) * sigma2
// and scaled coordinates x~ = x / ell:
//    S_d   = sum_ij h_ij (x~_id - x~_jd)^2 = sum_i x~_id (x~_id u_i - 2 P_id) + sum_j x~_jd^2 v_j
//    P     = H X~2   (tile product on DMMA; the accumulator fragment of the distance tile IS the A operand: k-step e of
//                     fragment (i, j) takes the columns {2 kk + e}, whose values each lane already holds)
//    u_i   = sum_j h_ij,  v_j = sum_i h_ij     (thread-local partial sums)
//    gX2_jd ~ Pc_jd - x~_jd v_j,  Pc = H^T X~1 (needs H transposed: staged through shared memory, G2 only)
// -- the expansion the reference's autograd differentiates (gptorch/util.py:82-88), so the rounding behaviour is the
// reference's.  The distance tile itself is X~1 X~2^T on DMMA as in the forward kernel.  One CTA owns a block of 128
// columns and a strip of 64-row tiles; per-CTA partial sums are combined in a fixed order (deterministic).

// D <= 8 keeps a second copy of the column coordinates whose row stride makes the P-product operand loads
// bank-conflict free; wider D shares one copy (2-way conflicts on those few loads) so that two CTAs still fit an SM.
template <int DP, bool GPR>
static inline size_t kbwd_mma_smem_bytes() {
  size_t dbl = 64 * BM_HLD + 128 * (DP + 4) + (DP == 8 ? 128 * (DP + 2) : 0) + 64 * (DP + 4) + 128 + 64 + 32 + 8 * DP +
               2 * 128 + DP;
  if (GPR) dbl += 128 * BM_DY + 64 * BM_DY;
  return dbl * sizeof(double);
}

// cp.async of 16 / 8 bytes global -> shared; `bytes` valid source bytes (0 = fill the destination with zeros)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int KIND, int DP, bool GPR, bool G2>
__global__ void __launch_bounds__(BM_THREADS, (DP <= 16 ? 2 : 1)) kern_bwd_mma_kernel(const KbwdParams p) {
  constexpr int LDA = DP + 4;   // distance product: fragment rows r * LDA + kk hit 16 distinct banks
  constexpr bool SPLIT = DP == 8;
  constexpr int LDB = SPLIT ? DP + 2 : LDA;   // P product: rows (2 kk + e) * (DP + 2) + g hit 16 distinct banks
  constexpr int NF = DP / 8;
  extern __shared__ __align__(16) double bm_smem[];
  double* Gs = bm_smem;                  // [64][BM_HLD] upstream-gradient tile (cp.async), overwritten by h when G2
  double* X2a = Gs + 64 * BM_HLD;        // [128][LDA]  scaled column coordinates (B operand of the distance product)
  double* X2b = SPLIT ? X2a + 128 * LDA : X2a;   // [128][LDB]  second layout (B operand of P = H X~2)
  double* X1s = X2a + 128 * LDA + (SPLIT ? 128 * LDB : 0);   // [64][LDA]  scaled row coordinates of the current tile
  double* nb = X1s + 64 * LDA;           // [128] |x~_j|^2
  double* na = nb + 128;                 // [64]  |x~_i|^2
  double* red = na + 64;                 // [32]
  double* wS = red + 32;                 // [8][DP]  per-warp S_d
  double* vs = wS + 8 * DP;              // [2][128] per-warp-row column sums
  double* ellv = vs + 2 * 128;           // [DP]
  double* acol = ellv + DP;              // [BM_DY][128]  a_j  (GPR only; output-major: conflict-free fragment reads)
  double* arow = acol + 128 * BM_DY;     // [BM_DY][64]   a_i  (GPR only)
  double* Hs = Gs;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int wm = warp & 1, wn = warp >> 1;
  const int r = lane >> 2, kk = lane & 3;
  const int cb = blockIdx.x, strip = blockIdx.y;
  const int c0 = cb * BM_TN;
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int D = p.D;
  const int dy = GPR ? p.dy : 0;

  if (t < DP) ellv[t] = t < D ? 1.0 / p.ell[p.ell_len == 1 ? 0 : t] : 1.0;     // reciprocal length scales
  __syncthreads();
  for (int idx = t; idx < 128 * DP; idx += BM_THREADS) {
    const int cc = idx / DP, d = idx - cc * DP;
    double v = 0.0;
    if (d < D && c0 + cc < p.n2) v = p.X2[static_cast<long>(c0 + cc) * p.ldx2 + d] * ellv[d];
    X2a[cc * LDA + d] = v;
    if (SPLIT) X2b[cc * LDB + d] = v;
  }
  if (GPR) {
    for (int idx = t; idx < 128 * BM_DY; idx += BM_THREADS) {
      const int o = idx >> 7, cc = idx & 127;
      acol[idx] = (c0 + cc < p.n2 && o < dy) ? p.a[static_cast<long>(c0 + cc) * p.lda + o] : 0.0;
    }
  }
  __syncthreads();
  if (t < 128) {
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) s += X2a[t * LDA + d] * X2a[t * LDA + d];
    nb[t] = s;
  }
  const double sig2 = *p.sigma2;
  const bool vec_ok = GPR ? (((reinterpret_cast<uintptr_t>(p.Kinv) | reinterpret_cast<uintptr_t>(p.kd)) & 15) == 0 && (p.ldk & 1) == 0)
                          : (((reinterpret_cast<uintptr_t>(p.G) & 15) == 0) && ((p.ldg & 1) == 0));

  double Sacc[NF][2], vcol[4][2], Pc[2][NF][2];
#pragma unroll
  for (int nf = 0; nf < NF; ++nf) Sacc[nf][0] = Sacc[nf][1] = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) vcol[j][0] = vcol[j][1] = 0.0;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) Pc[q][nf][0] = Pc[q][nf][1] = 0.0;
  double sK = 0.0, gn = 0.0;

  const int ntr = (p.n1 + BM_TM - 1) / BM_TM;
  const int rt0 = GPR ? (c0 / BM_TM) : 0;     // GPR: lower triangle only, rows start at the column block
  for (int rt = rt0 + strip; rt < ntr; rt += p.strips) {
    const int m0 = rt * BM_TM;
    const bool same_blk = GPR && (m0 / NB == c0 / NB);
    __syncthreads();     // the previous tile's readers of X1s / arow / Gs are done
    // ---- upstream gradient of this tile: asynchronous copies into shared memory, consumed after the distance product
    {
      const double* gbase = GPR ? (same_blk ? p.kd : p.Kinv) : p.G;
      const long gld = GPR ? (same_blk ? static_cast<long>(NB) : p.ldk) : p.ldg;
      const int gcol0 = same_blk ? 0 : c0;
#pragma unroll 4
      for (int c = t; c < 64 * 64; c += BM_THREADS) {
        const int rr = c >> 6, cc = (c & 63) * 2;
        const int row = m0 + rr, col = c0 + cc;
        const double* src = gbase + static_cast<long>(row) * gld + gcol0 + cc;
        double* dst = Gs + rr * BM_HLD + cc;
        const bool rok = row < p.n1;
        if (vec_ok && rok && col + 1 < p.n2) {
          cp_async16(dst, src, 16);
        } else {
          cp_async8(dst, rok && col < p.n2 ? src : gbase, rok && col < p.n2 ? 8 : 0);
          cp_async8(dst + 1, rok && col + 1 < p.n2 ? src + 1 : gbase, rok && col + 1 < p.n2 ? 8 : 0);
        }
      }
    }
    for (int idx = t; idx < 64 * DP; idx += BM_THREADS) {
      const int rr = idx / DP, d = idx - rr * DP;
      double v = 0.0;
      if (d < D && m0 + rr < p.n1) v = p.X1[static_cast<long>(m0 + rr) * p.ldx1 + d] * ellv[d];
      X1s[rr * LDA + d] = v;
    }
    if (GPR) {
      for (int idx = t; idx < 64 * BM_DY; idx += BM_THREADS) {
        const int o = idx >> 6, rr = idx & 63;
        arow[idx] = (m0 + rr < p.n1 && o < dy) ? p.a[static_cast<long>(m0 + rr) * p.lda + o] : 0.0;
      }
    }
    cp_async_commit_wait_all();
    __syncthreads();     // X1s, arow, Gs
    if (t < 64) {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < DP; ++d) s += X1s[t * LDA + d] * X1s[t * LDA + d];
      na[t] = s;
    }
    __syncthreads();     // na

    // Two halves of 16 fragment rows each: halves the live accumulators (distance tile, P) so that two CTAs fit an SM.
#pragma unroll 1
    for (int ih = 0; ih < 2; ++ih) {
      // ---- distance tile: acc = X~1 X~2^T ----------------------------------------------------------------------
      double acc[2][4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < DP / 4; ++ks) {
        double a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = X1s[(wm * 32 + 8 * (2 * ih + i) + r) * LDA + ks * 4 + kk];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = X2a[(wn * 32 + 8 * j + r) * LDA + ks * 4 + kk];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }

      // ---- element-wise: h = g * fac * sigma2 (in place of acc), sK, tr W, row / column sums ------------------------
      double urow[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int lr = wm * 32 + 8 * (2 * ih + i) + r;
        const int row = m0 + lr;
        const double nrow = na[lr];
        double us = 0.0;
        double aa[4][2];
        if (GPR) {
          // a_i . a_j for the 8 elements of this fragment row: one pass over the outputs, row value loaded once
#pragma unroll
          for (int j = 0; j < 4; ++j) aa[j][0] = aa[j][1] = 0.0;
          for (int o = 0; o < dy; ++o) {
            const double ar = arow[o * 64 + lr];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double2 ac = *reinterpret_cast<const double2*>(acol + o * 128 + wn * 32 + 8 * j + 2 * kk);
              aa[j][0] = fma(ar, ac.x, aa[j][0]);
              aa[j][1] = fma(ar, ac.y, aa[j][1]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 g2v = *reinterpret_cast<const double2*>(Gs + lr * BM_HLD + wn * 32 + 8 * j + 2 * kk);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int lc = wn * 32 + 8 * j + 2 * kk + e;
            const int col = c0 + lc;
            const bool diag = (GPR || p.symmetric) && row == col;
            double r2 = (nrow + nb[lc]) - 2.0 * acc[i][j][e];
            r2 = fmax(r2, 0.0);
            r2 = diag ? 0.0 : r2;        // the distance of a point to itself is exactly 0 (as in the forward kernel)
            double kbase, fac1;
            kern_base_fac(KIND, r2, kbase, fac1);
            double g = e ? g2v.y : g2v.x;
            if (GPR) {
              const bool valid = row < p.n1 && col < p.n2 && col <= row;
              const double w = valid ? 0.5 * (static_cast<double>(dy) * g - aa[j][e]) : 0.0;
              gn += diag ? w : 0.0;
              g = diag ? w : 2.0 * w;
            }
            sK += g * kbase;
            const double h = diag ? 0.0 : g * fac1 * sig2;
            acc[i][j][e] = h;
            us += h;
            vcol[j][e] += h;
          }
        }
        us += __shfl_xor_sync(0xffffffffu, us, 1);
        us += __shfl_xor_sync(0xffffffffu, us, 2);
        urow[i] = us;
      }

      // ---- P = H X~2 over this warp's 32 columns, straight from the accumulator fragments -------------------------
      double Pacc[2][NF][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) Pacc[i][nf][0] = Pacc[i][nf][1] = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double bf[NF];
#pragma unroll
          for (int nf = 0; nf < NF; ++nf) bf[nf] = X2b[(wn * 32 + 8 * j + 2 * kk + e) * LDB + nf * 8 + r];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) dmma884(Pacc[i][nf][0], Pacc[i][nf][1], acc[i][j][e], bf[nf]);
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int lr = wm * 32 + 8 * (2 * ih + i) + r;
#pragma unroll
        for (int nf = 0; nf < NF; ++nf)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double xv = X1s[lr * LDA + nf * 8 + 2 * kk + e];
            Sacc[nf][e] += xv * (xv * urow[i] - 2.0 * Pacc[i][nf][e]);
          }
      }
      if (G2) {
        // each thread overwrites exactly the gradient elements it consumed itself
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2*>(Hs + (wm * 32 + 8 * (2 * ih + i) + r) * BM_HLD + wn * 32 + 8 * j + 2 * kk) =
                make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }  // halves

    if (G2) {
      // ---- Pc^T = X~1^T H over the 64 rows of the tile: H went through shared memory to change fragment roles -------
      __syncthreads();
#pragma unroll 4
      for (int ks = 0; ks < 16; ++ks) {
        double af[NF], bq[2];
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) af[nf] = X1s[(ks * 4 + kk) * LDA + nf * 8 + r];   // A[m = d][k = row]
#pragma unroll
        for (int q = 0; q < 2; ++q) bq[q] = Hs[(ks * 4 + kk) * BM_HLD + warp * 16 + 8 * q + r];   // B[k = row][n = col]
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int nf = 0; nf < NF; ++nf) dmma884(Pc[q][nf][0], Pc[q][nf][1], af[nf], bq[q]);
      }
    }
  }  // row tiles

  // ---- column sums v_j of this CTA (all its row tiles), deterministic order -------------------------------------------
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = vcol[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (r == 0) vs[wm * 128 + wn * 32 + 8 * j + 2 * kk + e] = v;
    }
  // ---- per-warp S_d: lanes with equal kk hold the same feature dimensions -----------------------------------------------
#pragma unroll
  for (int nf = 0; nf < NF; ++nf)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double s = Sacc[nf][e];
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      if (r == 0) wS[warp * DP + nf * 8 + 2 * kk + e] = s;
    }
  __syncthreads();
  double vj = 0.0;
  if (t < 128) vj = vs[t] + vs[128 + t];
  for (int d = 0; d < D; ++d) {       // uniform trip count
    double s = 0.0;
    if (t < 128) {
      const double x = X2a[t * LDA + d];
      s = x * x * vj;
    }
    s = block_sum(s, red);
    if (t == 0) {
      double tot = s;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += wS[w * DP + d];
      p.part_h[static_cast<long>(cta) * (D + 2) + d] = tot;
    }
  }
  sK = block_sum(sK, red);
  if (t == 0) p.part_h[static_cast<long>(cta) * (D + 2) + D] = sK;
  gn = block_sum(gn, red);
  if (t == 0) p.part_h[static_cast<long>(cta) * (D + 2) + D + 1] = gn;
  if (G2 && p.part_g2) {
    // Pc^T fragment: lane holds rows d = nf * 8 + r, columns j = warp * 16 + 8 q + 2 kk + e
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int lc = warp * 16 + 8 * q + 2 * kk + e;
        const int j = c0 + lc;
        if (j < p.n2) {
          const double vcj = vs[lc] + vs[128 + lc];
#pragma unroll
          for (int nf = 0; nf < NF; ++nf) {
            const int d = nf * 8 + r;
            if (d < D)
              p.part_g2[(static_cast<long>(strip) * p.n2 + j) * D + d] = Pc[q][nf][e] - X2a[lc * LDA + d] * vcj;
          }
        }
      }
  }
}

// ---- dispatch over (padded D, GPR / dense, with / without the column gradient) for ONE family ----------------------
template <int KIND, int DP, bool GPR, bool G2>
static int kbwd_mma_launch_one(const KbwdParams& p, int ncb, cudaStream_t stream) {
  static std::atomic<int> smem_state[GPB_MAX_DEVICES];
  const size_t smem = kbwd_mma_smem_bytes<DP, GPR>();
  if (int rc = ensure_dynamic_smem(kern_bwd_mma_kernel<KIND, DP, GPR, G2>, static_cast<int>(smem), smem_state)) return rc;
  dim3 grid(ncb, p.strips);
  kern_bwd_mma_kernel<KIND, DP, GPR, G2><<<grid, BM_THREADS, smem, stream>>>(p);
  count_launch();
  GPB_CUDA_CHECK(cudaGetLastError());
  return GPB_OK;
}

template <int KIND, bool GPR, bool G2>
static int kbwd_mma_launch_dp(const KbwdParams& p, int ncb, cudaStream_t stream) {
  if (p.D <= 8) return kbwd_mma_launch_one<KIND, 8, GPR, G2>(p, ncb, stream);
  if (p.D <= 16) return kbwd_mma_launch_one<KIND, 16, GPR, G2>(p, ncb, stream);
  return kbwd_mma_launch_one<KIND, 32, GPR, G2>(p, ncb, stream);
}

template <int KIND>
static int kbwd_mma_launch_family(const KbwdParams& p, int ncb, bool gpr, bool g2, cudaStream_t stream) {
  if (gpr) return kbwd_mma_launch_dp<KIND, true, false>(p, ncb, stream);
  return g2 ? kbwd_mma_launch_dp<KIND, false, true>(p, ncb, stream) : kbwd_mma_launch_dp<KIND, false, false>(p, ncb, stream);
}

}  // namespace gpb
