// gpb_kern_sop.cu -- composite covariance in ONE pass: K = sum_t prod_{l in t} k_l(X, X2)  (+ noise I).
//
// Replaces the Sum / Product combinators of the reference (gptorch/kernels.py:286-306) over its leaf kernels
// (Rbf/Exp/Matern/Periodic :182-235, Linear :238-265, Constant/White :83-101), which materialise one N x N tensor per
// child plus one per combinator (and ~4 temporaries inside every stationary child).  Any tree of + and * expands into
// a sum of products of leaves; this kernel evaluates that normal form tile by tile: for every leaf the scaled
// X.X'^T term runs on FP64 DMMA (each leaf has its own length scales), the leaf value is formed in the accumulator
// epilogue, multiplied into the term's running product and added to the tile total, all in registers.  One N^2
// write, no temporaries.  BASELINE config #1 (examples/regression_1d.py:42) is Linear + Rbf + Constant.
#include "gpb_kernfn.cuh"

namespace gpb {

constexpr int SP_TM = 64, SP_TN = 128;        // CTA tile: 2 x 4 warps of 32 x 32
constexpr int SP_MI = 4, SP_NI = 4;           // 8x8 DMMA fragments per warp tile
constexpr int SP_DC = 16, SP_LD = 20;         // feature chunk, padded smem row stride (conflict-free fragment loads)
constexpr int SP_THREADS = 256;
constexpr int SP_MAX_LEAVES = 16, SP_MAX_TERMS = 8;

struct SopLeaf {
  int kind, ell_len;
  const double* ell;       // length scales (Linear: per-dimension variances); unused for Constant / White
  const double* sigma2;    // unused for Linear
};

struct SopParams {
  const double* X; int n1; long ldx;
  const double* X2; int n2; long ldx2;
  int D;
  int nterms;
  int term_end[SP_MAX_TERMS];    // leaves of term t: [term_end[t-1], term_end[t])
  SopLeaf leaf[SP_MAX_LEAVES];
  const double* noise;
  int symmetric, lower;
  double* K; long ldk;
  int tiles_n;
};

// 96 fp64 values per thread live across the leaf loop (accumulators, running product, total): one CTA per SM.
__global__ void __launch_bounds__(SP_THREADS, 1) kern_sop_fwd_kernel(const __grid_constant__ SopParams p) {
  __shared__ double As[SP_TM * SP_LD];
  __shared__ double Bs[SP_TN * SP_LD];
  __shared__ double na[SP_TM], nbv[SP_TN];
  __shared__ double scale[SP_DC];

  int tm, tn;
  {
    const int t = blockIdx.x;
    if (p.lower) {   // pairs of 64-row tile rows (2q, 2q+1) each cover column tiles 0 .. q
      int q = static_cast<int>((sqrt(4.0 * t + 1.0) - 1.0) * 0.5);
      while ((q + 1) * (q + 2) <= t) ++q;
      while (q * (q + 1) > t) --q;
      const int rem = t - q * (q + 1);
      tm = 2 * q + rem / (q + 1);
      tn = rem % (q + 1);
    } else {
      tm = t / p.tiles_n;
      tn = t - tm * p.tiles_n;
    }
  }
  const int m0 = tm * SP_TM, n0 = tn * SP_TN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 1, wn = warp >> 1;
  const int r = lane >> 2, kk = lane & 3;

  double tot[SP_MI][SP_NI][2], prod[SP_MI][SP_NI][2];
#pragma unroll
  for (int i = 0; i < SP_MI; ++i)
#pragma unroll
    for (int j = 0; j < SP_NI; ++j) tot[i][j][0] = tot[i][j][1] = 0.0;

  int l = 0;
  for (int t = 0; t < p.nterms; ++t) {
#pragma unroll
    for (int i = 0; i < SP_MI; ++i)
#pragma unroll
      for (int j = 0; j < SP_NI; ++j) prod[i][j][0] = prod[i][j][1] = 1.0;
    for (; l < p.term_end[t]; ++l) {
      const int kind = p.leaf[l].kind;
      if (kind >= KERN_CONSTANT) {
        const double s2 = *p.leaf[l].sigma2;
        const bool white = kind == KERN_WHITE;
#pragma unroll
        for (int i = 0; i < SP_MI; ++i) {
          const int row = m0 + wm * 32 + 8 * i + r;
#pragma unroll
          for (int j = 0; j < SP_NI; ++j) {
            const int col = n0 + wn * 32 + 8 * j + 2 * kk;
#pragma unroll
            for (int e = 0; e < 2; ++e) prod[i][j][e] *= (!white || (p.symmetric && row == col + e)) ? s2 : 0.0;
          }
        }
        continue;
      }
      const bool linear = kind == KERN_LINEAR;
      const double* ell = p.leaf[l].ell;
      const int ell_len = p.leaf[l].ell_len;
      double acc[SP_MI][SP_NI][2];
#pragma unroll
      for (int i = 0; i < SP_MI; ++i)
#pragma unroll
        for (int j = 0; j < SP_NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      double my_n = 0.0;
      for (int d0 = 0; d0 < p.D; d0 += SP_DC) {
        __syncthreads();
        if (tid < SP_DC) {
          const int d = d0 + tid;
          scale[tid] = d < p.D ? ell[ell_len == 1 ? 0 : d] : 1.0;
        }
        __syncthreads();
        for (int idx = tid; idx < (SP_TM + SP_TN) * SP_DC; idx += SP_THREADS) {
          const int row = idx >> 4, k = idx & 15;
          const int d = d0 + k;
          double v = 0.0;
          if (row < SP_TM) {
            if (d < p.D && m0 + row < p.n1) {
              const double x = p.X[static_cast<long>(m0 + row) * p.ldx + d];
              v = linear ? x * scale[k] : x / scale[k];      // Linear: (X * v) X'^T (gptorch/kernels.py:260-262)
            }
            As[row * SP_LD + k] = v;
          } else {
            const int rb = row - SP_TM;
            if (d < p.D && n0 + rb < p.n2) {
              const double x = p.X2[static_cast<long>(n0 + rb) * p.ldx2 + d];
              v = linear ? x : x / scale[k];
            }
            Bs[rb * SP_LD + k] = v;
          }
        }
        __syncthreads();
        if (tid < SP_TM + SP_TN) {
          const double* src = tid < SP_TM ? &As[tid * SP_LD] : &Bs[(tid - SP_TM) * SP_LD];
#pragma unroll
          for (int k = 0; k < SP_DC; ++k) my_n += src[k] * src[k];
        }
        const int ksteps = min(SP_DC, p.D - d0 + 3) / 4;
        for (int ks = 0; ks < ksteps; ++ks) {
          double a[SP_MI], b[SP_NI];
#pragma unroll
          for (int i = 0; i < SP_MI; ++i) a[i] = As[(wm * 32 + 8 * i + r) * SP_LD + ks * 4 + kk];
#pragma unroll
          for (int j = 0; j < SP_NI; ++j) b[j] = Bs[(wn * 32 + 8 * j + r) * SP_LD + ks * 4 + kk];
#pragma unroll
          for (int i = 0; i < SP_MI; ++i)
#pragma unroll
            for (int j = 0; j < SP_NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
      }
      __syncthreads();   // everyone is done with the norms of the previous leaf
      if (tid < SP_TM) na[tid] = my_n;
      else if (tid < SP_TM + SP_TN) nbv[tid - SP_TM] = my_n;
      __syncthreads();
      const double sig2 = linear ? 1.0 : *p.leaf[l].sigma2;
#pragma unroll
      for (int i = 0; i < SP_MI; ++i) {
        const int lr = wm * 32 + 8 * i + r;
        const int row = m0 + lr;
        const double nrow = na[lr];
#pragma unroll
        for (int j = 0; j < SP_NI; ++j) {
          const int lc = wn * 32 + 8 * j + 2 * kk;
          const int col = n0 + lc;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double dot = acc[i][j][e];
            double v;
            if (linear) {
              v = dot;
            } else {
              double r2 = fmax((nrow + nbv[lc + e]) - 2.0 * dot, 0.0);   // gptorch/util.py:84-88
              if (p.symmetric && row == col + e) r2 = 0.0;               // a point's distance to itself is exactly 0
              v = sig2 * kern_base(kind, r2);
            }
            prod[i][j][e] *= v;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < SP_MI; ++i)
#pragma unroll
      for (int j = 0; j < SP_NI; ++j) {
        tot[i][j][0] += prod[i][j][0];
        tot[i][j][1] += prod[i][j][1];
      }
  }

  const double noise = (p.symmetric && p.noise) ? *p.noise : 0.0;
#pragma unroll
  for (int i = 0; i < SP_MI; ++i) {
    const int row = m0 + wm * 32 + 8 * i + r;
    if (row >= p.n1) continue;
    double* krow = p.K + static_cast<long>(row) * p.ldk;
#pragma unroll
    for (int j = 0; j < SP_NI; ++j) {
      const int col = n0 + wn * 32 + 8 * j + 2 * kk;
      double v0 = tot[i][j][0], v1 = tot[i][j][1];
      if (p.symmetric && row == col) v0 += noise;
      if (p.symmetric && row == col + 1) v1 += noise;
      if (col + 1 < p.n2 && ((p.ldk & 1) == 0)) {
        __stcs(reinterpret_cast<double2*>(krow + col), make_double2(v0, v1));
      } else {
        if (col < p.n2) krow[col] = v0;
        if (col + 1 < p.n2) krow[col + 1] = v1;
      }
    }
  }
}

int kern_sop_fwd(int n_terms, const int* term_len, const int* leaf_kind, const double* const* leaf_ell,
                 const int* leaf_ell_len, const double* const* leaf_sigma2, const double* X, int n1, long ldx,
                 const double* X2, int n2, long ldx2, int D, const double* noise, int fill, double* K, long ldk,
                 cudaStream_t stream) {
  if (n_terms <= 0 || n_terms > SP_MAX_TERMS || !term_len || !leaf_kind || !leaf_ell || !leaf_ell_len || !leaf_sigma2)
    return GPB_ERR_BADARG;
  if (!X || !K || D <= 0) return GPB_ERR_BADARG;
  SopParams p;
  int nl = 0;
  for (int t = 0; t < n_terms; ++t) {
    if (term_len[t] <= 0) return GPB_ERR_BADARG;
    nl += term_len[t];
    if (nl > SP_MAX_LEAVES) return GPB_ERR_UNSUPPORTED;
    p.term_end[t] = nl;
  }
  for (int l = 0; l < nl; ++l) {
    const int kind = leaf_kind[l];
    if (kind < 0 || kind > KERN_WHITE) return GPB_ERR_BADARG;
    if (kind != KERN_LINEAR && !leaf_sigma2[l]) return GPB_ERR_BADARG;
    if (kind < KERN_CONSTANT) {
      if (!leaf_ell[l] || (leaf_ell_len[l] != 1 && leaf_ell_len[l] != D)) return GPB_ERR_BADARG;
      if (kind == KERN_LINEAR && leaf_ell_len[l] != D) return GPB_ERR_BADARG;
    }
    p.leaf[l].kind = kind;
    p.leaf[l].ell_len = leaf_ell_len[l];
    p.leaf[l].ell = leaf_ell[l];
    p.leaf[l].sigma2 = leaf_sigma2[l];
  }
  p.nterms = n_terms;
  p.X = X; p.n1 = n1; p.ldx = ldx;
  p.symmetric = (X2 == nullptr);
  if (p.symmetric) { p.X2 = X; p.n2 = n1; p.ldx2 = ldx; }
  else { p.X2 = X2; p.n2 = n2; p.ldx2 = ldx2; }
  if (n1 <= 0 || p.n2 <= 0) return GPB_OK;
  if (ldx < D || p.ldx2 < D || ldk < p.n2) return GPB_ERR_BADARG;
  if (reinterpret_cast<uintptr_t>(K) & 15) return GPB_ERR_ALIGN;
  p.D = D; p.noise = noise;
  p.lower = (fill == 1);
  if (p.lower && !p.symmetric) return GPB_ERR_BADARG;
  p.K = K; p.ldk = ldk;
  const int tiles_m = (n1 + SP_TM - 1) / SP_TM;
  p.tiles_n = (p.n2 + SP_TN - 1) / SP_TN;
  int ntiles = tiles_m * p.tiles_n;
  if (p.lower) {
    const int q = tiles_m / 2;
    ntiles = q * (q + 1) + ((tiles_m & 1) ? (q + 1) : 0);
  }
  kern_sop_fwd_kernel<<<ntiles, SP_THREADS, 0, stream>>>(p);
  count_launch();
  GPB_CUDA_CHECK(cudaGetLastError());
  return GPB_OK;
}

}  // namespace gpb
