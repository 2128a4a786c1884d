// gpb_gemm.cu -- FP64 GEMM / SYRK engine for sm_100a: DMMA.8x8x4 warp tiles fed by a TMA + mbarrier ring.
//
// This is the kernel every O(N^3) step of the hot path runs on: the Cholesky trailing updates and panel
// solves (reference: torch.cholesky behind gptorch/functions.py:46-47), the triangular inverse and
// L^-T L^-1 product that replace autograd's CholeskyBackward0, and the
// Kuf-panel products of the sparse models (gptorch/models/sparse_gpr.py:132-137).
//
// Design (B200): CTA tile 128x128 (64x64 for launches too small to fill the chip), K-chunk 16 (one 128-byte
// swizzled TMA row per operand row), ring of 5-6 stages, 8 consumer warps (2 x 4, warp tile 64 x 32 = 8 x 4
// DMMA fragments, 64 fp64 accumulators per thread) + 1 producer warp whose elected lane issues
// cp.async.bulk.tensor.  FP64 peak
// on B200 is 64 FMA/clk/SM for DFMA and DMMA alike (measured 37.0 TFLOP/s); DMMA needs 8x fewer issue
// slots and 12 LDS.64 per 32 DMMA, so shared-memory bandwidth and issue are far from limiting and the
// k-permuted fragment addressing below is bank-conflict free under SWIZZLE_128B.
#include "gpb_gemm.cuh"
#include <cstdlib>

namespace gpb {

constexpr int BK = 16;                 // k-chunk: one 128-byte (swizzled) row per operand row
constexpr int NT_BOX_ROWS = 32;        // rows of a K-contiguous TMA box (all NT/NN-A tensor maps use this)

// Tile configurations.  L: the throughput shape (1 CTA/SM, 64 accumulators per thread).  S: for launches that
// would leave most SMs idle with 128x128 tiles (the latency-bound small products at the bottom of the
// Cholesky / TRSM recursions) -- 4x the CTAs, each 1/4 of the work.
template <int BM_, int BN_, int WM_, int WN_, int STAGES_>
struct GemmCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_;
  static constexpr int MI = WM / 8, NI = WN / 8;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N;
  static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
  static constexpr int A_TILE_BYTES = BM * BK * 8;
  static constexpr int B_TILE_BYTES = BN * BK * 8;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};
using CfgL = GemmCfg<128, 128, 64, 32, 5>;
// H: the 128 x 128 tile split into a left and a right 128 x 64 half, one CTA each (4 consumer warps with the same 64 x 32
// warp tiles, 4 stages of 24 KB): TWO CTAs per SM, so that the epilogue (C read-modify-write), the launch gap and the
// pipeline fill of one overlap the main loop of the other instead of idling the FP64 pipe.  Tiles are enumerated in 128 x 128
// units exactly as for L (same super-block order); blockIdx.x & 1 selects the half.
using CfgH = GemmCfg<128, 64, 64, 32, 4>;
using CfgS = GemmCfg<64, 64, 32, 32, 6>;
// T: row strips for the in-place right-TRSM base case (N <= 128 = BN: one CTA owns all columns of its rows, so
// every TMA read of those rows completes before the CTA's own epilogue overwrites them).
using CfgT = GemmCfg<32, 128, 32, 32, 6>;

struct GemmKParams {
  int M, N, K;
  double alpha, beta;
  double* C;
  long ldc, c_batch;
  double* Cdiag;
  long ldd;
  int ax, ay, bx, by, dax, day, dbx, dby;
  unsigned flags;
  int tiles_m, tiles_n;
  unsigned zero;   // always 0 at run time; opaque to the compiler (see the stage-release dependency in the main loop)
  long long stagger;            // CfgH: start delay (clocks) of the CTAs with stagger_lo <= blockIdx.x < stagger_hi
  unsigned stagger_lo, stagger_hi;
};

// Swizzled byte offset inside a "row-tile" (rows of 16 doubles = 128 B, SWIZZLE_128B): element (row, col).
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t col) {
  return row * 128u + ((((col >> 1) ^ (row & 7u)) << 4) | ((col & 1u) << 3));
}

template <int MODE, class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, (Cfg::BM == 128 && Cfg::BN == 64) ? 2 : 1)
gemm_dmma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const GemmKParams p) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, WM = Cfg::WM, WN = Cfg::WN, MI = Cfg::MI, NI = Cfg::NI;
  constexpr int STAGES = Cfg::STAGES, STAGE_BYTES = Cfg::STAGE_BYTES, A_TILE_BYTES = Cfg::A_TILE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_al + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile coordinates -------------------------------------------------------------------------
  // Tiles are visited in super-blocks of SUPER x SUPER tiles (about one wave of 148 CTAs): the CTAs that are
  // resident together then share SUPER row panels and SUPER column panels, which stay in L2 instead of being
  // re-fetched from HBM for every tile.  Super-block rows ascend, so LAUUM's heaviest tiles still go first.
  constexpr bool HALF = (BM == 128 && BN == 64);
  int tm, tn;
  {
    constexpr int SUPER = 12;
    int t = HALF ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    if (p.flags & GF_LOWER_TILES) {
      const int T = p.tiles_m;
      int sm = 0, r0 = 0, h = min(SUPER, T);
      for (;;) {
        const int cnt = h * SUPER * sm + h * (h + 1) / 2;
        if (t < cnt) break;
        t -= cnt;
        ++sm;
        r0 += SUPER;
        h = min(SUPER, T - r0);
      }
      const int off_diag = h * SUPER * sm;
      if (t < off_diag) {
        const int sn = t / (h * SUPER), q = t - sn * (h * SUPER);
        tm = r0 + q / SUPER;
        tn = sn * SUPER + q % SUPER;
      } else {
        const int q = t - off_diag;
        int i = static_cast<int>((sqrtf(8.0f * q + 1.0f) - 1.0f) * 0.5f);
        while ((i + 1) * (i + 2) / 2 <= q) ++i;
        while (i * (i + 1) / 2 > q) --i;
        tm = r0 + i;
        tn = r0 + (q - i * (i + 1) / 2);
      }
    } else {
      const int Tn = p.tiles_n;
      const int rows_per_super = SUPER * Tn;            // tiles in one full super row
      const int sm = t / rows_per_super;
      const int r0 = sm * SUPER;
      const int h = min(SUPER, p.tiles_m - r0);
      t -= sm * rows_per_super;
      const int per_block = h * SUPER;                    // tiles in one full super-block of this super row
      const int sn = t / per_block;
      const int c0 = sn * SUPER;
      const int w = min(SUPER, Tn - c0);
      const int q = t - sn * per_block;
      tm = r0 + q / w;
      tn = c0 + q % w;
    }
  }
  const int m0 = tm * BM, n0 = HALF ? tn * 128 + static_cast<int>(blockIdx.x & 1) * 64 : tn * BN;
  const int bz = blockIdx.y;
  if (HALF && n0 >= p.N) return;     // right half of a ragged last tile column
  if (HALF && p.stagger > 0 && blockIdx.x >= p.stagger_lo && blockIdx.x < p.stagger_hi) {
    // first wave only: the CTAs that land in the second slot of each SM start half a tile late, so that the two
    // co-resident CTAs of an SM alternate between main loop and epilogue instead of idling the pipe together
    const long long t0 = clock64();
    while (clock64() - t0 < p.stagger) { __nanosleep(200); }
  }

  int k_lo = 0, k_hi = p.K;
  if (p.flags & GF_KLO_M) k_lo = max(k_lo, m0);
  if (p.flags & GF_KLO_N) k_lo = max(k_lo, n0);
  if (p.flags & GF_KHI_M) k_hi = min(k_hi, m0 + BM);
  if (p.flags & GF_KHI_N) k_hi = min(k_hi, n0 + BN);
  k_lo &= ~(BK - 1);
  const int nk = k_hi > k_lo ? (k_hi - k_lo + BK - 1) / BK : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], Cfg::CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == Cfg::CONSUMER_WARPS) {
    // =============================== TMA producer ==============================================
    // The whole warp walks the k-chunks in lock step; lane 0 arms the barrier and issues the TMA loads.  During
    // the last PREFETCH_AHEAD chunks all 32 lanes also pull the C tile into L2 (beta != 0), a few 64-byte
    // segments per chunk, so the epilogue's read-modify-write of C does not pay DRAM latency.
    constexpr int PREFETCH_AHEAD = 24;
    constexpr int SEG = BN / 8;                 // 64-byte segments per tile row
    constexpr int TOTAL_SEG = BM * SEG;
    const int pit = (p.beta != 0.0) ? max(nk - PREFETCH_AHEAD, 0) : nk;
    const int seg_per_it = nk > pit ? (TOTAL_SEG + (nk - pit) - 1) / (nk - pit) : 0;
    const bool diag_ws = (p.flags & GF_DIAG_TO_WS) && (m0 / NB) == (n0 / NB);
    const double* Cp = diag_ws ? p.Cdiag : p.C + static_cast<long>(bz) * p.c_batch;
    const long ldp = diag_ws ? p.ldd : p.ldc;
    const int shift = diag_ws ? (n0 / NB) * NB : 0;
    if (lane == 0) {
      tma_prefetch_desc(&mapA);
      tma_prefetch_desc(&mapB);
    }
    const int ax = p.ax + bz * p.dax, ay = p.ay + bz * p.day;
    const int bx = p.bx + bz * p.dbx, by = p.by + bz * p.dby;
    for (int it = 0; it < nk; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      if (lane == 0) {
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        uint8_t* a_dst = smem_al + s * STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_TILE_BYTES;
        const int k = k_lo + it * BK;
        if (MODE == GEMM_TN) {
#pragma unroll
          for (int j = 0; j < BM / 16; ++j) tma_load_2d(a_dst + j * 2048, &mapA, ax + m0 + 16 * j, ay + k, &full_bar[s]);
        } else {
#pragma unroll
          for (int j = 0; j < BM / NT_BOX_ROWS; ++j)
            tma_load_2d(a_dst + j * (NT_BOX_ROWS * 128), &mapA, ax + k, ay + m0 + NT_BOX_ROWS * j, &full_bar[s]);
        }
        if (MODE == GEMM_NT) {
#pragma unroll
          for (int j = 0; j < BN / NT_BOX_ROWS; ++j)
            tma_load_2d(b_dst + j * (NT_BOX_ROWS * 128), &mapB, bx + k, by + n0 + NT_BOX_ROWS * j, &full_bar[s]);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 16; ++j) tma_load_2d(b_dst + j * 2048, &mapB, bx + n0 + 16 * j, by + k, &full_bar[s]);
        }
      }
      if (it >= pit) {
        const int first = (it - pit) * seg_per_it;
        for (int q = lane; q < seg_per_it; q += 32) {
          const int idx = first + q;
          const int row = m0 + idx / SEG, col = n0 + (idx % SEG) * 8;
          if (idx < TOTAL_SEG && row < p.M && col < p.N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Cp + static_cast<long>(row) * ldp - shift + col));
        }
      }
      __syncwarp();
    }
    return;
  }

  // ================================= DMMA consumers ==============================================
  const int wm = warp % Cfg::WARPS_M, wn = warp / Cfg::WARPS_M;
  const int r = lane >> 2, kk = lane & 3;

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // Per-lane fragment offsets (bytes, relative to the A / B tile of a stage).  For k4-step s (0..3):
  //  K-contiguous tile [row][16 k]: element (row, kcol), kcol = 2s + (kk&1) + 8(kk>>1): chunk = s + 4(kk>>1), so
  //     off(s) = off(0) ^ (s << 4)                       (the XOR swizzle makes the step a bit flip)
  //  MN-contiguous tile [16 k][16 x] per sub-box: element (krow, x) with
  //     TN (both operands): krow = 2kk + (s&1) + 8(s>>1):  off(s) = (off(0) ^ ((s&1) << 4)) + (s&1)*128 + (s>>1)*1024
  //     NN (B operand):     krow = kcol above:             off(s) = (off(0) ^ ((s&1) << 5)) + (s&1)*256 ... (see kn_off)
  // Only the s = 0 offsets live in registers; the per-step part is an immediate.
  uint32_t a_off0, b_off0;
  {
    const uint32_t kcol0 = (kk & 1) + 8 * (kk >> 1);
    const uint32_t krow0 = 2 * kk;
    if (MODE == GEMM_TN) {
      const uint32_t ml = wm * WM + r;  // + 8i added below (i even/odd changes the sub-box column half)
      a_off0 = (ml >> 4) * 2048u + swz(krow0, ml & 15u);
      const uint32_t nl = wn * WN + r;
      b_off0 = (nl >> 4) * 2048u + swz(krow0, nl & 15u);
    } else {
      a_off0 = swz(wm * WM + r, kcol0);
      if (MODE == GEMM_NT) {
        b_off0 = swz(wn * WN + r, kcol0);
      } else {
        const uint32_t nl = wn * WN + r;
        b_off0 = (nl >> 4) * 2048u + swz(kcol0, nl & 15u);
      }
    }
  }
  // step offsets as functions of the compile-time step index
  auto kc_off = [](uint32_t off0, int s) -> uint32_t { return off0 ^ (static_cast<uint32_t>(s) << 4); };
  auto tn_off = [](uint32_t off0, int s) -> uint32_t {
    return (off0 ^ (static_cast<uint32_t>(s & 1) << 4)) + static_cast<uint32_t>(s & 1) * 128u + static_cast<uint32_t>(s >> 1) * 1024u;
  };
  // NN B operand: krow = 2s + (kk&1) + 8(kk>>1): row offset +256 s bytes; (krow & 7) = (2s + (kk&1)) & 7 flips chunk
  // bits 1-2 by s (2s < 8): chunk ^= 2s  -> byte offset bits 5-6 ^= s.
  auto kn_off = [](uint32_t off0, int s) -> uint32_t {
    return (off0 ^ (static_cast<uint32_t>(s) << 5)) + static_cast<uint32_t>(s) * 256u;
  };

  for (int it = 0; it < nk; ++it) {
    const int s = it % STAGES;
    const uint32_t ph = (it / STAGES) & 1;
    mbar_wait(&full_bar[s], ph);
    const uint32_t a_base = smem_base + s * STAGE_BYTES;
    const uint32_t b_base = a_base + A_TILE_BYTES;
    uint32_t dep = 0;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      double a[MI], b[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        uint32_t off;
        if (MODE == GEMM_TN) {
          // rows 8i of the warp tile: i odd -> columns 8..15 of the sub-box (chunk index + 4), i>>1 -> next sub-box
          off = tn_off(a_off0, ks) + (i >> 1) * 2048u;
          if (i & 1) off ^= 64u;  // (x>>1) + 4 under the XOR swizzle == flip bit 6 of the byte offset
        } else {
          off = kc_off(a_off0, ks) + i * 1024u;  // 8 rows x 128 B; row & 7 unchanged
        }
        a[i] = ld_shared_f64(a_base + off);
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        uint32_t off;
        if (MODE == GEMM_NT) {
          off = kc_off(b_off0, ks) + j * 1024u;
        } else {
          off = (MODE == GEMM_TN ? tn_off(b_off0, ks) : kn_off(b_off0, ks)) + (j >> 1) * 2048u;
          if (j & 1) off ^= 64u;
        }
        b[j] = ld_shared_f64(b_base + off);
      }
      if (ks == 3) {
        // Every fragment register of the last k-step feeds `dep`: the stage-release below cannot issue before
        // these (in-order) shared-memory loads have returned.
#pragma unroll
        for (int i = 0; i < MI; ++i) dep |= static_cast<uint32_t>(__double2hiint(a[i]));
#pragma unroll
        for (int j = 0; j < NI; ++j) dep |= static_cast<uint32_t>(__double2hiint(b[j]));
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    // Release the stage to the TMA producer.  The arrive must not overtake any LDS of this stage (the compiler
    // and ptxas are free to sink the MMAs below it), so its address carries a data dependency on the loaded
    // fragments of ALL lanes: warp-wide OR, masked with a run-time zero.
    dep = __reduce_or_sync(0xffffffffu, dep) & p.zero;
    if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(&empty_bar[s]) + dep));
  }

  // ---------------------------------- epilogue ---------------------------------------------------
  double* Cb = p.C + static_cast<long>(bz) * p.c_batch;
  long ldc = p.ldc;
  int col_shift = 0;
  if ((p.flags & GF_DIAG_TO_WS) && (m0 / NB) == (n0 / NB)) {
    Cb = p.Cdiag;
    ldc = p.ldd;
    col_shift = (n0 / NB) * NB;
  }
  const double alpha = p.alpha, beta = p.beta;
  const int col_base = n0 + wn * WN + 2 * kk;
  const bool full_cols = (n0 + wn * WN + WN) <= p.N;   // every column pair of this warp tile is in range
  if (full_cols) {
    // Fast path.  The C tile was prefetched into L2 by the producer warp's spare lanes; it is read in batches
    // of EB row groups (EB x NI double2 loads in flight per thread) before anything is stored, so the memory
    // latency is paid once per batch instead of once per element pair.
    constexpr int EB = (MI * NI > 16) ? 1 : 2;   // row groups per batch (register budget: 168 with 9 warps)
#pragma unroll
    for (int i0 = 0; i0 < MI; i0 += EB) {
      double cx[EB][NI], cy[EB][NI];
#pragma unroll
      for (int ii = 0; ii < EB; ++ii) {
        const int row = m0 + wm * WM + 8 * (i0 + ii) + r;
        const double* crow = Cb + static_cast<long>(row) * ldc - col_shift + col_base;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          cx[ii][j] = 0.0;
          cy[ii][j] = 0.0;
          if (beta != 0.0 && row < p.M) {
            const double2 t = *reinterpret_cast<const double2*>(crow + 8 * j);
            cx[ii][j] = t.x;
            cy[ii][j] = t.y;
          }
        }
      }
#pragma unroll
      for (int ii = 0; ii < EB; ++ii) {
        const int row = m0 + wm * WM + 8 * (i0 + ii) + r;
        double* crow = Cb + static_cast<long>(row) * ldc - col_shift + col_base;
        if (row < p.M) {
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            double2 v;
            v.x = alpha * acc[i0 + ii][j][0] + beta * cx[ii][j];
            v.y = alpha * acc[i0 + ii][j][1] + beta * cy[ii][j];
            *reinterpret_cast<double2*>(crow + 8 * j) = v;
          }
        }
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = m0 + wm * WM + 8 * i + r;
    if (row >= p.M) continue;
    double* crow = Cb + static_cast<long>(row) * ldc - col_shift;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int col = n0 + wn * WN + 8 * j + 2 * kk;
      if (col + 1 < p.N) {
        double2* ptr = reinterpret_cast<double2*>(crow + col);
        double2 v;
        if (beta != 0.0) {
          v = *ptr;
          v.x = alpha * acc[i][j][0] + beta * v.x;
          v.y = alpha * acc[i][j][1] + beta * v.y;
        } else {
          v.x = alpha * acc[i][j][0];
          v.y = alpha * acc[i][j][1];
        }
        *ptr = v;
      } else if (col < p.N) {
        double v = alpha * acc[i][j][0];
        if (beta != 0.0) v += beta * crow[col];
        crow[col] = v;
      }
    }
  }
}

template <int MODE, class Cfg>
static int launch_cfg(const CUtensorMap& mapA, const CUtensorMap& mapB, GemmKParams kp, const GemmArgs& a,
                      cudaStream_t stream) {
  static std::atomic<int> smem_state[GPB_MAX_DEVICES];
  if (int rc = ensure_dynamic_smem(gemm_dmma_kernel<MODE, Cfg>, Cfg::SMEM_BYTES, smem_state)) return rc;
  constexpr bool HALF = (Cfg::BM == 128 && Cfg::BN == 64);
  kp.tiles_m = (a.M + Cfg::BM - 1) / Cfg::BM;
  kp.tiles_n = HALF ? (a.N + 127) / 128 : (a.N + Cfg::BN - 1) / Cfg::BN;
  int ntiles;
  if (a.flags & GF_LOWER_TILES) {
    if (kp.tiles_m != kp.tiles_n) return GPB_ERR_BADARG;
    ntiles = kp.tiles_m * (kp.tiles_m + 1) / 2;
  } else {
    ntiles = kp.tiles_m * kp.tiles_n;
  }
  if (HALF) {
    ntiles *= 2;
    static std::atomic<int> carve_state[GPB_MAX_DEVICES];
    {
      int d0 = 0;
      cudaGetDevice(&d0);
      if (d0 >= 0 && d0 < GPB_MAX_DEVICES && carve_state[d0].load(std::memory_order_relaxed) == 0) {
        // two CTAs of ~100 KB each: ask for the largest shared-memory carve-out
        cudaFuncSetAttribute(gemm_dmma_kernel<MODE, Cfg>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        carve_state[d0].store(1, std::memory_order_relaxed);
      }
    }
    static const long stagger_env = []() { const char* v = getenv("GPB_GEMM_STAGGER"); return v && *v ? atol(v) : -1L; }();
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static int sm_count[GPB_MAX_DEVICES] = {0};
    if (dev >= 0 && dev < GPB_MAX_DEVICES) {
      if (sm_count[dev] == 0) cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
      sms = sm_count[dev];
    }
    // half a tile of FP64-pipe time when two CTAs share the pipe: 128 x 64 x K FMA at 64 FMA/clk/SM
    const long long half_tile = static_cast<long long>(a.K) * 128 * 64 / 64 / 2;
    // measured: the two CTAs of an SM fall out of phase on their own (stagger 0 = auto = 35.9 TFLOP/s); the explicit
    // first-wave delay stays available through GPB_GEMM_STAGGER=<clocks> for experiments
    (void)half_tile;
    kp.stagger = stagger_env > 0 ? stagger_env : 0;
    kp.stagger_lo = static_cast<unsigned>(sms);
    kp.stagger_hi = static_cast<unsigned>(2 * sms);
  }
  dim3 grid(ntiles, a.batch, 1);
  gemm_dmma_kernel<MODE, Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(mapA, mapB, kp);
  count_launch();
  GPB_CUDA_CHECK(cudaGetLastError());
  return GPB_OK;
}

template <int MODE>
static int launch_mode(const CUtensorMap& mapA, const CUtensorMap& mapB, const GemmKParams& kp, const GemmArgs& a,
                       cudaStream_t stream) {
  // 128x128 tiles unless that would occupy fewer than ~half of the 148 SMs.
  const long t128 = static_cast<long>((a.M + 127) / 128) * ((a.N + 127) / 128) * a.batch;
  const long tiles = (a.flags & GF_LOWER_TILES) ? (t128 + a.batch * ((a.M + 127) / 128)) / 2 : t128;
  static const long small_env = []() { const char* v = getenv("GPB_GEMM_SMALL_TILES"); return v && *v ? atol(v) : 0L; }();
  // Lower-tile (SYRK) launches of up to ~2 waves of 128x128 tiles fill the 148 SMs better with 64x64 tiles (measured:
  // m = 1536: 0.234 -> 0.193 ms, m = 3072: 1.27 -> 0.88 ms); full launches only below half a wave (NT/TN at 2048^3 are
  // 1-2 % slower with small tiles).  GPB_GEMM_SMALL_TILES overrides both (tuning aid, read once).
  const long small = small_env > 0 ? small_env
                                   : ((a.flags & GF_LOWER_TILES) ? GEMM_SMALL_TILE_THRESHOLD_LOWER : GEMM_SMALL_TILE_THRESHOLD);
  if (a.flags & GF_ROWS_INPLACE) {
    if (a.N > 128) return GPB_ERR_BADARG;
    if (tiles <= GEMM_SMALL_TILE_THRESHOLD) return launch_cfg<MODE, CfgT>(mapA, mapB, kp, a, stream);
    return launch_cfg<MODE, CfgL>(mapA, mapB, kp, a, stream);
  }
  // batched lower-tile launches (the split-K Gram products of the sparse models) are judged per problem: with M = 1024 a
  // slice has 36 tiles of 128 x 128, 8 of them half-empty diagonal tiles -- 136 tiles of 64 x 64 waste a third of that
  // (VFE forward statistics 52.2 -> 50.0 ms per 1.25e6-row shard)
  const long tiles_rule = (a.flags & GF_LOWER_TILES) && a.batch > 1 ? tiles / a.batch : tiles;
  if (tiles_rule <= small) return launch_cfg<MODE, CfgS>(mapA, mapB, kp, a, stream);
  // From about two waves of 128 x 128 tiles on, the half-tile configuration (two CTAs per SM) wins: measured SYRK
  // m = 28672, k = 2048: 35.19 -> 35.88 TFLOP/s, m = 8192: 33.15 -> 34.54; NN 131072 x 1024 x 1024: 34.94 -> 35.69
  // (tools/bench_gemm_half.py).  GPB_GEMM_HALF=<min tiles> overrides the threshold, a huge value disables it (tuning aid).
  static const long half_env = []() { const char* v = getenv("GPB_GEMM_HALF"); return v && *v ? atol(v) : 0L; }();
  const long half_min = half_env > 0 ? half_env : GEMM_HALF_TILE_THRESHOLD;
  if (tiles >= half_min) return launch_cfg<MODE, CfgH>(mapA, mapB, kp, a, stream);
  return launch_cfg<MODE, CfgL>(mapA, mapB, kp, a, stream);
}

int gemm_launch(GemmMode mode, const CUtensorMap& mapA, const CUtensorMap& mapB, const GemmArgs& a,
                cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return GPB_OK;
  if (a.K < 0 || a.C == nullptr || (a.ldc & 1) || (reinterpret_cast<uintptr_t>(a.C) & 15)) return GPB_ERR_ALIGN;
  if (a.c_batch & 1) return GPB_ERR_ALIGN;
  GemmKParams kp;
  kp.M = a.M; kp.N = a.N; kp.K = a.K;
  kp.alpha = a.alpha; kp.beta = a.beta;
  kp.C = a.C; kp.ldc = a.ldc; kp.c_batch = a.c_batch;
  kp.Cdiag = a.Cdiag; kp.ldd = a.ldd;
  kp.ax = a.ax; kp.ay = a.ay; kp.bx = a.bx; kp.by = a.by;
  kp.dax = a.dax; kp.day = a.day; kp.dbx = a.dbx; kp.dby = a.dby;
  kp.flags = a.flags;
  kp.tiles_m = kp.tiles_n = 0;
  kp.zero = 0u;
  kp.stagger = 0; kp.stagger_lo = kp.stagger_hi = 0u;
  if ((a.flags & GF_DIAG_TO_WS) && (a.Cdiag == nullptr || (a.ldd & 1))) return GPB_ERR_BADARG;
  switch (mode) {
    case GEMM_NT: return launch_mode<GEMM_NT>(mapA, mapB, kp, a, stream);
    case GEMM_TN: return launch_mode<GEMM_TN>(mapA, mapB, kp, a, stream);
    case GEMM_NN: return launch_mode<GEMM_NN>(mapA, mapB, kp, a, stream);
  }
  return GPB_ERR_BADARG;
}

}  // namespace gpb
