// Diagonal-block kernel of the blocked Cholesky: factor one 128x128 tile and invert the factor,
// entirely inside one CTA.  This kernel sits on the serial critical path of the factorisation
// (one launch per 128-wide block column), so it is organised for latency:
//   * the tile lives in shared memory (column-major, stride 132 => conflict-free DMMA fragments);
//   * it is factored as 4 block columns of width 32:
//       - the 32x32 diagonal block by ONE warp in registers (lane = row), right-looking, with the
//         pivot broadcast and the rank-1 update done by warp shuffles;  1/L_jj comes from rsqrt so
//         the per-column dependency chain is shuffle -> rsqrt -> multiply -> fma;
//       - the sub-diagonal 32x32 blocks by forward substitution (lane = row, L_kk broadcast from smem);
//       - the trailing 32x32 blocks by DMMA 8x8x4 tiles spread over the 16 warps;
//   * D = L^-1 is then built by recursive doubling: four 32x32 inverses (one warp each, lane =
//     column), then X21 = -X22 (L21 X11) at block sizes 32 and 64 with DMMA tiles.
// Replaces (for one diagonal block) LAPACK dpotrf / dtrtri behind Matrix::chol / chol2inv
// (reference R/GPfunctions.R:826-829).
#pragma once
#include "dgemm_tile.cuh"

namespace gpedm {

constexpr int DG_THREADS = 512;
#ifdef GPEDM_DIAG_TIMING
__device__ long long g_diag_ts[32];
#define DG_TS(i) do { if (threadIdx.x == 0) g_diag_ts[i] = clock64(); } while (0)
#else
#define DG_TS(i) do { } while (0)
#endif
constexpr int DG_LD = 132;   // stride of the 128x128 tile in smem  (== 4 mod 16)
constexpr int DG_WLD = 68;   // stride of the 64x64 scratch          (== 4 mod 16)
constexpr int DG_SMEM_BYTES = (TB * DG_LD + 64 * DG_WLD + 2 * TB) * (int)sizeof(double);

// One 8x8 output tile: c += (+/-) A[8 x K] * op(B).  Abase -> A[row0][k0] (column-major, lda).
// B_TRANS: element (k,n) = Bm[n][k], Bbase -> Bm[col0][k0];  else element (k,n) = Bm[k][n],
// Bbase -> Bm[k0][col0].
template <bool B_TRANS>
__device__ __forceinline__ void tile_mma(double& c0, double& c1, const double* Abase, int lda, const double* Bbase,
                                         int ldb, int K, bool negate, int g, int t) {
  for (int k = 0; k < K; k += 4) {
    double a = Abase[g + (k + t) * lda];
    double b = B_TRANS ? Bbase[g + (k + t) * ldb] : Bbase[(k + t) + g * ldb];
    dmma884(c0, c1, negate ? -a : a, b);
  }
}

__global__ void __launch_bounds__(DG_THREADS, 1)
    diag_block_kernel(double* __restrict__ Akk, size_t lda, double* __restrict__ Dkk, size_t ldd, int kblock,
                      int n_valid, double* __restrict__ logdet_part, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* s = sm;                   // [128][132]
  double* w = sm + TB * DG_LD;      // [64][68] scratch
  double* rdiag = w + 64 * DG_WLD;  // 1 / L_jj
  double* dvec = rdiag + TB;        // L_jj, later reduction scratch
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const unsigned FULL = 0xffffffffu;

  DG_TS(0);
#pragma unroll 8
  for (int idx = tid; idx < TB * TB; idx += DG_THREADS) {
    int r = idx & 127, c = idx >> 7;
    s[r + c * DG_LD] = Akk[(size_t)r + (size_t)c * lda];
  }
  __syncthreads();
  DG_TS(1);

  // ---------------- Cholesky, 4 block columns of width 32 ----------------
#pragma unroll 1
  for (int kb = 0; kb < 4; ++kb) {
    const int base = kb * 32;
    if (warp == 0) {
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = s[(base + lane) + (base + c) * DG_LD];
      // software-pipelined: the next pivot (broadcast + rsqrt) is started right after the first
      // update of column j so its latency hides under the remaining rank-1 update instructions.
      double ajj = __shfl_sync(FULL, a[0], 0);
      double rs = rsqrt(ajj);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        double l = a[j] * rs;
        a[j] = l;
        if (lane == j) {
          if (!(ajj > 0.0) && (kblock * TB + base + j < n_valid)) atomicCAS(info, 0, kblock * TB + base + j + 1);
          rdiag[base + j] = rs;
          dvec[base + j] = l;
        }
        double ajj_n = 1.0, rs_n = 1.0;
        if (j + 1 < 32) {
          double lc1 = __shfl_sync(FULL, l, j + 1);
          a[j + 1] = fma(-l, lc1, a[j + 1]);
          ajj_n = __shfl_sync(FULL, a[j + 1], j + 1);
          rs_n = rsqrt(ajj_n);
        }
#pragma unroll
        for (int c = j + 2; c < 32; ++c) {
          double lc = __shfl_sync(FULL, l, c);
          a[c] = fma(-l, lc, a[c]);
        }
        ajj = ajj_n;
        rs = rs_n;
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) s[(base + lane) + (base + c) * DG_LD] = (c <= lane) ? a[c] : 0.0;
    }
    __syncthreads();
    DG_TS(2 + 3 * kb);
    // panel: rows below the diagonal block, forward substitution per row
    if (warp < 3 - kb) {
      const int row = base + 32 * (warp + 1) + lane;
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = s[row + (base + c) * DG_LD];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        double x = a[c] * rdiag[base + c];
        a[c] = x;
#pragma unroll
        for (int c2 = c + 1; c2 < 32; ++c2) a[c2] = fma(-x, s[(base + c2) + (base + c) * DG_LD], a[c2]);
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) s[row + (base + c) * DG_LD] = a[c];
    }
    __syncthreads();
    DG_TS(3 + 3 * kb);
    // trailing update on the remaining lower 32x32 blocks (16 DMMA tiles each)
    const int m = 3 - kb;
    const int ntile = m * (m + 1) / 2 * 16;
    for (int q = warp; q < ntile; q += DG_THREADS / 32) {
      int blk = q >> 4, tile = q & 15;
      int bi = blk < 1 ? 0 : (blk < 3 ? 1 : 2);
      int bj = blk - bi * (bi + 1) / 2;
      int row0 = base + 32 * (bi + 1) + 8 * (tile >> 2);
      int col0 = base + 32 * (bj + 1) + 8 * (tile & 3);
      double* cp = s + (row0 + g) + (col0 + 2 * t) * DG_LD;
      double c0 = cp[0], c1 = cp[DG_LD];
      tile_mma<true>(c0, c1, s + row0 + base * DG_LD, DG_LD, s + col0 + base * DG_LD, DG_LD, 32, true, g, t);
      cp[0] = c0;
      cp[DG_LD] = c1;
    }
    __syncthreads();
    DG_TS(4 + 3 * kb);
  }

  // ---------------- write L, log-determinant part ----------------
  for (int idx = tid; idx < TB * TB; idx += DG_THREADS) {
    int r = idx & 127, c = idx >> 7;
    Akk[(size_t)r + (size_t)c * lda] = (r >= c) ? s[r + c * DG_LD] : 0.0;
  }
  {
    double v = 0.0;
    if (tid < TB && kblock * TB + tid < n_valid) v = log(dvec[tid]);
    __syncthreads();
    if (tid < TB) dvec[tid] = v;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
      if (tid < off) dvec[tid] += dvec[tid + off];
      __syncthreads();
    }
    if (tid == 0) logdet_part[kblock] = dvec[0];
  }
  DG_TS(14);

  // ---------------- inverse: 32x32 diagonal blocks (warp b, lane = column) ----------------
  if (warp < 4) {
    const int base = warp * 32;
    double x[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int mm = 0; mm < r; ++mm) {
        double lv = s[(base + r) + (base + mm) * DG_LD];
        if ((mm & 3) == 0) s0 = fma(lv, x[mm], s0);
        else if ((mm & 3) == 1) s1 = fma(lv, x[mm], s1);
        else if ((mm & 3) == 2) s2 = fma(lv, x[mm], s2);
        else s3 = fma(lv, x[mm], s3);
      }
      s0 = (s0 + s1) + (s2 + s3);
      s1 = 0.0;
      double rd = rdiag[base + r];
      double v = -(s0 + s1) * rd;
      x[r] = (r == lane) ? rd : ((r < lane) ? 0.0 : v);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) s[(base + r) + (base + lane) * DG_LD] = x[r];
  }
  __syncthreads();
  DG_TS(15);
  // ---------------- level 1: 32-blocks, pairs (0,1) and (2,3) ----------------
  for (int q = warp; q < 32; q += DG_THREADS / 32) {  // W = L21 * X11
    int p = q >> 4, tile = q & 15;
    int o = p * 64;
    int r0 = 8 * (tile >> 2), c0i = 8 * (tile & 3);
    double c0 = 0.0, c1 = 0.0;
    tile_mma<false>(c0, c1, s + (o + 32 + r0) + o * DG_LD, DG_LD, s + o + (o + c0i) * DG_LD, DG_LD, 32, false, g, t);
    double* wp = w + (p * 32 + r0 + g) + (c0i + 2 * t) * DG_WLD;
    wp[0] = c0;
    wp[DG_WLD] = c1;
  }
  __syncthreads();
  for (int q = warp; q < 32; q += DG_THREADS / 32) {  // X21 = -X22 * W
    int p = q >> 4, tile = q & 15;
    int o = p * 64;
    int r0 = 8 * (tile >> 2), c0i = 8 * (tile & 3);
    double c0 = 0.0, c1 = 0.0;
    tile_mma<false>(c0, c1, s + (o + 32 + r0) + (o + 32) * DG_LD, DG_LD, w + p * 32 + c0i * DG_WLD, DG_WLD, 32, true, g,
                    t);
    double* xp = s + (o + 32 + r0 + g) + (o + c0i + 2 * t) * DG_LD;
    xp[0] = c0;
    xp[DG_LD] = c1;
  }
  __syncthreads();
  DG_TS(16);
  // ---------------- level 2: 64-blocks ----------------
  for (int q = warp; q < 64; q += DG_THREADS / 32) {  // W = L21 * X11, k >= 32*colblock
    int r0 = 8 * (q >> 3), c0i = 8 * (q & 7);
    int kbeg = (c0i >= 32) ? 32 : 0;
    double c0 = 0.0, c1 = 0.0;
    tile_mma<false>(c0, c1, s + (64 + r0) + kbeg * DG_LD, DG_LD, s + kbeg + c0i * DG_LD, DG_LD, 64 - kbeg, false, g, t);
    double* wp = w + (r0 + g) + (c0i + 2 * t) * DG_WLD;
    wp[0] = c0;
    wp[DG_WLD] = c1;
  }
  __syncthreads();
  for (int q = warp; q < 64; q += DG_THREADS / 32) {  // X21 = -X22 * W, k < 32*(rowblock+1)
    int r0 = 8 * (q >> 3), c0i = 8 * (q & 7);
    int kend = (r0 >= 32) ? 64 : 32;
    double c0 = 0.0, c1 = 0.0;
    tile_mma<false>(c0, c1, s + (64 + r0) + 64 * DG_LD, DG_LD, w + c0i * DG_WLD, DG_WLD, kend, true, g, t);
    double* xp = s + (64 + r0 + g) + (c0i + 2 * t) * DG_LD;
    xp[0] = c0;
    xp[DG_LD] = c1;
  }
  __syncthreads();
  DG_TS(17);
  for (int idx = tid; idx < TB * TB; idx += DG_THREADS) {
    int r = idx & 127, c = idx >> 7;
    Dkk[(size_t)r + (size_t)c * ldd] = (r >= c) ? s[r + c * DG_LD] : 0.0;
  }
  DG_TS(18);
}

}  // namespace gpedm
