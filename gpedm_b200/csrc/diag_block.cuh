// Diagonal-block kernel of the blocked Cholesky: factor one 128x128 tile and invert the factor,
// entirely inside one CTA.  This kernel sits on the serial critical path of the factorisation
// (one launch per 128-wide block column), so it is organised for latency:
//   * the tile lives in shared memory (column-major, stride 132 => conflict-free DMMA fragments);
//   * it is factored as 4 block columns of width 32:
//       - the 32x32 diagonal block by ONE warp in registers (lane = row), right-looking; the
//         pivot column is broadcast through a double-buffered shared-memory vector read back with
//         LDS.128, 1/L_jj comes from rsqrt, and the next pivot (own-lane fma -> shuffle -> rsqrt)
//         is started before the rank-1 update so its latency hides under the update;
//       - the sub-diagonal 32x32 blocks by forward substitution (lane = row, L_kk broadcast);
//       - the trailing 32x32 blocks by DMMA 8x8x4 tiles; the next diagonal block is updated
//         first so that its factorisation (warp 0) overlaps the rest of the update (warps 1-15);
//   * D = L^-1 is built by recursive doubling X21 = -X22 (L21 X11) from 8x8 diagonal inverses
//     (one thread per column) through block sizes 8, 16, 32, 64, all products on DMMA tiles.
// Replaces (for one diagonal block) LAPACK dpotrf / dtrtri behind Matrix::chol / chol2inv
// (reference R/GPfunctions.R:826-829).
#pragma once
#include "dgemm_tile.cuh"
#include "tile_stream.cuh"

namespace gpedm {

constexpr int DG_THREADS = 512;
#ifdef GPEDM_DIAG_TIMING
__device__ long long g_diag_ts[32];
#define DG_TS(i) do { if (threadIdx.x == 0) g_diag_ts[i] = clock64(); __syncwarp(); } while (0)
#else
#define DG_TS(i) do { } while (0)
#endif
constexpr int DG_LD = 132;   // stride of the 128x128 tile in smem  (== 4 mod 16)
constexpr int DG_WLD = 68;   // stride of the 64x64 scratch          (== 4 mod 16)
constexpr int DG_SMEM_BYTES = (TB * DG_LD + 64 * DG_WLD + 2 * TB + 64) * (int)sizeof(double);

// One 8x8 output tile: c += (+/-) A[8 x K] * op(B), K a multiple of 8.  Abase -> A[row0][k0]
// (column-major, lda).  B_TRANS: element (k,n) = Bm[n][k], Bbase -> Bm[col0][k0]; else element
// (k,n) = Bm[k][n], Bbase -> Bm[k0][col0].  Fragments of the next 8 k are fetched while the
// current two DMMAs issue.
template <bool B_TRANS>
__device__ __forceinline__ void tile_mma(double& c0, double& c1, const double* Abase, int lda, const double* Bbase,
                                         int ldb, int K, unsigned sgn, int g, int t) {
  const double* ap = Abase + g + t * lda;
  const double* bp = B_TRANS ? Bbase + g + t * ldb : Bbase + t + g * ldb;
  const int astep = 4 * lda, bstep = B_TRANS ? 4 * ldb : 4;
  double a0 = ap[0], a1 = ap[astep], b0 = bp[0], b1 = bp[bstep];
  for (int k = 8; k < K; k += 8) {
    ap += 2 * astep;
    bp += 2 * bstep;
    double na0 = ap[0], na1 = ap[astep], nb0 = bp[0], nb1 = bp[bstep];
    dmma884(c0, c1, flip_sign(a0, sgn), b0);
    dmma884(c0, c1, flip_sign(a1, sgn), b1);
    a0 = na0;
    a1 = na1;
    b0 = nb0;
    b1 = nb1;
  }
  dmma884(c0, c1, flip_sign(a0, sgn), b0);
  dmma884(c0, c1, flip_sign(a1, sgn), b1);
}


// RG output tiles at once that share one operand: SHARE_B (tiles stacked along rows: distinct A
// row blocks 8 rows apart, common B) or shared A (tiles side by side: common A, distinct B column
// blocks 8 columns apart).  The RG independent DMMA chains hide each other's latency.
// All operands are plain (non-transposed) here: element (k,n) of B = Bm[k][n].
template <int RG, bool SHARE_B>
__device__ __forceinline__ void tile_mma_multi(double (&c)[RG][2], const double* Abase, int lda, const double* Bbase,
                                               int ldb, int K, unsigned sgn, int g, int t) {
  const double* ap = Abase + g + t * lda;
  const double* bp = Bbase + t + g * ldb;
  for (int k = 0; k < K; k += 4) {
    if (SHARE_B) {
      const double b = bp[0];
#pragma unroll
      for (int r = 0; r < RG; ++r) dmma884(c[r][0], c[r][1], flip_sign(ap[8 * r], sgn), b);
    } else {
      const double a = flip_sign(ap[0], sgn);
#pragma unroll
      for (int r = 0; r < RG; ++r) dmma884(c[r][0], c[r][1], a, bp[8 * r * ldb]);
    }
    ap += 4 * lda;
    bp += 4;
  }
}

// One level of the recursive doubling (block size sz, RG = tiles per warp).
template <int RG, int NTHR = DG_THREADS>
__device__ __forceinline__ void doubling_level(double* s, double* w, int sz, int warp, int g, int t) {
  constexpr int NW = NTHR / 32;
  const int tps = sz >> 3;              // 8x8 tiles per block edge
  const int gpp = tps * tps / RG;       // tile groups per pair
  const int ngroups = (TB / (2 * sz)) * gpp;
  const int gpl = tps / RG;             // groups per column (W) / per row (X)
  // W = L21 * X11: a group = RG vertically stacked tiles of one tile column (same k range, same B)
  for (int q = warp; q < ngroups; q += NW) {
    const int p = q / gpp, gi = q % gpp;
    const int o = p * 2 * sz;
    const int c0i = 8 * (gi / gpl), r0 = 8 * RG * (gi % gpl);
    double c[RG][2];
#pragma unroll
    for (int r = 0; r < RG; ++r) c[r][0] = c[r][1] = 0.0;
    tile_mma_multi<RG, true>(c, s + (o + sz + r0) + (o + c0i) * DG_LD, DG_LD, s + (o + c0i) + (o + c0i) * DG_LD, DG_LD,
                             sz - c0i, 0u, g, t);
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      double* wp = w + (p * sz + r0 + 8 * r + g) + (c0i + 2 * t) * DG_WLD;
      wp[0] = c[r][0];
      wp[DG_WLD] = c[r][1];
    }
  }
  __syncthreads();
  // X21 = -X22 * W: a group = RG tiles side by side in one tile row (same k range, same A)
  for (int q = warp; q < ngroups; q += NW) {
    const int p = q / gpp, gi = q % gpp;
    const int o = p * 2 * sz;
    const int r0 = 8 * (gi / gpl), c0i = 8 * RG * (gi % gpl);
    double c[RG][2];
#pragma unroll
    for (int r = 0; r < RG; ++r) c[r][0] = c[r][1] = 0.0;
    tile_mma_multi<RG, false>(c, s + (o + sz + r0) + (o + sz) * DG_LD, DG_LD, w + p * sz + c0i * DG_WLD, DG_WLD, r0 + 8,
                              0x80000000u, g, t);
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      double* xp = s + (o + sz + r0 + g) + (o + c0i + 8 * r + 2 * t) * DG_LD;
      xp[0] = c[r][0];
      xp[DG_LD] = c[r][1];
    }
  }
  __syncthreads();
}

// Reciprocal square root for the pivot chain: the library rsqrt() is a ~35-instruction dependent
// sequence (special-case handling + refinement) and dominates the per-column latency of the
// warp-level factorisation.  Pivots here are ordinary positive numbers (ve <= pivot <= sigma2+ve), so
// start from the hardware approximation (rsqrt.approx.f64, ~2^-22) and apply one cubically convergent
// Halley step: y' = y (1 + e/2 + 3e^2/8), e = 1 - x y^2  ->  relative error ~1e-19 before rounding.
// Non-positive / NaN pivots still produce NaN / inf and are flagged by the caller.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);
  const double q = e * fma(0.375, e, 0.5);
  return fma(y, q, y);
}

// 32x32 Cholesky of the diagonal block at `base` by one warp (lane = row).
__device__ __noinline__ void chol32_warp(double* s, double* rdiag, double* dvec, double* colbuf, int base,
                                            int lane, int kblock, int n_valid, int* info) {
  const unsigned FULL = 0xffffffffu;
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = s[(base + lane) + (base + c) * DG_LD];
  double ajj = __shfl_sync(FULL, a[0], 0);
  double rs = fast_rsqrt(ajj);
  // Branch-free bookkeeping: every lane keeps 1/L_jj and L_jj of ITS column (select, no divergent
  // branch inside the latency-bound column loop) and the first bad pivot it saw; stores after the loop.
  double rs_mine = 0.0, l_mine = 0.0;
  int bad = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double l = a[j] * rs;  // L[lane][j]
    a[j] = l;
    rs_mine = (lane == j) ? rs : rs_mine;
    l_mine = (lane == j) ? l : l_mine;
    bad = (!(ajj > 0.0) && j < bad) ? j : bad;  // uniform across the warp (ajj is a broadcast value)
    double ajj_n = 1.0, rs_n = 1.0;
    if (j + 1 < 32) {
      // lane j+1 owns the next pivot and can update it with its own l: start the broadcast + rsqrt now
      const double piv = fma(-l, l, a[j + 1]);
      ajj_n = __shfl_sync(FULL, piv, j + 1);
      rs_n = fast_rsqrt(ajj_n);
    }
    double* cb = colbuf + (j & 1) * 32;
    cb[lane] = l;
    __syncwarp();
#pragma unroll
    for (int c2 = ((j + 1) & ~1); c2 < 32; c2 += 2) {
      const double2 v = *reinterpret_cast<const double2*>(cb + c2);
      if (c2 > j) a[c2] = fma(-l, v.x, a[c2]);
      a[c2 + 1] = fma(-l, v.y, a[c2 + 1]);
    }
    ajj = ajj_n;
    rs = rs_n;
  }
  rdiag[base + lane] = rs_mine;
  dvec[base + lane] = l_mine;
  if (lane == 0 && bad != 0x7fffffff && kblock * TB + base + bad < n_valid) atomicCAS(info, 0, kblock * TB + base + bad + 1);
#pragma unroll
  for (int c = 0; c < 32; ++c) s[(base + lane) + (base + c) * DG_LD] = (c <= lane) ? a[c] : 0.0;
}

// Cholesky of the 128x128 tile held in shared memory `s` (stride DG_LD), in place (lower part = L;
// the part above the diagonal is left undefined except inside the 32x32 diagonal blocks, which are
// zeroed).  rdiag <- 1/L_jj, dvec <- L_jj.  All NTHR threads of the CTA must call it (NTHR = 512 in the
// standalone diagonal-block / small-fit kernels, 256 inside the persistent factorisation kernel).
// `pub` lets the caller hand block columns of L on as they become final (the persistent factorisation kernel
// publishes them to the CTAs that solve the next panel tiles against them).  Block column q < 3 (rows >= 32 q, 32
// columns) is final when phase B of step q starts: pub.background(q) is called there by warps 1 .. NW-1 only, so the
// stores, the fence and the flag never sit in front of warp 0, which factors the next diagonal block meanwhile.
// pub.store(3) is called by ALL threads at the end; the caller raises the flag after its next CTA barrier.
struct DiagNoPublish {
  __device__ __forceinline__ void background(int) const {}
  __device__ __forceinline__ void store(int) const {}
  __device__ __forceinline__ void flag(int) const {}
};
template <int NTHR = DG_THREADS, class Pub = DiagNoPublish>
__device__ __forceinline__ void diag_cholesky_smem(double* s, double* rdiag, double* dvec, double* colbuf, int kblock,
                                                   int n_valid, int* info, const Pub pub = Pub()) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  constexpr int NW = NTHR / 32;
  // ---------------- Cholesky, 4 block columns of width 32 ----------------
  if (warp == 0) chol32_warp(s, rdiag, dvec, colbuf, 0, lane, kblock, n_valid, info);
  __syncthreads();
  DG_TS(2);
#pragma unroll 1
  for (int kb = 0; kb < 3; ++kb) {
    const int base = kb * 32;
    // panel: rows below the diagonal block, forward substitution per row (lane = row)
    if (warp < 3 - kb) {
      const int row = base + 32 * (warp + 1) + lane;
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = s[row + (base + c) * DG_LD];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const double x = a[c] * rdiag[base + c];
        a[c] = x;
        const double* lcol = s + base + (base + c) * DG_LD;  // column c of L_kk, contiguous
#pragma unroll
        for (int c2 = ((c + 1) & ~1); c2 < 32; c2 += 2) {
          const double2 v = *reinterpret_cast<const double2*>(lcol + c2);
          if (c2 > c) a[c2] = fma(-x, v.x, a[c2]);
          a[c2 + 1] = fma(-x, v.y, a[c2 + 1]);
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) s[row + (base + c) * DG_LD] = a[c];
    }
    __syncthreads();
    DG_TS(3 + 3 * kb);
    // trailing update, phase A: the next diagonal block (16 tiles, one per warp with 16 warps)
    for (int tl = warp; tl < 16; tl += NW) {
      const int nb0 = base + 32;
      const int row0 = nb0 + 8 * (tl >> 2), col0 = nb0 + 8 * (tl & 3);
      double* cp = s + (row0 + g) + (col0 + 2 * t) * DG_LD;
      double c0 = cp[0], c1 = cp[DG_LD];
      tile_mma<true>(c0, c1, s + row0 + base * DG_LD, DG_LD, s + col0 + base * DG_LD, DG_LD, 32, 0x80000000u, g, t);
      cp[0] = c0;
      cp[DG_LD] = c1;
    }
    __syncthreads();
    DG_TS(4 + 3 * kb);
    // phase B: warp 0 factors the next diagonal block while warps 1..15 update the blocks below it
    if (warp == 0) {
      chol32_warp(s, rdiag, dvec, colbuf, base + 32, lane, kblock, n_valid, info);
    } else {
      pub.background(kb);
      const int m = 2 - kb;                     // block rows below the next diagonal block
      const int nblk = m * (m + 3) / 2;         // blocks (bi, bj), bi in 1..m, bj in 0..bi  (relative to kb+1)
      for (int q = warp - 1; q < nblk * 16; q += NW - 1) {
        const int blk = q >> 4, tile = q & 15;
        int bi = (blk < 2) ? 1 : 2;             // m <= 2: blocks (1,0),(1,1),(2,0),(2,1),(2,2)
        int bj = blk - (bi == 1 ? 0 : 2);
        const int row0 = base + 32 * (bi + 1) + 8 * (tile >> 2);
        const int col0 = base + 32 * (bj + 1) + 8 * (tile & 3);
        double* cp = s + (row0 + g) + (col0 + 2 * t) * DG_LD;
        double c0 = cp[0], c1 = cp[DG_LD];
        tile_mma<true>(c0, c1, s + row0 + base * DG_LD, DG_LD, s + col0 + base * DG_LD, DG_LD, 32, 0x80000000u, g, t);
        cp[0] = c0;
        cp[DG_LD] = c1;
      }
    }
    __syncthreads();
    DG_TS(5 + 3 * kb);
  }
  pub.store(3);
}

// In-place inverse of the lower-triangular factor held in `s` (needs rdiag from diag_cholesky_smem):
// on return the lower part of `s` is L^-1.  `w` is a 64 x DG_WLD scratch.
template <int NTHR = DG_THREADS>
__device__ __forceinline__ void diag_invert_smem(double* s, double* w, const double* rdiag) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  // ---------------- inverse: 8x8 diagonal blocks, one thread per column ----------------
  {
    double x[8];
    const int b8 = (tid >> 3) * 8, c = tid & 7;  // threads 0..127: block tid/8, column tid%8
    if (tid < TB) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int mm = 0; mm < r; ++mm) acc = fma(s[(b8 + r) + (b8 + mm) * DG_LD], x[mm], acc);
        const double rd = rdiag[b8 + r];
        x[r] = (r == c) ? rd : ((r < c) ? 0.0 : -acc * rd);
      }
    }
    __syncthreads();
    if (tid < TB) {
#pragma unroll
      for (int r = 0; r < 8; ++r) s[(b8 + r) + (b8 + c) * DG_LD] = x[r];
    }
  }
  __syncthreads();
  DG_TS(15);
  // ---------------- recursive doubling: block sizes 8, 16, 32, 64 ----------------
  doubling_level<1, NTHR>(s, w, 8, warp, g, t);    //  8 tiles
  doubling_level<1, NTHR>(s, w, 16, warp, g, t);   // 16 tiles
  doubling_level<2, NTHR>(s, w, 32, warp, g, t);   // 32 tiles, 2 per warp
  doubling_level<4, NTHR>(s, w, 64, warp, g, t);   // 64 tiles, 4 per warp
}

__global__ void __launch_bounds__(DG_THREADS, 1)
    diag_block_kernel(double* __restrict__ Akk, size_t lda, double* __restrict__ Dkk, size_t ldd, int kblock,
                      int n_valid, double* __restrict__ logdet_part, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* s = sm;                   // [128][132]
  double* w = sm + TB * DG_LD;      // [64][68] scratch
  double* rdiag = w + 64 * DG_WLD;  // 1 / L_jj
  double* dvec = rdiag + TB;        // L_jj, later reduction scratch
  double* colbuf = dvec + TB;       // [2][32] pivot-column broadcast buffer
  const int tid = threadIdx.x;
#ifdef GPEDM_TIMELINE
  const unsigned long long tl_t0 = tl_now();
#endif

  // The tile moves between global and shared memory as its 10 lower 32x32 blocks, two rows per thread
  // (16-byte accesses).  Nothing reads the part of L_kk above the diagonal blocks (panel / trailing /
  // inverse kernels take D_kk and the off-diagonal tiles), and the upper part of the D_kk tile in
  // the X buffer is zeroed once when the handle is created and never written again.
  auto lower_block_elem = [](int idx, int& r, int& c, bool& on_diag) {
    const int b = idx >> 9, e = idx & 511;  // block 0..9 (row-major over the lower triangle), 512 row pairs each
    const int bi = (b >= 6) ? 3 : ((b >= 3) ? 2 : ((b >= 1) ? 1 : 0)), bj = b - bi * (bi + 1) / 2;
    r = 32 * bi + 2 * (e & 15);
    c = 32 * bj + (e >> 4);
    on_diag = bi == bj;
  };
  DG_TS(0);
#pragma unroll 5
  for (int idx = tid; idx < 10 * 512; idx += DG_THREADS) {
    int r, c;
    bool dg;
    lower_block_elem(idx, r, c, dg);
    *reinterpret_cast<double2*>(&s[r + c * DG_LD]) = *reinterpret_cast<const double2*>(&Akk[(size_t)r + (size_t)c * lda]);
  }
  __syncthreads();
  DG_TS(1);
  diag_cholesky_smem(s, rdiag, dvec, colbuf, kblock, n_valid, info);
  // ---------------- write L, log-determinant part ----------------
#pragma unroll 5
  for (int idx = tid; idx < 10 * 512; idx += DG_THREADS) {
    int r, c;
    bool dg;
    lower_block_elem(idx, r, c, dg);
    double2 v = *reinterpret_cast<const double2*>(&s[r + c * DG_LD]);
    if (dg) {
      if (r < c) v.x = 0.0;
      if (r + 1 < c) v.y = 0.0;
    }
    *reinterpret_cast<double2*>(&Akk[(size_t)r + (size_t)c * lda]) = v;
  }
  {
    // sum of log L_jj over the valid rows: shuffle tree per warp, then the 4 warp partials in order
    double v = 0.0;
    if (tid < TB && kblock * TB + tid < n_valid) v = log(dvec[tid]);
    __syncthreads();
    if (tid < TB) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if ((tid & 31) == 0) dvec[tid >> 5] = v;
    }
    __syncthreads();
    if (tid == 0) logdet_part[kblock] = ((dvec[0] + dvec[1]) + dvec[2]) + dvec[3];
  }
  DG_TS(14);

  diag_invert_smem(s, w, rdiag);
  DG_TS(17);
#pragma unroll 5
  for (int idx = tid; idx < 10 * 512; idx += DG_THREADS) {
    int r, c;
    bool dg;
    lower_block_elem(idx, r, c, dg);
    double2 v = *reinterpret_cast<const double2*>(&s[r + c * DG_LD]);
    if (dg) {
      if (r < c) v.x = 0.0;
      if (r + 1 < c) v.y = 0.0;
    }
    *reinterpret_cast<double2*>(&Dkk[(size_t)r + (size_t)c * ldd]) = v;
  }
  DG_TS(18);
#ifdef GPEDM_TIMELINE
  __syncthreads();
  tl_log(tl_t0, 100 | (1 << 8));
#endif
}

}  // namespace gpedm
