// Tile operations of the evaluation pipeline.  Every O(N^3) step (Cholesky panel and trailing
// update, triangular inverse, L^-T L^-1 product, prediction solve) is a set of independent
// 128x128 output tiles  C = (+/-) sum_k Aop Bop (+ C)  with tile-dependent k ranges.  One launch =
// one such set, described by a POD `TileOp`; `decode_tile` maps a tile index to operand origins
// (longest k range first, so the hardware block scheduler balances the tail) and `tile_once_kernel`
// runs the DMMA contraction of dgemm_tile.cuh on it.
#pragma once
#include "dgemm_tile.cuh"

namespace gpedm {

enum TileOpType {
  OP_PANEL = 0,     // L_ik = A_ik D_k^T, i > k                         (MN, MN)
  OP_TRAIL = 1,     // A_ij -= sum_{q=k}^{k+kcount-1} L_iq L_jq^T, j >= jlo, i >= j   (MN, MN)
  OP_TRTRI_W = 2,   // W = L21 X11   (pairs of size s)                  (MN, K)
  OP_TRTRI_X = 3,   // X21 = -X22 W                                     (MN, K)
  OP_LAUUM = 4,     // S_ij = sum_{k>=i} X_ki^T X_kj, i >= j            (K, K)
  OP_TRI_APPLY = 5  // U = X V  (X lower triangular, V = N x M)         (MN, K)
};

#ifdef GPEDM_TIMELINE
// Debug build (-DGPEDM_TIMELINE): every CTA logs {start ns, end ns, SM id, op type | tag << 8} so a
// host tool can reconstruct the SM occupancy of one evaluation (tools/timeline.py).
constexpr unsigned int TL_MAX = 1u << 18;
__device__ unsigned long long g_tl_rec[4 * TL_MAX];
__device__ unsigned int g_tl_count;
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned int tl_smid() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tl_log(unsigned long long t0, long long code) {
  if (threadIdx.x == 0) {
    const unsigned int i = atomicAdd(&g_tl_count, 1u);
    if (i < TL_MAX) {
      g_tl_rec[4 * i] = t0;
      g_tl_rec[4 * i + 1] = tl_now();
      g_tl_rec[4 * i + 2] = tl_smid();
      g_tl_rec[4 * i + 3] = (unsigned long long)code;
    }
  }
}
#endif

struct TileOp {
  int type;
  int ntiles;
  int nb;        // tiles per matrix edge
  int k;         // block column (panel) / first block column of the update (trail)
  int kcount;    // trail: number of consecutive block columns applied (rank = 128 * kcount)
  int s;         // trtri block size in tiles
  int p0;        // trtri: first pair covered by this launch
  int nfull;     // trtri: number of full (s x s tile) pairs covered; a ragged last pair may follow
  int jlo, jhi;  // trail column range
  int mt;        // tri_apply: number of column tiles of V
  int tag;       // which stream launched it (0 main, 1 chain, 2 fill): timeline builds only
  const double* P0;  // operand matrices (meaning depends on type)
  const double* P1;
  double* P2;
  size_t ld0, ld1, ld2;
};

// What the kernel needs to know about one tile.  Strides are per-launch constants (TileOp) so
// they stay out of the per-tile register state.
struct TileWork {
  const double* A;  // operand tile origins, already advanced to the first k slab
  const double* B;
  double* C;        // output tile (also the accumulate-from tile when flags & 2)
  int nslab;        // number of BK-wide k slabs
  int flags;        // bit0: negate the product, bit1: accumulate onto C
};

// tiles of the lower triangle of an m x m tile grid enumerated column by column:
// column c holds rows c..m-1; returns (row, col) of linear index t.
__device__ __forceinline__ void colmajor_tri_decode(int t, int m, int& row, int& col) {
  double b = 2.0 * m + 1.0;
  int c = (int)((b - sqrt(b * b - 8.0 * (double)t)) * 0.5);
  if (c < 0) c = 0;
  if (c > m - 1) c = m - 1;
  while (c > 0 && (long long)c * m - (long long)c * (c - 1) / 2 > t) --c;
  while (c < m - 1 && (long long)(c + 1) * m - (long long)(c + 1) * c / 2 <= t) ++c;
  col = c;
  row = c + (t - (c * m - c * (c - 1) / 2));
}

// Operand strides of a launch: lda/ldb apply to the A/B operands, ldc to the output.
__device__ __forceinline__ void op_strides(const TileOp& op, size_t& lda, size_t& ldb, size_t& ldc) {
  switch (op.type) {
    case OP_PANEL: lda = op.ld2; ldb = op.ld1; ldc = op.ld2; break;
    case OP_TRAIL: lda = ldb = ldc = op.ld2; break;
    case OP_LAUUM: lda = ldb = op.ld0; ldc = op.ld2; break;
    default: lda = op.ld0; ldb = op.ld1; ldc = op.ld2; break;
  }
}

// Tile t of a launch -> operand origins and k range.  A_KMAJOR/B_KMAJOR say how a k offset moves
// the operand pointers.  Kept out of line: it runs once per tile, not per slab.
template <bool A_KMAJOR, bool B_KMAJOR, int BK>
__device__ __forceinline__ bool decode_tile(const TileOp& op, int t, TileWork& w) {
  if (t >= op.ntiles) {
    w.nslab = 0;
    return false;
  }
  const double *A, *B;
  double* C;
  size_t lda, ldb, ldc;
  op_strides(op, lda, ldb, ldc);
  int kbeg = 0, kend = TB, flags = 0;
  switch (op.type) {
    case OP_PANEL: {
      const int i = op.k + 1 + t;
      C = op.P2 + (size_t)i * TB + (size_t)op.k * TB * ldc;
      A = C;
      B = op.P1 + (size_t)op.k * TB + (size_t)op.k * TB * ldb;  // D_k
      break;
    }
    case OP_TRAIL: {
      int r, c;
      colmajor_tri_decode(t, op.nb - op.jlo, r, c);
      const int i = op.jlo + r, j = op.jlo + c;
      A = op.P2 + (size_t)i * TB + (size_t)op.k * TB * ldc;
      B = op.P2 + (size_t)j * TB + (size_t)op.k * TB * ldc;
      C = op.P2 + (size_t)i * TB + (size_t)j * TB * ldc;
      kend = op.kcount * TB;
      flags = 3;
      break;
    }
    case OP_TRTRI_W:
    case OP_TRTRI_X: {
      // pairs p cover tile rows/cols [2ps, 2ps+2s); all but possibly the last are full (s x s
      // tiles), the last may have only s2 < s rows (ragged nb).  No empty tiles are enumerated.
      const int s = op.s, per = s * s;
      int p, rem, s2;
      if (t < op.nfull * per) {
        p = op.p0 + t / per;
        rem = t % per;
        s2 = s;
      } else {  // ragged last pair of the level: only s2 < s block rows exist below the split
        p = op.nb / (2 * s);
        rem = t - op.nfull * per;
        s2 = op.nb - 2 * s * p - s;
      }
      const int o = p * 2 * s;
      if (op.type == OP_TRTRI_W) {
        const int j = rem / s2, i = rem % s2;  // ascending j: longest k range first
        A = op.P0 + (size_t)(o + s + i) * TB + (size_t)o * TB * lda;   // L21 row block
        B = op.P1 + (size_t)o * TB + (size_t)(o + j) * TB * ldb;       // X11 column block
        C = op.P2 + (size_t)(o + s + i) * TB + (size_t)(o + j) * TB * ldc;
        kbeg = j * TB;
        kend = s * TB;
      } else {
        const int i = s2 - 1 - rem / s, j = rem % s;  // descending i: longest k range first
        A = op.P0 + (size_t)(o + s + i) * TB + (size_t)(o + s) * TB * lda;  // X22 row block
        B = op.P1 + (size_t)(o + s) * TB + (size_t)(o + j) * TB * ldb;      // W column block
        C = op.P2 + (size_t)(o + s + i) * TB + (size_t)(o + j) * TB * ldc;
        kend = (i + 1) * TB;
        flags = 1;
      }
      break;
    }
    case OP_LAUUM: {
      int i, j;
      tri_decode(t, i, j);  // ascending i: longest k range first
      A = op.P0 + (size_t)i * TB * lda;
      B = op.P0 + (size_t)j * TB * ldb;
      C = op.P2 + (size_t)i * TB + (size_t)j * TB * ldc;
      kbeg = i * TB;
      kend = op.nb * TB;
      break;
    }
    default: {  // OP_TRI_APPLY
      const int i = op.nb - 1 - t / op.mt, jt = t % op.mt;
      A = op.P0 + (size_t)i * TB;
      B = op.P1 + (size_t)jt * TB * ldb;
      C = op.P2 + (size_t)i * TB + (size_t)jt * TB * ldc;
      kend = (i + 1) * TB;
      break;
    }
  }
  w.A = A + (A_KMAJOR ? (size_t)kbeg : (size_t)kbeg * lda);
  w.B = B + (B_KMAJOR ? (size_t)kbeg : (size_t)kbeg * ldb);
  w.C = C;
  w.nslab = (kend - kbeg) / BK;
  w.flags = flags;
  return true;
}

// One tile per CTA (grid = ntiles, hardware block scheduler hands out tiles in launch order).
// With Cfg::TM < 128 / Cfg::TN < 128 each 128 x 128 tile is split into row strips / column halves
// handled by consecutive CTAs.  Row strips keep the in-place panel update safe (a CTA reads only
// the rows of A it overwrites); column halves are used for the out-of-place operations only.
template <bool A_KMAJOR, bool B_KMAJOR, class Cfg = GemmCfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINCTA) tile_once_kernel(const TileOp op) {
  extern __shared__ double smem[];
  constexpr int SPLIT_M = TB / Cfg::TM, SPLIT_N = TB / Cfg::TN;
  static_assert(SPLIT_M == 1 || !A_KMAJOR, "row-strip split is implemented for MN-major A only");
#ifdef GPEDM_TIMELINE
  const unsigned long long tl_t0 = tl_now();
#endif
  TileWork w;
  if (!decode_tile<A_KMAJOR, B_KMAJOR, Cfg::BK>(op, blockIdx.x / (SPLIT_M * SPLIT_N), w)) return;
  size_t lda, ldb, ldc;
  op_strides(op, lda, ldb, ldc);
  const int sub = blockIdx.x % (SPLIT_M * SPLIT_N);
  const int roff = (sub % SPLIT_M) * Cfg::TM, coff = (sub / SPLIT_M) * Cfg::TN;
  const double* Bp = w.B + (B_KMAJOR ? (size_t)coff * ldb : (size_t)coff);
  double* Cp = w.C + roff + (size_t)coff * ldc;
  gemm_tile_mb<A_KMAJOR, B_KMAJOR, Cfg>(w.A + roff, lda, Bp, ldb, 0, w.nslab * Cfg::BK, (w.flags & 1) != 0,
                                        (w.flags & 2) ? Cp : nullptr, ldc, Cp, ldc, smem);
#ifdef GPEDM_TIMELINE
  __syncthreads();
  // code = op type | stream tag << 8 | k slabs << 16 | output rows << 40 | output cols << 52
  tl_log(tl_t0, op.type | (op.tag << 8) | ((long long)w.nslab << 16) | ((long long)Cfg::TM << 40) | ((long long)Cfg::TN << 52));
#endif
}

}  // namespace gpedm
