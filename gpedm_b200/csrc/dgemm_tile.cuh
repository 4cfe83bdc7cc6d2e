// FP64 tensor-pipe (DMMA) tile contraction for sm_100a.
//
// One CTA (256 threads, 8 warps) computes one 128x128 tile
//     C = (+/-) sum_k Aop[m,k] * Bop[k,n]  (+ Cin)
// over an arbitrary k range (multiple of 16).  This is the workhorse behind every
// O(N^3) step that replaces the reference's LAPACK calls (dpotrf / dpotri behind
// Matrix::chol / chol2inv, reference R/GPfunctions.R:826-829): the Cholesky panel
// and trailing update, the triangular inverse and the L^-T L^-1 product.
//
// Hardware mapping: tcgen05/UMMA has no f64 kind, so the FP64 tensor path on
// sm_100a is warp-level `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4).  Each warp owns
// a 64x32 sub-tile = 8x4 DMMA tiles = 64 FP64 accumulators per thread; operands
// are staged global -> shared with a 4-deep cp.async (LDGSTS) pipeline of
// 16-wide k slabs, in layouts padded so that every fragment LDS.64 is
// bank-conflict-free.
//
// Operand layouts (all matrices are column-major in global memory):
//   A "MN-major": element (m,k) at A[m + k*lda]   (m contiguous; plain  "N")
//   A "K-major" : element (m,k) at A[k + m*lda]   (k contiguous; i.e.   "T")
//   B "MN-major": element (k,n) at B[n + k*ldb]   (n contiguous; i.e.   "T")
//   B "K-major" : element (k,n) at B[k + n*ldb]   (k contiguous; plain  "N")
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpedm {

constexpr int TB = 128;  // tile edge

// Compile-time shape of the contraction kernel: a TM x 128 output tile computed by WM x WN warps,
// each owning a (TM/WM) x (128/WN) sub-tile; k slabs of BK columns, NSTAGE-deep cp.async pipeline.
// TM = 128 is the bulk shape; TM = 32 splits one 128 x 128 tile over four CTAs for the
// latency-critical chain of the Cholesky (panel + next-column update), where there are fewer
// tiles than SMs.
template <int WM_, int WN_, int BK_, int NSTAGE_, int TM_ = 128, int TN_ = 128, int MINCTA_ = 1>
struct GemmCfgT {
  static constexpr int WM = WM_, WN = WN_, BK = BK_, NSTAGE = NSTAGE_, TM = TM_, TN = TN_, MINCTA = MINCTA_;
  static constexpr int THREADS = WM * WN * 32;
  static constexpr int MT = TM / WM / 8;  // 8-row DMMA tiles per warp along m
  static constexpr int NT = TN / WN / 8;  // 8-col DMMA tiles per warp along n
  static constexpr int LD_MN_A = TM + 4;  // smem row stride (doubles) of an MN-major A slab [BK][TM+4]
  static constexpr int LD_MN_B = TN + 4;  // ... of an MN-major B slab [BK][TN+4]
  static constexpr int LD_K = BK + 4;     // smem row stride of a K-major slab [rows][BK+4]
  static constexpr int SLAB_A = (BK * LD_MN_A > TM * LD_K) ? BK * LD_MN_A : TM * LD_K;
  static constexpr int SLAB_B = (BK * LD_MN_B > TN * LD_K) ? BK * LD_MN_B : TN * LD_K;
  static constexpr int SMEM_BYTES = NSTAGE * (SLAB_A + SLAB_B) * (int)sizeof(double);
  static_assert(TM % (WM * 8) == 0 && TN % (WN * 8) == 0, "warp grid must tile the output");
  static_assert((TM + 4) % 16 == 4 && (TN + 4) % 16 == 4 && (BK + 4) % 16 == 4,
                "strides must be 4 mod 16 for conflict-free LDS.64");
};
#ifndef GPEDM_GEMM_CFG
#define GPEDM_GEMM_CFG GemmCfgT<2, 4, 16, 4>
#endif
using GemmCfg = GPEDM_GEMM_CFG;
using GemmCfgSplit = GemmCfgT<1, 8, 16, 4, 32>;  // 32 x 128 tiles, 8 warps of 32 x 16
// 128 x 64 half tiles, 4 warps of 64 x 32, two CTAs resident per SM: the prologue (accumulate-from
// load, pipeline fill) and epilogue (store) of one CTA overlap the k loop of the other.
using GemmCfgHalf = GemmCfgT<2, 2, 16, 4, 128, 64, 2>;
constexpr int GEMM_THREADS = GemmCfg::THREADS;
constexpr int GEMM_SMEM_BYTES = GemmCfg::SMEM_BYTES;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double flip_sign(double x, unsigned mask) {
  return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

// Stage one BK-wide k slab of a ROWS-row operand into shared memory.
template <class Cfg, bool KMAJOR, int ROWS>
__device__ __forceinline__ void load_slab(double* __restrict__ s, const double* __restrict__ g, size_t ld,
                                          int k0, int tid) {
  constexpr int CHUNKS = ROWS * Cfg::BK / 2;  // 16-byte chunks per slab
  constexpr int ITERS = (CHUNKS + Cfg::THREADS - 1) / Cfg::THREADS;
  if (!KMAJOR) {
    // global: BK columns (k) of ROWS contiguous doubles.
    constexpr int CPR = ROWS / 2;  // chunks per k row
    constexpr int LDS = ROWS + 4;
#pragma unroll
    for (int r = 0; r < ITERS; ++r) {
      int chunk = tid + r * Cfg::THREADS;
      if (CHUNKS % Cfg::THREADS == 0 || chunk < CHUNKS) {
        int row = chunk / CPR;         // k within slab
        int c2 = (chunk % CPR) << 1;   // m offset
        cp_async16(s + row * LDS + c2, g + (size_t)(k0 + row) * ld + c2);
      }
    }
  } else {
    // global: ROWS columns (m) of BK contiguous doubles (k).
    constexpr int CPR = Cfg::BK / 2;  // chunks per m row
#pragma unroll
    for (int r = 0; r < ITERS; ++r) {
      int chunk = tid + r * Cfg::THREADS;
      if (CHUNKS % Cfg::THREADS == 0 || chunk < CHUNKS) {
        int row = chunk / CPR;         // m
        int c2 = (chunk % CPR) << 1;   // k offset within slab
        cp_async16(s + row * Cfg::LD_K + c2, g + (size_t)row * ld + k0 + c2);
      }
    }
  }
}

// Main contraction of one TM x 128 tile.  `smem` must provide Cfg::SMEM_BYTES.  kbeg/kend are
// element indices along k (multiples of BK) applied to both operands' k axes.
// If Cin != nullptr the accumulator starts from the Cin tile (fetched before the
// k loop so its latency hides under the first slabs), else from zero.
template <bool A_KMAJOR, bool B_KMAJOR, class Cfg = GemmCfg>
__device__ __forceinline__ void gemm_tile(const double* __restrict__ A, size_t lda, const double* __restrict__ B,
                                          size_t ldb, int kbeg, int kend, bool negate, const double* Cin,
                                          size_t ldcin, double* __restrict__ Cout, size_t ldc, double* smem) {
  constexpr int MT = Cfg::MT, NT = Cfg::NT, BK = Cfg::BK, NS = Cfg::NSTAGE;
  constexpr int SLAB_A = Cfg::SLAB_A, SLAB_B = Cfg::SLAB_B;
  constexpr int LD_A = Cfg::LD_MN_A, LD_B = Cfg::LD_MN_B, LD_K = Cfg::LD_K;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = (warp % Cfg::WM) * (MT * 8), n0 = (warp / Cfg::WM) * (NT * 8);

  double* sA = smem;
  double* sB = smem + NS * SLAB_A;

  const int nk = (kend - kbeg) / BK;
  const unsigned sgn = negate ? 0x80000000u : 0u;  // sign flip on the integer pipe, not the FP64 pipe
#pragma unroll
  for (int s = 0; s < NS - 1; ++s) {
    if (s < nk) {
      load_slab<Cfg, A_KMAJOR, Cfg::TM>(sA + s * SLAB_A, A, lda, kbeg + s * BK, tid);
      load_slab<Cfg, B_KMAJOR, Cfg::TN>(sB + s * SLAB_B, B, ldb, kbeg + s * BK, tid);
    }
    cp_async_commit();
  }

  double acc[MT][NT][2];
  if (Cin != nullptr) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          acc[i][j][e] = Cin[(size_t)(m0 + 8 * i + g) + (size_t)(n0 + 8 * j + 2 * t + e) * ldcin];
  } else {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<NS - 2>();
    __syncthreads();
    {
      int nxt = kt + NS - 1;
      if (nxt < nk) {
        int sb = nxt % NS;
        load_slab<Cfg, A_KMAJOR, Cfg::TM>(sA + sb * SLAB_A, A, lda, kbeg + nxt * BK, tid);
        load_slab<Cfg, B_KMAJOR, Cfg::TN>(sB + sb * SLAB_B, B, ldb, kbeg + nxt * BK, tid);
      }
      cp_async_commit();
    }
    const double* a_s = sA + (kt % NS) * SLAB_A;
    const double* b_s = sB + (kt % NS) * SLAB_B;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double a[MT], b[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        double x = A_KMAJOR ? a_s[(m0 + 8 * i + g) * LD_K + k4 * 4 + t] : a_s[(k4 * 4 + t) * LD_A + m0 + 8 * i + g];
        a[i] = flip_sign(x, sgn);
      }
#pragma unroll
      for (int j = 0; j < NT; ++j)
        b[j] = B_KMAJOR ? b_s[(n0 + 8 * j + g) * LD_K + k4 * 4 + t] : b_s[(k4 * 4 + t) * LD_B + n0 + 8 * j + g];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        Cout[(size_t)(m0 + 8 * i + g) + (size_t)(n0 + 8 * j + 2 * t + e) * ldc] = acc[i][j][e];
}

// ---- mbarrier helpers (shared::cta) -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
// arrive-on-completion of all cp.async issued so far by this thread (does not change the pending count)
__device__ __forceinline__ void mbar_cp_async_arrive(uint64_t* bar) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @!p bra WAIT_%=;\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}

// Same contraction as gemm_tile, but the CTA-wide __syncthreads per k slab is replaced by per-stage
// mbarriers: full[s] completes when every thread's cp.async copies of the slab have landed, empty[s]
// when every warp has finished reading it.  Slabs are prefetched NS-2 ahead into a ring of NS
// stages, so warps may drift up to two slabs apart instead of meeting at a barrier every slab:
// a warp that waits (LDS latency, copy not landed yet) no longer stops the other warp on its
// scheduler from feeding the DMMA pipe.
template <bool A_KMAJOR, bool B_KMAJOR, class Cfg = GemmCfg>
__device__ __forceinline__ void gemm_tile_mb(const double* __restrict__ A, size_t lda, const double* __restrict__ B,
                                             size_t ldb, int kbeg, int kend, bool negate, const double* Cin,
                                             size_t ldcin, double* __restrict__ Cout, size_t ldc, double* smem) {
  constexpr int MT = Cfg::MT, NT = Cfg::NT, BK = Cfg::BK, NS = Cfg::NSTAGE, PREF = NS - 2;
  constexpr int SLAB_A = Cfg::SLAB_A, SLAB_B = Cfg::SLAB_B;
  constexpr int LD_A = Cfg::LD_MN_A, LD_B = Cfg::LD_MN_B, LD_K = Cfg::LD_K;
  static_assert(NS >= 3, "need at least 3 stages");
  __shared__ uint64_t full_bar[NS], empty_bar[NS];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = (warp % Cfg::WM) * (MT * 8), n0 = (warp / Cfg::WM) * (NT * 8);
  double* sA = smem;
  double* sB = smem + NS * SLAB_A;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], Cfg::THREADS);
      mbar_init(&empty_bar[s], Cfg::THREADS / 32);
    }
  }
  __syncthreads();

  const int nk = (kend - kbeg) / BK;
  const unsigned sgn = negate ? 0x80000000u : 0u;
  auto issue = [&](int j) {  // copy slab j into stage j % NS
    const int s = j % NS, r = j / NS;
    if (r > 0) mbar_wait(&empty_bar[s], (unsigned)((r - 1) & 1));
    load_slab<Cfg, A_KMAJOR, Cfg::TM>(sA + s * SLAB_A, A, lda, kbeg + j * BK, tid);
    load_slab<Cfg, B_KMAJOR, Cfg::TN>(sB + s * SLAB_B, B, ldb, kbeg + j * BK, tid);
    mbar_cp_async_arrive(&full_bar[s]);
  };
#pragma unroll
  for (int j = 0; j < PREF; ++j)
    if (j < nk) issue(j);

  double acc[MT][NT][2];
  if (Cin != nullptr) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          acc[i][j][e] = Cin[(size_t)(m0 + 8 * i + g) + (size_t)(n0 + 8 * j + 2 * t + e) * ldcin];
  } else {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }

  for (int kt = 0; kt < nk; ++kt) {
    if (kt + PREF < nk) issue(kt + PREF);
    const int s = kt % NS;
    mbar_wait(&full_bar[s], (unsigned)((kt / NS) & 1));
    const double* a_s = sA + s * SLAB_A;
    const double* b_s = sB + s * SLAB_B;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double a[MT], b[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        double x = A_KMAJOR ? a_s[(m0 + 8 * i + g) * LD_K + k4 * 4 + t] : a_s[(k4 * 4 + t) * LD_A + m0 + 8 * i + g];
        a[i] = flip_sign(x, sgn);
      }
#pragma unroll
      for (int j = 0; j < NT; ++j)
        b[j] = B_KMAJOR ? b_s[(n0 + 8 * j + g) * LD_K + k4 * 4 + t] : b_s[(k4 * 4 + t) * LD_B + n0 + 8 * j + g];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        Cout[(size_t)(m0 + 8 * i + g) + (size_t)(n0 + 8 * j + 2 * t + e) * ldc] = acc[i][j][e];
}

// Decode a linear index into lower-triangle coordinates (i >= j), row-major over rows:
// t = i*(i+1)/2 + j.
__device__ __forceinline__ void tri_decode(int t, int& i, int& j) {
  int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((long long)(r + 1) * (r + 2) / 2 <= t) ++r;
  while ((long long)r * (r + 1) / 2 > t) --r;
  i = r;
  j = t - r * (r + 1) / 2;
}

}  // namespace gpedm
