// Persistent, dependency-driven factorisation kernel: Cholesky (left-looking by tiles) and the
// triangular inverse X = L^-1 (row form) as ONE launch per evaluation.
//
// Replaces LAPACK dpotrf + the dtrtri half of dpotri behind Matrix::chol / chol2inv (reference
// R/GPfunctions.R:826-829).  Round 1 ran these as ~600 stream-ordered launches (diagonal block,
// panel, grouped trailing updates on three priority streams); the serial chain of the Cholesky -
// and the SMs its kernels could not get while long low-priority CTAs were resident - cost 2-4 ms
// of the 20 ms evaluation at N = 8182 and most of the step below N = 4096.  Here one CTA per SM stays
// resident and pulls TASKS (one 128 x 128 output tile each) from an ordered list:
//
//   DIAG (j)      T = A_jj - sum_{k<j} L_jk L_jk^T ;  L_jj = chol(T) ;  D_j = L_jj^-1   (in shared memory)
//   POTRF (i,j)   T = A_ij - sum_{k<j} L_ik L_jk^T ;  L_ij = T D_j^T                    (i > j)
//   TRTRI (i,j)   T = sum_{k=j}^{i-1} L_ik X_kj ;     X_ij = -D_i T                     (i > j, X_jj = D_j)
//   LAUUM (i,j)   S_ij = sum_{k>=i} X_ki^T X_kj                                         (i >= j; Sigma^-1 = X^T X)
//
// Every tile is written exactly once, by one CTA, after ONE long k loop (the DMMA contraction of
// dgemm_tile.cuh at its long-k rate) followed by a 128^3 epilogue product against the diagonal-block
// inverse.  Dependencies are per-tile ready flags in global memory (release/acquire at gpu scope):
// the k loop polls the flags of the operand tiles it is about to stream, so a CTA pre-accumulates
// everything that is already final and only its last k tiles wait for the chain.  Tasks are claimed in
// list order through an atomic counter and every dependency of a task precedes it in the list, so the
// lowest unfinished task is always held by a running CTA whose inputs are complete: no deadlock,
// whatever else shares the GPU.  The list interleaves the rows of the inverse two block columns
// behind the Cholesky, so CTAs that would wait for the chain find inverse tiles whose inputs have
// long been final.
// Results are bit-reproducible: each tile's summation order is fixed (k ascending), scheduling
// only decides when a tile is computed, never how.
// The serial chain DIAG(c) -> L(c+1, c) -> DIAG(c+1) bounds every factorisation below ~6000 rows; two hand-off rules
// shorten it: tiles next to the diagonal are handed on in 32-column STRIPS (dag_strips_ready),
// and the POTRF tiles L(c+1 .. c+t, c) are SOLVED against the block columns of L_cc as DIAG(c) publishes them
// (DagPublishL / dag_trsm_chain) instead of waiting for D_c.  Which tiles are solved depends on their position only.
#pragma once
#include "dgemm_tile.cuh"
#include "diag_block.cuh"

namespace gpedm {

constexpr int DAG_THREADS = 256;
using DagCfg = GemmCfgT<2, 4, 16, 4>;
constexpr int DAG_NS = DagCfg::NSTAGE;
// shared memory (doubles): [0, P): operand ring of the k loop, 4 stages x (A slab + B slab); the same
// region then holds the resident tile T (128 x 132, column-major) for the epilogue / the in-CTA
// Cholesky.  [P, P+E): ring of the one streamed operand of the epilogue (4 x 16 x 132), aliased by
// the diagonal-block scratch (64 x 68 + pivots).
constexpr int DAG_P = DAG_NS * (DagCfg::SLAB_A + DagCfg::SLAB_B);
constexpr int DAG_ESLAB = 16 * DG_LD;
constexpr int DAG_E = DAG_NS * DAG_ESLAB;
static_assert(DAG_P >= TB * DG_LD, "resident tile must fit the operand ring");
static_assert(DAG_E >= 64 * DG_WLD + 2 * TB + 64, "diag scratch must fit the epilogue ring");
constexpr int DAG_BAR_OFF = (DAG_P + DAG_E) * (int)sizeof(double);  // 2 x NS mbarriers behind the data
constexpr int DAG_SMEM_BYTES = DAG_BAR_OFF + 2 * DAG_NS * 8;

static_assert(DagCfg::MT == 8 && DagCfg::NT == 4, "the triangular variants below assume 64 x 32 warp tiles");
enum DagTaskType { DAG_DIAG = 0, DAG_POTRF = 1, DAG_TRTRI = 2, DAG_LAUUM = 3 };
constexpr int DAG_NFLAG = 6;  // flag arrays per fit (see DagFit::flags)

// One factorisation problem.  A launch may carry the task lists of MANY independent fits (the lockstep
// batch of mid-size fits, host_batch.cuh): their tasks are interleaved in one list, so a CTA that would
// wait for one fit's chain picks up another fit's tile instead.
struct DagFit {
  double* A;      // Sigma (lower tiles) -> L, in place
  double* X;      // L^-1 (lower tiles; diagonal tiles = D_j)
  double* W;      // Sigma^-1 (lower tiles), written by the LAUUM tasks; before that, scratch for the running
                  // partial sums of TRTRI tiles that are computed in pieces
  size_t ld;
  int nb, n_valid;
  int* flags;     // DAG_NFLAG nb*nb int arrays: L final, X final, pieces done of L tile / X tile / S tile, column strips
                  // stored of the chain tiles L(c+1, c) (zeroed per evaluation)
  double* logdet_part;
  int* info;
};

struct DagArgs {
  const DagFit* fits;
  const int4* tasks;  // {type | fit << 4, i, j, kt0 | kt1 << 10 | piece << 20 | last << 30}: k tiles [kt0, kt1) of the tile
  int ntasks;
  int* counter;       // next task to claim (zeroed per evaluation)
  int strips;         // 1: the chain tile L(c+1, c) is handed to DIAG(c+1) in 32-column strips (see dag_strips_ready)
  int trsm;           // t > 0 (needs strips): the tiles L(c+1 .. c+t, c) are SOLVED against the block columns of L_cc as DIAG(c)
                      // publishes them (dag_trsm_chain) instead of multiplied by D_c = L_cc^-1 once that is complete
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Strip hand-off on the Cholesky chain.  L(c+1, c) = T D_c^T is lower-triangular in D, so its 32-column strip s is
// complete once the epilogue has consumed the D slabs k < 32 (s + 1): the two warps that own the strip store it and
// add 1 to byte s of one packed counter, and DIAG(c+1) - whose last k tile is exactly this tile, as both operands -
// streams k slabs 2 s, 2 s + 1 as soon as byte s reads 2.  The epilogue of the producer, the store of the tile and
// the consumer's last k tile overlap instead of running one after the other (about a sixth of the chain step).
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool dag_strips_ready(const int* p, int shift) { return ((ld_acquire(p) >> shift) & 0xff) >= 2; }
// One lane polls, the warp follows (__syncwarp orders the other lanes' later reads after the acquire).
#ifdef GPEDM_TIMELINE
#define DAG_TL(...) __VA_ARGS__
#else
#define DAG_TL(...)
#endif
#ifdef GPEDM_TIMELINE
// chain marks: extra timeline records {time, time, smid, 150 | mark << 8 | column << 20} logged by thread 0 of the
// two chain tasks of a block column (DIAG(c), POTRF(c+1, c)); tools/dag_timeline.py prints the per-step breakdown
__device__ __forceinline__ void tl_mark_any(unsigned long long t, int mark, int col) {
  const unsigned int i = atomicAdd(&g_tl_count, 1u);
  if (i < TL_MAX) {
    g_tl_rec[4 * i] = t;
    g_tl_rec[4 * i + 1] = t;
    g_tl_rec[4 * i + 2] = tl_smid();
    g_tl_rec[4 * i + 3] = 150ull | ((unsigned long long)mark << 8) | ((unsigned long long)col << 20);
  }
}
__device__ __forceinline__ void tl_mark(unsigned long long t, int mark, int col) {
  if (threadIdx.x == 0) {
    const unsigned int i = atomicAdd(&g_tl_count, 1u);
    if (i < TL_MAX) {
      g_tl_rec[4 * i] = t;
      g_tl_rec[4 * i + 1] = t;
      g_tl_rec[4 * i + 2] = tl_smid();
      g_tl_rec[4 * i + 3] = 150ull | ((unsigned long long)mark << 8) | ((unsigned long long)col << 20);
    }
  }
}
#endif
__device__ __forceinline__ void warp_wait_flag(const int* f, int lane DAG_TL(, unsigned long long& waited)) {
  if (lane == 0) {
    if (ld_acquire(f) == 0) {
      DAG_TL(const unsigned long long w0 = tl_now();)
      while (ld_acquire(f) == 0) __nanosleep(32);
      DAG_TL(waited += tl_now() - w0;)
    }
  }
  __syncwarp();
}

// Ring state shared by the k loop and the epilogue: slabs issued / consumed since the kernel started
// (identical in every thread), so stage and mbarrier parity carry over from task to task.
// The barriers and the stages are addressed by their 32-bit shared-window addresses, formed once per kernel: going
// through generic pointers costs an S2R (CTA rank in the cluster) + address arithmetic + cvta on every barrier wait,
// arrive and copy of every slab hand-off (5.7 % of the warp-stall samples of the round-2 v3 capture).
struct DagRing {
  uint32_t base;  // shared address of the dynamic shared memory; everything else is a constant offset from it
  unsigned gi, gc;
  __device__ __forceinline__ uint32_t full(unsigned s) const { return base + (uint32_t)DAG_BAR_OFF + 8u * s; }
  __device__ __forceinline__ uint32_t empty(unsigned s) const { return base + (uint32_t)DAG_BAR_OFF + 8u * (DAG_NS + s); }
  __device__ __forceinline__ uint32_t sA(unsigned s) const { return base + s * (uint32_t)(DagCfg::SLAB_A * 8); }
  __device__ __forceinline__ uint32_t sB(unsigned s) const { return base + (uint32_t)(DAG_NS * DagCfg::SLAB_A * 8) + s * (uint32_t)(DagCfg::SLAB_B * 8); }
  __device__ __forceinline__ uint32_t sE(unsigned s) const { return base + (uint32_t)(DAG_P * 8) + s * (uint32_t)(DAG_ESLAB * 8); }
};
__device__ __forceinline__ void mbar_init_s(uint32_t a, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t a, unsigned parity) {
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
__device__ __forceinline__ void mbar_arrive_s(uint32_t a) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_cp_async_arrive_s(uint32_t a) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(a) : "memory");
}
// load_slab (dgemm_tile.cuh) with the destination given as a shared address
template <class Cfg, bool KMAJOR, int ROWS>
__device__ __forceinline__ void load_slab_s(uint32_t s, const double* __restrict__ g, size_t ld, int k0, int tid) {
  constexpr int CHUNKS = ROWS * Cfg::BK / 2;  // 16-byte chunks per slab
  static_assert(CHUNKS % Cfg::THREADS == 0, "whole rounds of copies");
  constexpr int ITERS = CHUNKS / Cfg::THREADS;
#pragma unroll
  for (int r = 0; r < ITERS; ++r) {
    const int chunk = tid + r * Cfg::THREADS;
    uint32_t dst;
    const double* src;
    if (!KMAJOR) {  // global: BK columns (k) of ROWS contiguous doubles
      constexpr int CPR = ROWS / 2;
      const int row = chunk / CPR, c2 = (chunk % CPR) << 1;
      dst = s + (uint32_t)(row * (ROWS + 4) + c2) * 8u;
      src = g + (size_t)(k0 + row) * ld + c2;
    } else {  // global: ROWS columns (m) of BK contiguous doubles (k)
      constexpr int CPR = Cfg::BK / 2;
      const int row = chunk / CPR, c2 = (chunk % CPR) << 1;
      dst = s + (uint32_t)(row * Cfg::LD_K + c2) * 8u;
      src = g + (size_t)row * ld + k0 + c2;
    }
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
  }
}

// DIAG(c) side of the chain solve: block column q of L_cc (rows >= 32 q, 32 columns, upper part of the diagonal block
// zeroed) goes to the L tile in global memory as soon as it is final, and `prog` counts the block columns published.
struct DagPublishL {
  const double* T;
  double* Lg;
  size_t ld;
  int* prog;
  bool on;
  int col;
  __device__ __forceinline__ void copy(int q, int first, int nthr) const {
    for (int idx = first; idx < (4 - q) * 512; idx += nthr) {
      const int b = idx >> 9, e = idx & 511;
      const int r = 32 * (q + b) + 2 * (e & 15), c = 32 * q + (e >> 4);
      double2 v = *reinterpret_cast<const double2*>(&T[r + c * DG_LD]);
      if (b == 0) {
        if (r < c) v.x = 0.0;
        if (r + 1 < c) v.y = 0.0;
      }
      *reinterpret_cast<double2*>(&Lg[(size_t)r + (size_t)c * ld]) = v;
    }
  }
  // warps 1 .. 7, while warp 0 factors the next diagonal block: store, meet at a named barrier, one thread flags
  __device__ __forceinline__ void background(int q) const {
    if (!on) return;
    DAG_TL(if (threadIdx.x == 32) tl_mark_any(tl_now(), 40 + q, col);)
    copy(q, (int)threadIdx.x - 32, DAG_THREADS - 32);
    asm volatile("bar.sync 1, %0;" ::"n"(DAG_THREADS - 32) : "memory");
    if (threadIdx.x == 32) {
      st_release(prog, q + 1);
      DAG_TL(tl_mark_any(tl_now(), 44 + q, col);)
    }
  }
  __device__ __forceinline__ void store(int q) const {
    if (!on) return;
    DAG_TL(tl_mark(tl_now(), 40 + q, col);)
    copy(q, (int)threadIdx.x, DAG_THREADS);
  }
  __device__ __forceinline__ void flag(int q) const {  // (after a CTA barrier that follows store(q))
    if (on && threadIdx.x == 0) {
      st_release(prog, q + 1);
      DAG_TL(tl_mark(tl_now(), 44 + q, col);)
    }
  }
};

// POTRF(i, c) side, i = c+1 .. c+t: L(i, c) = T L_cc^-T by blocked forward substitution, one 32-column strip per block
// column of L_cc, each started as soon as DIAG(c) has published that block column - so the tile is complete one 32-wide
// solve after the Cholesky of L_cc instead of after its inverse, the store of D_c and a 128^3 product.  (Same operation
// as LAPACK dtrsm; the tiles further down the column are off the chain and keep the product with D_c.)
//   T   resident tile in shared memory (128 x DG_LD), overwritten with L(c+1, c)
//   Lq  scratch for the staged block column (128 x 32, stride DG_LD) + 32 reciprocals
// Every strip is stored and counted as it completes, for the Cholesky tasks of column c+1 whose last k tile this is
// (dag_strips_ready).
__device__ __forceinline__ void dag_trsm_chain(double* T, double* Lq, const double* Lcc, size_t ld, double* Cg, const int* prog, int* strip, int col) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  double* rdv = Lq + 32 * DG_LD;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    const int base = 32 * q;
    if (lane == 0) {
      while (ld_acquire(prog) <= q) __nanosleep(32);
    }
    __syncwarp();
    DAG_TL(if (col >= 0 && q == 0) tl_mark(tl_now(), 50 + q, col);)
    for (int idx = tid; idx < (4 - q) * 512; idx += DAG_THREADS) {
      const int b = idx >> 9, e = idx & 511;
      const int r = base + 32 * b + 2 * (e & 15), c = e >> 4;
      const double2 v = __ldcg(reinterpret_cast<const double2*>(Lcc + (size_t)r + (size_t)(base + c) * ld));
      *reinterpret_cast<double2*>(&Lq[r + c * DG_LD]) = v;
    }
    __syncthreads();
    if (tid < 32) rdv[tid] = 1.0 / Lq[(base + tid) + tid * DG_LD];
    __syncthreads();
    if (warp < 4) {  // forward substitution, lane = row (as the panel step of diag_cholesky_smem)
      const int row = 32 * warp + lane;
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = T[row + (base + c) * DG_LD];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const double x = a[c] * rdv[c];
        a[c] = x;
        const double* lcol = Lq + base + c * DG_LD;  // column c of the diagonal block, contiguous
#pragma unroll
        for (int c2 = ((c + 1) & ~1); c2 < 32; c2 += 2) {
          const double2 v = *reinterpret_cast<const double2*>(lcol + c2);
          if (c2 > c) a[c2] = fma(-x, v.x, a[c2]);
          a[c2 + 1] = fma(-x, v.y, a[c2 + 1]);
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) T[row + (base + c) * DG_LD] = a[c];
    }
    __syncthreads();
    // strip q is final: to global memory for DIAG(c+1) and everybody else
    for (int idx = tid; idx < 4 * 512; idx += DAG_THREADS) {
      const int r = 2 * (idx & 63), c = base + (idx >> 6);
      *reinterpret_cast<double2*>(&Cg[(size_t)r + (size_t)c * ld]) = *reinterpret_cast<const double2*>(&T[r + c * DG_LD]);
    }
    // columns right of the strip: T[:, n] -= X_q L_cc[n, strip]^T, 32 x 8 pieces (four DMMA chains per warp)
    for (int u = warp; u < 16 * (3 - q); u += DAG_THREADS / 32) {
      const int row0 = 32 * (u & 3), col0 = base + 32 + 8 * (u >> 2);
      double c[4][2];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double* cp = T + (row0 + 8 * r + g) + (col0 + 2 * t) * DG_LD;
        c[r][0] = cp[0];
        c[r][1] = cp[DG_LD];
      }
      const double* ap = T + (row0 + g) + (base + t) * DG_LD;
      const double* bp = Lq + (col0 + g) + t * DG_LD;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        const double b = bp[0];
#pragma unroll
        for (int r = 0; r < 4; ++r) dmma884(c[r][0], c[r][1], -ap[8 * r], b);
        ap += 4 * DG_LD;
        bp += 4 * DG_LD;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double* cp = T + (row0 + 8 * r + g) + (col0 + 2 * t) * DG_LD;
        cp[0] = c[r][0];
        cp[DG_LD] = c[r][1];
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release_add(strip, 2 << (8 * q));
      DAG_TL(if (col >= 0) tl_mark(tl_now(), 58 + q, col);)
    }
  }
}

// The k loop of one task: acc (+)= (+/-) sum over `nkt` 128-wide k tiles of A B, operands streamed through the
// mbarrier ring (see gemm_tile_mb in dgemm_tile.cuh).  A is MN-major; B is MN-major (Cholesky) or K-major
// (inverse).  Operand tile kt is final once fa[kt] / fb[kt * fsb] are set.  Most of them already are when
// the task starts: every warp scans the flags once (32 per load round, acquire) and only the k tiles past
// the first unset flag are polled as the loop reaches them - the ones right behind the chain.
// TRI (diagonal tiles of the Cholesky): the warp tile sits at (m0, n0) of a symmetric tile of which only the lower
// part is used, so the 8 x 8 blocks strictly above the diagonal are skipped.
template <bool A_KMAJOR, bool B_KMAJOR, bool TRI = false, bool STRIPS = false>
__device__ __forceinline__ void dag_kloop(double (&acc)[DagCfg::MT][DagCfg::NT][2], DagRing& ring, double* sA, double* sB,
                                          const double* __restrict__ Ap, const double* __restrict__ Bp, size_t ld,
                                          int nkt, const int* fa, int fsa, const int* fb, int fsb, bool active,
                                          int tid, int m0, int n0, const int* stripa, const int* stripb DAG_TL(, unsigned long long& tl_wait, int tl_col = -1)) {
  using Cfg = DagCfg;
  constexpr int MT = Cfg::MT, NT = Cfg::NT, BK = Cfg::BK, NS = Cfg::NSTAGE, PREF = NS - 2;
  constexpr int SLAB_A = Cfg::SLAB_A, SLAB_B = Cfg::SLAB_B;
  constexpr int LD_A = Cfg::LD_MN_A, LD_B = Cfg::LD_MN_B, LD_K = Cfg::LD_K;
  constexpr int SPT = TB / BK;  // slabs per k tile
  const int lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  // TRI: a DMMA that is predicated off still occupies the sub-partition's DMMA pipe for its 16 cycles (measured),
  // so the skipping is done with warp-uniform BRANCHES between three fully unrolled variants.  The warp tiles of a
  // diagonal task sit at (m0 - n0) / 8 = dl 8 x 8 blocks below the diagonal: dl >= 3 all 32 blocks are needed,
  // dl == 0 the blocks j <= i (26), dl == -4 the blocks j <= i - 4 (10).
  const int tri_cls = !TRI ? 0 : ((m0 - n0) >= 24 ? 0 : ((m0 - n0) == 0 ? 1 : 2));
  int nready = 0;
  for (int base = 0; base < nkt; base += 32) {
    const int kt = base + lane;
    int ok = 1;
    if (kt < nkt) ok = ld_acquire(fa + kt * fsa) & ld_acquire(fb + kt * fsb);
    const unsigned m = __ballot_sync(0xffffffffu, ok != 0);
    const int fz = __ffs(~m) - 1;  // first lane whose flags are not both set; -1 if none
    if (fz >= 0) {
      nready = base + fz;
      break;
    }
    nready = base + 32;
  }
  nready = nready < nkt ? nready : nkt;
  __syncwarp();  // the other lanes' acquires order every lane's later operand reads
  const int nk = nkt * SPT;
  auto issue = [&](int j) {  // copy slab j of this task into ring stage gi % NS
    const unsigned s = ring.gi % NS, r = ring.gi / NS;
    if (r > 0) mbar_wait_s(ring.empty(s), (r - 1) & 1u);
    load_slab_s<Cfg, A_KMAJOR, Cfg::TM>(ring.sA(s), Ap, ld, j * BK, tid);
    load_slab_s<Cfg, B_KMAJOR, Cfg::TN>(ring.sB(s), Bp, ld, j * BK, tid);
    mbar_cp_async_arrive_s(ring.full(s));
    ring.gi++;
  };
  // Slabs are issued PREF ahead of the one being consumed.  The first slab of a k tile that is not yet known to be
  // final needs that tile's flags: while older slabs are still waiting to be consumed the flags are only PEEKED
  // (no blocking: the contraction of what has landed goes first), and the warp blocks on them only once its
  // pipeline is empty - so when the chain tile arrives nothing but that tile's own work is left to do.
  int issued = 0, ready_slabs = nready * SPT;
  for (int kt = 0; kt < nk; ++kt) {
    const int want = kt + PREF + 1 < nk ? kt + PREF + 1 : nk;
    if (want <= ready_slabs) {  // steady state: every slab up to `want` is known to be final, no flag traffic
      while (issued < want) {
        issue(issued);
        ++issued;
      }
    }
    while (issued < want) {
      if (issued >= ready_slabs) {
        const int ktile = issued / SPT;
        if (STRIPS && stripa != nullptr && ktile == nkt - 1) {
          // tiles that are produced strip by strip (two slabs per 32-column strip); same peek / block rule as for whole tiles
          const int sh = ((issued % SPT) >> 1) << 3;
          if (issued > kt) {
            int ok = 0;
            if (lane == 0) ok = dag_strips_ready(stripa, sh) && dag_strips_ready(stripb, sh);
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (!ok) break;
          } else {
            if (lane == 0 && !(dag_strips_ready(stripa, sh) && dag_strips_ready(stripb, sh))) {
              DAG_TL(const unsigned long long w0 = tl_now();)
              while (!(dag_strips_ready(stripa, sh) && dag_strips_ready(stripb, sh))) __nanosleep(32);
              DAG_TL(tl_wait += tl_now() - w0;)
            }
            __syncwarp();
            DAG_TL(if (tl_col >= 0 && (issued % SPT) == 0) tl_mark(tl_now(), 20, tl_col);)
          }
          ready_slabs = (issued & ~1) + 2;
        } else {
          if (issued > kt) {
            int ok = 0;
            if (lane == 0) ok = ld_acquire(fa + ktile * fsa) & ld_acquire(fb + ktile * fsb);
            ok = __shfl_sync(0xffffffffu, ok, 0);  // (also orders the other lanes' operand reads after the acquire)
            if (!ok) break;
          } else {
            warp_wait_flag(fa + ktile * fsa, lane DAG_TL(, tl_wait));
            warp_wait_flag(fb + ktile * fsb, lane DAG_TL(, tl_wait));
            DAG_TL(if (tl_col >= 0 && ktile == nkt - 1) tl_mark(tl_now(), 20, tl_col);)
          }
          ready_slabs = (ktile + 1) * SPT;
        }
      }
      issue(issued);
      ++issued;
    }
    const unsigned s = ring.gc % NS;
    mbar_wait_s(ring.full(s), (ring.gc / NS) & 1u);
    DAG_TL(if (tl_col >= 0 && kt >= nk - SPT) tl_mark(tl_now(), 21 + (kt - (nk - SPT)), tl_col);)
    if (active) {
      const double* a_s = sA + s * SLAB_A;
      const double* b_s = sB + s * SLAB_B;
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; ++k4) {
        double av[MT], bv[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i)
          av[i] = A_KMAJOR ? a_s[(m0 + 8 * i + g) * LD_K + k4 * 4 + t] : a_s[(k4 * 4 + t) * LD_A + m0 + 8 * i + g];
#pragma unroll
        for (int j = 0; j < NT; ++j)
          bv[j] = B_KMAJOR ? b_s[(n0 + 8 * j + g) * LD_K + k4 * 4 + t] : b_s[(k4 * 4 + t) * LD_B + n0 + 8 * j + g];
        if (!TRI || tri_cls == 0) {
#pragma unroll
          for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
        } else if (tri_cls == 1) {
#pragma unroll
          for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
              if (j <= i) dmma884(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
        } else {
#pragma unroll
          for (int i = 4; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
              if (j <= i - 4) dmma884(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_s(ring.empty(s));
    ring.gc++;
  }
}

__global__ void __launch_bounds__(DAG_THREADS, 1) dag_factor_kernel(const DagArgs a) {
  extern __shared__ double smem_raw[];
  // The shared-window address of a CTA carries its rank in the cluster; left to itself the compiler re-derives it
  // (S2R SR_CgaCtaId + arithmetic) at every use under this kernel's register pressure.  One opaque register
  // instead; the generic pointer for ordinary loads / stores is formed from it (and inferred back to shared).
  uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
  asm volatile("" : "+r"(smem_base));
  double* const smem = (double*)__cvta_shared_to_generic(smem_base);
  using Cfg = DagCfg;
  constexpr int MT = Cfg::MT, NT = Cfg::NT, BK = Cfg::BK, NS = Cfg::NSTAGE, PREF = NS - 2;
  constexpr int SLAB_A = Cfg::SLAB_A;
  __shared__ int s_task;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  // Warp tile (64 x 32) origins inside the 128 x 128 tile.  Every SM sub-partition has its own DMMA pipe (one
  // m8n8k4 per 16.75 cycles, tools/smsp_probe.cu) shared by warps w and w + 4, so wherever a triangular operand
  // makes the warp tiles unequal the map pairs heavy with light:
  //   k loops over full tiles        (m0, n0) = (64 (w & 1), 32 (w >> 1))
  //   diagonal tiles (lower part)    the six warp tiles that touch the lower triangle: the two full ones alone on a
  //                                  sub-partition, a 26-block with a 10-block tile on the two others (dm0, dn0)
  //   epilogue  T D_j^T   (POTRF)    work grows with the column strip: strips {0, 96} and {32, 64} paired (n0p)
  //   epilogue -D_i T     (TRTRI)    work grows with the row half: one upper and one lower half per sub-partition (m0t)
  const int m0 = (warp % Cfg::WM) * (MT * 8), n0 = (warp / Cfg::WM) * (NT * 8);
  const int dm0 = (int)((0x8Bu >> warp) & 1u) * 64;                       // w = 0..7 -> 64 64  0 64  0  0  0 64
  const int dn0 = (int)((0x31322010u >> (4 * warp)) & 0xfu) * 32;         //          ->  0 32  0 64 64 96 32 96
  const int n0p = (int)((0x2310u >> (4 * (warp >> 1))) & 0xfu) * 32;      // w >> 1 = 0..3 -> 0 32 96 64
  const int m0t = (int)(((warp & 1) ^ (warp >> 2)) & 1) * 64;
  double* sA = smem;
  double* sB = smem + NS * SLAB_A;
  double* T = smem;               // resident tile, aliases the operand ring
  double* sE = smem + DAG_P;      // epilogue ring / diag scratch
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init_s(smem_base + (uint32_t)DAG_BAR_OFF + 8u * s, DAG_THREADS);
      mbar_init_s(smem_base + (uint32_t)DAG_BAR_OFF + 8u * (NS + s), DAG_THREADS / 32);
    }
  }
  __syncthreads();
  DagRing ring{smem_base, 0u, 0u};
  int* counter = a.counter;

  for (;;) {
    if (tid == 0) s_task = atomicAdd(counter, 1);
    __syncthreads();
    const int tix = s_task;
    if (tix >= a.ntasks) break;
    const int4 tk = a.tasks[tix];
    const int type = tk.x & 15, ti = tk.y, tj = tk.z;
    const DagFit f = a.fits[tk.x >> 4];
    const int nb = f.nb;
    const size_t ld = f.ld;
    int* flagL = f.flags;
    int* flagX = flagL + nb * nb;
    int* progL = flagX + nb * nb;
    int* progX = progL + nb * nb;
    int* stripL = progX + 2 * nb * nb;  // (behind the S-tile piece counts)
    const int kt0 = tk.w & 1023, kt1 = (tk.w >> 10) & 1023, piece = (tk.w >> 20) & 1023;
    const bool last = (tk.w >> 30) & 1;
    DAG_TL(const unsigned long long tl_t0 = tl_now(); unsigned long long tl_wait = 0;)
    // phase marks of every 4th task (30 k loop entered, 31 k loop left, 32 resident tile in smem, 33 epilogue product done)
    DAG_TL(const bool tl_ph = (tix & 3) == 0;)

    // ---------------- operands of the k loop
    const double* Ap;
    const double* Bp;
    const int *fa, *fb;
    int fsa = 1, fsb, nkt;  // flag strides of the operands, number of 128-wide k tiles of this piece
    double* Cg;    // output tile in global memory
    const bool is_trtri = type == DAG_TRTRI;
    if (type == DAG_LAUUM) {
      // ---------------- S_ij = sum_{k >= i} X_ki^T X_kj: both operands K-major, rows of X become final in
      // ascending order, so the k loop follows the inverse as it completes; no epilogue, no ready flag (nothing
      // in this kernel reads S).  A tile may be summed in PIECES over block rows [ti + kt0, ti + kt1) of X
      // (partial sums kept in place in S; summation order unchanged), so that the chain-bound stretches of a
      // mid-size factorisation absorb the X^T X work instead of leaving it all for after the last column.
      const int r0 = ti + kt0;
      const double* LAp = f.X + (size_t)r0 * TB + (size_t)ti * TB * ld;
      const double* LBp = f.X + (size_t)r0 * TB + (size_t)tj * TB * ld;
      double* Sg = f.W + (size_t)ti * TB + (size_t)tj * TB * ld;
      int* progS = progX + nb * nb + ti * nb + tj;
      const bool act = !(ti == tj && n0 > m0 + MT * 8 - 1);
      if (piece > 0) {
        if (lane == 0) {
          if (ld_acquire(progS) < piece) {
            DAG_TL(const unsigned long long w0 = tl_now();)
            while (ld_acquire(progS) < piece) __nanosleep(32);
            DAG_TL(tl_wait += tl_now() - w0;)
          }
        }
        __syncwarp();
      }
      double acc[MT][NT][2];
      if (piece > 0 && act) {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) acc[i][j][e] = __ldcg(Sg + (size_t)(m0 + 8 * i + g) + (size_t)(n0 + 8 * j + 2 * t + e) * ld);
      } else {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      }
      DAG_TL(if (tl_ph) tl_mark(tl_now(), 30, type);)
      dag_kloop<true, true>(acc, ring, sA, sB, LAp, LBp, ld, kt1 - kt0, flagX + r0 * nb + ti, nb, flagX + r0 * nb + tj, nb, act,
                            tid, m0, n0, nullptr, nullptr DAG_TL(, tl_wait));
      DAG_TL(if (tl_ph) tl_mark(tl_now(), 31, type);)
      if (act) {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) Sg[(size_t)(m0 + 8 * i + g) + (size_t)(n0 + 8 * j + 2 * t + e) * ld] = acc[i][j][e];
      }
      if (!last) {
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          st_release(progS, piece + 1);
        }
      }
      DAG_TL(tl_log(tl_t0, (200 + type) | ((long long)ti << 8) | ((long long)tj << 20) | ((long long)(kt1 - kt0) << 32) | ((long long)(tl_wait > 0xfffff ? 0xfffff : tl_wait) << 44));)
      continue;
    }
    const double* Cin;  // where the accumulator starts from (nullptr = zero)
    double* Cpart;      // where a non-final piece leaves its partial result
    int* prog;
    if (!is_trtri) {
      Ap = f.A + (size_t)ti * TB + (size_t)kt0 * TB * ld;
      Bp = f.A + (size_t)tj * TB + (size_t)kt0 * TB * ld;
      fa = flagL + ti * nb + kt0;
      fb = flagL + tj * nb + kt0;
      fsb = 1;
      Cg = f.A + (size_t)ti * TB + (size_t)tj * TB * ld;
      Cin = Cg;  // in place: Sigma tile, then Sigma - partial sums
      Cpart = Cg;
      prog = progL + ti * nb + tj;
    } else {
      Ap = f.A + (size_t)ti * TB + (size_t)(tj + kt0) * TB * ld;
      Bp = f.X + (size_t)(tj + kt0) * TB + (size_t)tj * TB * ld;
      fa = flagL + ti * nb + tj + kt0;
      fb = flagX + (tj + kt0) * nb + tj;
      fsb = nb;
      Cg = f.X + (size_t)ti * TB + (size_t)tj * TB * ld;
      Cpart = f.W + (size_t)ti * TB + (size_t)tj * TB * ld;
      Cin = piece > 0 ? Cpart : nullptr;
      prog = progX + ti * nb + tj;
    }
    nkt = kt1 - kt0;
    if (piece > 0) {  // the previous piece of this tile must have left its partial result
      if (lane == 0) {
        if (ld_acquire(prog) < piece) {
          DAG_TL(const unsigned long long w0 = tl_now();)
          while (ld_acquire(prog) < piece) __nanosleep(32);
          DAG_TL(tl_wait += tl_now() - w0;)
        }
      }
      __syncwarp();
    }
    // diagonal task: balanced warp tile map (see the kernel head); warps 4 and 5 hold the two tiles above the diagonal
    const bool active = type != DAG_DIAG || (warp != 4 && warp != 5);
    const int km0 = type == DAG_DIAG ? dm0 : m0, kn0 = type == DAG_DIAG ? dn0 : n0;

    double acc[MT][NT][2];
    if (Cin != nullptr) {  // __ldcg: a partial result may have been written by another SM
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            // Cholesky tasks accumulate the NEGATED tile, -(A - sum L L^T) = -A + sum L L^T, so that the k loop needs
            // no sign flips on its fragments; negation commutes with rounding, so the bits are those of A - sum L L^T
            const double v = __ldcg(Cin + (size_t)(km0 + 8 * i + g) + (size_t)(kn0 + 8 * j + 2 * t + e) * ld);
            acc[i][j][e] = is_trtri ? v : -v;
          }
    } else {
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }
    DAG_TL(if (tl_ph) tl_mark(tl_now(), 30, type);)
    // The last k tile (block column tj - 1) of a Cholesky task is taken strip by strip when its operand tiles are
    // produced that way: the chain tile L(tj, tj - 1) with the strips on, L(ti, tj - 1) if it is solved (dag_trsm_chain).
    const bool last_col = last && kt1 == tj && nkt > 0;
    if (type == DAG_DIAG) {
      const int* strip = (a.strips && last_col) ? stripL + tj * nb + (tj - 1) : nullptr;
      dag_kloop<false, false, true, true>(acc, ring, sA, sB, Ap, Bp, ld, nkt, fa, fsa, fb, fsb, active, tid, km0, kn0, strip, strip DAG_TL(, tl_wait, last ? tj : -1));
    } else if (!is_trtri) {
      // (two instances: the strip logic stays out of the loop that all the other Cholesky tiles run)
      if (last_col && ti - (tj - 1) <= a.trsm)
        dag_kloop<false, false, false, true>(acc, ring, sA, sB, Ap, Bp, ld, nkt, fa, fsa, fb, fsb, true, tid, km0, kn0,
                                             stripL + ti * nb + (tj - 1), stripL + tj * nb + (tj - 1) DAG_TL(, tl_wait));
      else
        dag_kloop<false, false>(acc, ring, sA, sB, Ap, Bp, ld, nkt, fa, fsa, fb, fsb, true, tid, km0, kn0, nullptr, nullptr DAG_TL(, tl_wait));
    } else
      dag_kloop<false, true>(acc, ring, sA, sB, Ap, Bp, ld, nkt, fa, fsa, fb, fsb, true, tid, km0, kn0, nullptr, nullptr DAG_TL(, tl_wait));
    DAG_TL(if (tl_ph) tl_mark(tl_now(), 31, type);)
    if (!is_trtri) {  // back from the negated accumulation
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          acc[i][j][0] = -acc[i][j][0];
          acc[i][j][1] = -acc[i][j][1];
        }
    }
    if (!last) {
      // ---------------- non-final piece: leave the partial result, publish the piece count
      if (active) {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) Cpart[(size_t)(km0 + 8 * i + g) + (size_t)(kn0 + 8 * j + 2 * t + e) * ld] = acc[i][j][e];
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(prog, piece + 1);
      }
      DAG_TL(tl_log(tl_t0, (210 + type) | ((long long)ti << 8) | ((long long)tj << 20) | ((long long)nkt << 32) | ((long long)(tl_wait > 0xfffff ? 0xfffff : tl_wait) << 44));)
      continue;
    }
    DAG_TL(if (type == DAG_DIAG || ti == tj + 1) tl_mark(tl_now(), type == DAG_DIAG ? 0 : 10, tj);)
    __syncthreads();  // every warp is done with the operand ring: T may overwrite it
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) T[(km0 + 8 * i + g) + (kn0 + 8 * j + 2 * t + e) * DG_LD] = acc[i][j][e];
    __syncthreads();
    DAG_TL(if (tl_ph) tl_mark(tl_now(), 32, type);)

    if (type == DAG_DIAG) {
      // ---------------- in-CTA Cholesky + inverse of the diagonal block (diag_block.cuh)
      double* w = sE;
      double* rdiag = w + 64 * DG_WLD;
      double* dvec = rdiag + TB;
      double* colbuf = dvec + TB;
      DAG_TL(tl_mark(tl_now(), 1, tj);)
      const bool trsm = a.trsm != 0;  // the next panel tile is solved against the block columns of L_jj as they complete
      const DagPublishL pub{T, Cg, ld, stripL + tj * nb + tj, trsm, tj};
      diag_cholesky_smem<DAG_THREADS, DagPublishL>(T, rdiag, dvec, colbuf, tj, f.n_valid, f.info, pub);
      DAG_TL(tl_mark(tl_now(), 2, tj);)
      auto lower_block_elem = [](int idx, int& r, int& c, bool& on_diag) {
        const int b = idx >> 9, e = idx & 511;
        const int bi = (b >= 6) ? 3 : ((b >= 3) ? 2 : ((b >= 1) ? 1 : 0)), bj = b - bi * (bi + 1) / 2;
        r = 32 * bi + 2 * (e & 15);
        c = 32 * bj + (e >> 4);
        on_diag = bi == bj;
      };
      // The chain waits for D_j = L_jj^-1 only (the next column's panel is T D_j^T; nothing in this kernel reads the
      // diagonal tile of L): L_jj is parked in registers, the in-place inverse and the store of D_j go first and
      // the flag is released; L_jj and the log-determinant part follow off the critical path.
      constexpr int LPT = 10 * 512 / DAG_THREADS;  // double2 per thread: the 10 lower 32 x 32 blocks
      __syncthreads();
      pub.flag(3);  // (the last block column was stored before this barrier)
      // (L_jj is parked in either mode: making this conditional moves the array to local memory and costs the k loops
      // of every task type registers - measured in the SASS, 576 -> 613 instructions per slab in the X^T X loop)
      double2 lreg[LPT];
#pragma unroll
      for (int q = 0; q < LPT; ++q) {
        int r, c;
        bool dg;
        lower_block_elem(tid + q * DAG_THREADS, r, c, dg);
        double2 v = *reinterpret_cast<const double2*>(&T[r + c * DG_LD]);
        if (dg) {
          if (r < c) v.x = 0.0;
          if (r + 1 < c) v.y = 0.0;
        }
        lreg[q] = v;
      }
      double lv = 0.0;
      if (tid < TB && tj * TB + tid < f.n_valid) lv = log(dvec[tid]);
      __syncthreads();
      DAG_TL(tl_mark(tl_now(), 3, tj);)
      diag_invert_smem<DAG_THREADS>(T, w, rdiag);
      DAG_TL(tl_mark(tl_now(), 4, tj);)
      double* Dg = f.X + (size_t)tj * TB * (ld + 1);
#pragma unroll 4
      for (int idx = tid; idx < 10 * 512; idx += DAG_THREADS) {
        int r, c;
        bool dg;
        lower_block_elem(idx, r, c, dg);
        double2 v = *reinterpret_cast<const double2*>(&T[r + c * DG_LD]);
        if (dg) {
          if (r < c) v.x = 0.0;
          if (r + 1 < c) v.y = 0.0;
        }
        *reinterpret_cast<double2*>(&Dg[(size_t)r + (size_t)c * ld]) = v;
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(flagX + tj * nb + tj, 1);
      }
      DAG_TL(tl_mark(tl_now(), 5, tj);)
      if (!trsm) {  // (published block column by block column otherwise)
#pragma unroll
        for (int q = 0; q < LPT; ++q) {
          int r, c;
          bool dg;
          lower_block_elem(tid + q * DAG_THREADS, r, c, dg);
          *reinterpret_cast<double2*>(&Cg[(size_t)r + (size_t)c * ld]) = lreg[q];
        }
      }
      if (tid < TB) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lv += __shfl_xor_sync(0xffffffffu, lv, off);
        if ((tid & 31) == 0) dvec[tid >> 5] = lv;
      }
      __syncthreads();
      if (tid == 0) {
        f.logdet_part[tj] = ((dvec[0] + dvec[1]) + dvec[2]) + dvec[3];
        __threadfence();
        st_release(flagL + tj * nb + tj, 1);
      }
      // timeline code: 200 + type | i << 8 | j << 20 | k tiles << 32 | ns spent waiting for flags (warp 0) << 44
      DAG_TL(tl_log(tl_t0, (200 + type) | ((long long)ti << 8) | ((long long)tj << 20) | ((long long)nkt << 32) | ((long long)(tl_wait > 0xfffff ? 0xfffff : tl_wait) << 44));)
      continue;
    }

    if (!is_trtri && ti - tj <= a.trsm) {
      // ---------------- chain tile (and the tiles right below it, which the next column's chain tile needs first):
      // solved against L_jj block column by block column (dag_trsm_chain), handed on strip by strip
      dag_trsm_chain(T, sE, f.A + (size_t)tj * TB * (ld + 1), ld, Cg, stripL + tj * nb + tj, stripL + ti * nb + tj, ti == tj + 1 ? tj : -1);
      if (tid == 0) st_release(flagL + ti * nb + tj, 1);  // (behind the barrier and fence of the last strip)
      DAG_TL(tl_mark(tl_now(), 13, tj);)
      DAG_TL(tl_log(tl_t0, (200 + type) | ((long long)ti << 8) | ((long long)tj << 20) | ((long long)nkt << 32) | ((long long)(tl_wait > 0xfffff ? 0xfffff : tl_wait) << 44));)
      continue;
    }
    // ---------------- epilogue product against the diagonal-block inverse
    //   POTRF: C = T D_j^T   (A operand = T resident, B operand = D_j streamed, B[k][n] = D_j[n][k])
    //   TRTRI: C = -D_i T    (A operand = D_i streamed,  B operand = T resident)
    // D is lower triangular: DMMAs whose D fragment is identically zero are skipped.
    const int em0 = is_trtri ? m0t : m0, en0 = is_trtri ? n0 : n0p;  // balanced warp tile maps of the epilogue
    const int dblk = is_trtri ? ti : tj;
    const double* Dg = f.X + (size_t)dblk * TB * (ld + 1);
    // chain tile: every warp stores its 32-column strip as soon as the D slabs that reach it are consumed (dag_strips_ready)
    const bool chain_strips = a.strips && !is_trtri && ti == tj + 1;
    warp_wait_flag(flagX + dblk * nb + dblk, lane DAG_TL(, tl_wait));
    DAG_TL(if (!is_trtri && ti == tj + 1) tl_mark(tl_now(), 11, tj);)
    auto issue_e = [&](int j) {
      const unsigned s = ring.gi % NS, r = ring.gi / NS;
      if (r > 0) mbar_wait_s(ring.empty(s), (r - 1) & 1u);
      load_slab_s<Cfg, false, TB>(ring.sE(s), Dg, ld, j * BK, tid);
      mbar_cp_async_arrive_s(ring.full(s));
      ring.gi++;
    };
    constexpr int NKE = TB / BK;
#pragma unroll
    for (int j = 0; j < PREF; ++j) issue_e(j);
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < NKE; ++kt) {
      if (kt + PREF < NKE) issue_e(kt + PREF);
      const unsigned s = ring.gc % NS;
      mbar_wait_s(ring.full(s), (ring.gc / NS) & 1u);
      const double* e_s = sE + s * DAG_ESLAB;
      // D is lower triangular: D[r][k] = 0 for k > r.  The DMMAs whose D fragment is identically zero are skipped
      // with warp-uniform jumps into a fall-through switch (a predicated-off DMMA would still cost its pipe slot).
#define DAG_ROW(i_)                                                                              \
  _Pragma("unroll") for (int j = 0; j < NT; ++j) dmma884(acc[i_][j][0], acc[i_][j][1], av[i_], bv[j]);
#define DAG_COL(j_)                                                                              \
  _Pragma("unroll") for (int i = 0; i < MT; ++i) dmma884(acc[i][j_][0], acc[i][j_][1], av[i], bv[j_]);
#pragma unroll 1
      for (int k4 = 0; k4 < BK / 4; ++k4) {
        const int kk = kt * BK + k4 * 4;  // first k of this DMMA step
        double av[MT], bv[NT];
        if (is_trtri) {
          const int imin = kk - em0 > 0 ? (kk - em0) >> 3 : 0;  // row blocks i < imin: 8 i + 7 + em0 < kk
          if (imin >= MT) continue;
#pragma unroll
          for (int i = 0; i < MT; ++i) av[i] = e_s[(k4 * 4 + t) * DG_LD + em0 + 8 * i + g];
#pragma unroll
          for (int j = 0; j < NT; ++j) bv[j] = T[(en0 + 8 * j + g) * DG_LD + kk + t];
          switch (imin) {
            case 0: DAG_ROW(0)  // fall through
            case 1: DAG_ROW(1)
            case 2: DAG_ROW(2)
            case 3: DAG_ROW(3)
            case 4: DAG_ROW(4)
            case 5: DAG_ROW(5)
            case 6: DAG_ROW(6)
            default: DAG_ROW(7)
          }
        } else {
          const int jmin = kk - en0 > 0 ? (kk - en0) >> 3 : 0;  // column blocks j < jmin: 8 j + 7 + en0 < kk
          if (jmin >= NT) continue;
#pragma unroll
          for (int i = 0; i < MT; ++i) av[i] = T[(kk + t) * DG_LD + em0 + 8 * i + g];
#pragma unroll
          for (int j = 0; j < NT; ++j) bv[j] = e_s[(k4 * 4 + t) * DG_LD + en0 + 8 * j + g];
          switch (jmin) {
            case 0: DAG_COL(0)  // fall through
            case 1: DAG_COL(1)
            case 2: DAG_COL(2)
            default: DAG_COL(3)
          }
        }
      }
#undef DAG_ROW
#undef DAG_COL
      __syncwarp();
      if (lane == 0) mbar_arrive_s(ring.empty(s));
      ring.gc++;
      if (chain_strips && kt == (en0 >> 4) + 1) {  // columns [en0, en0 + 32) take D slabs k < en0 + 32 only
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) Cg[(size_t)(em0 + 8 * i + g) + (size_t)(en0 + 8 * j + 2 * t + e) * ld] = acc[i][j][e];
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          red_release_add(stripL + ti * nb + tj, 1 << ((en0 >> 5) << 3));
        }
      }
    }
    DAG_TL(if (!is_trtri && ti == tj + 1) tl_mark(tl_now(), 12, tj);)
    DAG_TL(if (tl_ph) tl_mark(tl_now(), 33, type);)
    if (!chain_strips) {
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e)  // TRTRI: X_ij = -(D_i T), the sign applied here instead of on every D fragment
            Cg[(size_t)(em0 + 8 * i + g) + (size_t)(en0 + 8 * j + 2 * t + e) * ld] = is_trtri ? -acc[i][j][e] : acc[i][j][e];
    }
    __syncthreads();  // all stores of the tile issued; T is free again for the next task's ring
    if (tid == 0) {
      __threadfence();
      st_release((is_trtri ? flagX : flagL) + ti * nb + tj, 1);
    }
    DAG_TL(if (!is_trtri && ti == tj + 1) tl_mark(tl_now(), 13, tj);)
    DAG_TL(tl_log(tl_t0, (200 + type) | ((long long)ti << 8) | ((long long)tj << 20) | ((long long)nkt << 32) | ((long long)(tl_wait > 0xfffff ? 0xfffff : tl_wait) << 44));)
  }
}

}  // namespace gpedm
