// CUDA kernels of the GP evaluation pipeline (sm_100a, FP64).
// Matrices are Np x Np column-major with Np = N rounded up to a multiple of 128; the padded
// rows/columns carry an identity so they are inert in every factorisation step.
//   bufA : Sigma (lower, full diagonal tiles)  -> L   (in-place Cholesky)
//   bufB : X = L^-1 (lower; diagonal tiles carry explicit zeros above the diagonal)
//   bufC : scratch W during the triangular inverse, then S = Sigma^-1 (lower)
#pragma once
#include "dgemm_tile.cuh"
#include "tile_stream.cuh"

namespace gpedm {

constexpr int MAXD = 32;

// Per-evaluation hyper-parameters, resident in device memory so that every kernel reads the
// current values (keeps launches replayable).
struct EvalParams {
  double sqphi[MAXD];
  double phi[MAXD];
  double ve, sigma2, rho;
  int d, np, use_rhotab, pad;
};

// -----------------------------------------------------------------------------------------
// Covariance entry, arithmetic order of getcov (reference R/GPfunctions.R:787-811):
//   lC0 = -sum_k (sqrt(phi_k) x2_ik - sqrt(phi_k) x_jk)^2   (laGP::distance: difference, square,
//         accumulate in coordinate order; no fused multiply-add so rounding matches a CPU)
//   Cd  = (sigma2 * exp(lC0)) * R,  R = 1 (same pop) | rho | rhomat[pop2_i, pop_j]
// xi / xj point at pre-scaled coordinates with stride `sx` between coordinates.
__device__ __forceinline__ double cov_entry(const double* xi, const double* xj, int sx, int d, double sigma2,
                                            double rho, int pi, int pj, const double* rhotab, int np) {
  double dist = 0.0;
  for (int k = 0; k < d; ++k) {
    double diff = __dsub_rn(xi[k * sx], xj[k * sx]);
    dist = __dadd_rn(dist, __dmul_rn(diff, diff));
  }
  double v = __dmul_rn(sigma2, exp(-dist));
  if (rhotab != nullptr) {
    // a new point of a population absent from the training set carries code np: the reference's Rmat
    // loop (:799-808) only visits training populations, so its entries keep the initial 1
    if (pi < np) v = __dmul_rn(v, rhotab[pi + pj * np]);
  } else if (pi != pj) {
    v = __dmul_rn(v, rho);  // code np never equals a training code: popsame = 0, the rho branch (:811)
  }
  return v;
}

// Sigma = Cd + ve*I on the lower tiles (reference R/GPfunctions.R:633-641; the symmetrisation
// at :641 is a no-op because each entry is computed once).  grid = nb*(nb+1)/2 tiles x csplit column
// ranges (csplit > 1 only when there are fewer tiles than SMs).  A thread owns TWO consecutive rows
// (16-byte loads of the staged predictors, 16-byte stores: one column of the tile is written as 64 x 16
// contiguous bytes) and every fourth column of its range; the arithmetic per entry is exactly cov_entry's.
// `Akeep` receives the same values: Sigma is factored in place, and the gradient trace terms read Cd = Sigma - ve I
// from this copy instead of recomputing every distance and exponential (the assembly is bound by FP64 issue, not by
// its stores, so the second stream of stores is free).
__device__ __forceinline__ void assemble_sigma_body(double* __restrict__ A, double* __restrict__ Akeep, size_t ld, int N,
                                                              const double* __restrict__ X,
                                                              const int* __restrict__ pop,
                                                              const double* __restrict__ rhotab,
                                                              const EvalParams* __restrict__ P, int csplit, const int bid) {
  extern __shared__ double sm[];
  const int d = P->d;
  double* xi = sm;             // [d][128]
  double* xj = sm + d * TB;    // [d][128]
  int* pi_s = (int*)(xj + d * TB);
  int* pj_s = pi_s + TB;
  int ti, tj;
  tri_decode(bid / csplit, ti, tj);
  const int ncols = TB / csplit, cbeg = (bid % csplit) * ncols;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < d * TB; idx += 256) {
    int k = idx >> 7, r = idx & 127;
    int gi = ti * TB + r, gj = tj * TB + r;
    double s = P->sqphi[k];
    xi[idx] = gi < N ? __dmul_rn(X[(size_t)gi + (size_t)k * N], s) : 0.0;
    xj[idx] = gj < N ? __dmul_rn(X[(size_t)gj + (size_t)k * N], s) : 0.0;
  }
  if (tid < TB) {
    int gi = ti * TB + tid, gj = tj * TB + tid;
    pi_s[tid] = gi < N ? pop[gi] : -1;
    pj_s[tid] = gj < N ? pop[gj] : -1;
  }
  __syncthreads();
  const double sigma2 = P->sigma2, rho = P->rho, ve = P->ve;
  const int np = P->np;
  const double* rt = P->use_rhotab ? rhotab : nullptr;
  const int r0 = 2 * (tid & 63), cg = tid >> 6;
  const int gi0 = ti * TB + r0;
  const int p0 = max(pi_s[r0], 0), p1 = max(pi_s[r0 + 1], 0);  // (padding rows: value overwritten below)
  for (int c = cbeg + cg; c < cbeg + ncols; c += 4) {
    const int gj = tj * TB + c;
    double dist0 = 0.0, dist1 = 0.0;
    for (int k = 0; k < d; ++k) {
      const double2 xr = *reinterpret_cast<const double2*>(xi + k * TB + r0);
      const double xc = xj[k * TB + c];
      const double f0 = __dsub_rn(xr.x, xc), f1 = __dsub_rn(xr.y, xc);
      dist0 = __dadd_rn(dist0, __dmul_rn(f0, f0));
      dist1 = __dadd_rn(dist1, __dmul_rn(f1, f1));
    }
    const int pj = max(pj_s[c], 0);
    double v0 = __dmul_rn(sigma2, exp(-dist0)), v1 = __dmul_rn(sigma2, exp(-dist1));
    if (rt != nullptr) {
      v0 = __dmul_rn(v0, rt[p0 + pj * np]);  // training codes are always < np
      v1 = __dmul_rn(v1, rt[p1 + pj * np]);
    } else {
      if (p0 != pj) v0 = __dmul_rn(v0, rho);
      if (p1 != pj) v1 = __dmul_rn(v1, rho);
    }
    if (gi0 == gj) v0 = __dadd_rn(v0, ve);
    if (gi0 + 1 == gj) v1 = __dadd_rn(v1, ve);
    if (gj >= N || gi0 >= N) v0 = (gi0 == gj) ? 1.0 : 0.0;          // identity in the padding
    if (gj >= N || gi0 + 1 >= N) v1 = (gi0 + 1 == gj) ? 1.0 : 0.0;
    *reinterpret_cast<double2*>(A + (size_t)gi0 + (size_t)gj * ld) = make_double2(v0, v1);
    *reinterpret_cast<double2*>(Akeep + (size_t)gi0 + (size_t)gj * ld) = make_double2(v0, v1);
  }
}
__global__ void __launch_bounds__(256) assemble_sigma_kernel(double* __restrict__ A, double* __restrict__ Akeep, size_t ld, int N,
                                                              const double* __restrict__ X,
                                                              const int* __restrict__ pop,
                                                              const double* __restrict__ rhotab,
                                                              const EvalParams* __restrict__ P, int csplit) {
  assemble_sigma_body(A, Akeep, ld, N, X, pop, rhotab, P, csplit, (int)blockIdx.x);
}

// General rectangular covariance (getcov with X2 != X): out[(row) + (col)*ldo] where
// transpose == 0: rows = new points (M), cols = training points (N)   -> Cd as R returns it
// transpose == 1: rows = training points, cols = new points            -> Cs^T for the solves
// Entries outside M x N (padding up to the allocated extents rows_alloc x cols_alloc) are zero.
__global__ void __launch_bounds__(256) cov_rect_kernel(double* __restrict__ out, size_t ldo, int rows_alloc,
                                                        int cols_alloc, int transpose, int N,
                                                        const double* __restrict__ X, const int* __restrict__ pop,
                                                        int M, const double* __restrict__ X2,
                                                        const int* __restrict__ pop2,
                                                        const double* __restrict__ rhotab,
                                                        const EvalParams* __restrict__ P, int add_ve_diag) {
  extern __shared__ double sm[];
  const int d = P->d;
  double* xr = sm;
  double* xc = sm + d * TB;
  int* pr_s = (int*)(xc + d * TB);
  int* pc_s = pr_s + TB;
  const int tr = blockIdx.x, tc = blockIdx.y;
  const int tid = threadIdx.x;
  // row-side / col-side sources
  const double* Xr = transpose ? X : X2;
  const double* Xc = transpose ? X2 : X;
  const int* popr = transpose ? pop : pop2;
  const int* popc = transpose ? pop2 : pop;
  const int nr = transpose ? N : M, nc = transpose ? M : N;
  for (int idx = tid; idx < d * TB; idx += 256) {
    int k = idx >> 7, r = idx & 127;
    int gr = tr * TB + r, gc = tc * TB + r;
    double s = P->sqphi[k];
    xr[idx] = gr < nr ? __dmul_rn(Xr[(size_t)gr + (size_t)k * nr], s) : 0.0;
    xc[idx] = gc < nc ? __dmul_rn(Xc[(size_t)gc + (size_t)k * nc], s) : 0.0;
  }
  if (tid < TB) {
    int gr = tr * TB + tid, gc = tc * TB + tid;
    pr_s[tid] = gr < nr ? popr[gr] : -1;
    pc_s[tid] = gc < nc ? popc[gc] : -1;
  }
  __syncthreads();
  const double sigma2 = P->sigma2, rho = P->rho, ve = P->ve;
  const int np = P->np;
  const double* rt = P->use_rhotab ? rhotab : nullptr;
  const int r = tid & 127, half = tid >> 7;
  const int gr = tr * TB + r;
  if (gr >= rows_alloc) return;
  for (int c = half * 64; c < half * 64 + 64; ++c) {
    int gc = tc * TB + c;
    if (gc >= cols_alloc) break;
    double v = 0.0;
    if (gr < nr && gc < nc) {
      // rhomat is addressed [pop2 (new), pop (train)] (reference :806)
      int p_new = transpose ? pc_s[c] : pr_s[r];
      int p_trn = transpose ? pr_s[r] : pc_s[c];
      // the squared distance is symmetric in its arguments term by term, so xr/xc order is free
      v = cov_entry(xr + r, xc + c, TB, d, sigma2, rho, p_new, p_trn, rt, np);
      if (add_ve_diag && gr == gc) v = __dadd_rn(v, ve);
    }
    out[(size_t)gr + (size_t)gc * ldo] = v;
  }
}

// getcovinv on a caller-supplied matrix (reference R/GPfunctions.R:818-830): load the symmetrised lower
// tiles of Sigma (N x N column-major, "(Sigma + t(Sigma))/2" as the reference does before calling it,
// :641, :1076) into the padded work matrix, identity in the padding.
__global__ void __launch_bounds__(256) load_sigma_kernel(double* __restrict__ A, double* __restrict__ Akeep, size_t ld, int N,
                                                          const double* __restrict__ S) {
  int ti, tj;
  tri_decode(blockIdx.x, ti, tj);
  const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int gi = ti * TB + r;
  for (int c = half * 64; c < half * 64 + 64; ++c) {
    const int gj = tj * TB + c;
    double v;
    if (gi < N && gj < N)
      v = 0.5 * (S[(size_t)gi + (size_t)gj * N] + S[(size_t)gj + (size_t)gi * N]);
    else
      v = (gi == gj) ? 1.0 : 0.0;
    A[(size_t)gi + (size_t)gj * ld] = v;
    Akeep[(size_t)gi + (size_t)gj * ld] = v;
  }
}

// popsame (reference :796) for export.
__global__ void popsame_kernel(double* out, int N, const int* pop) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * N) return;
  int i = (int)(idx % N), j = (int)(idx / N);
  out[idx] = pop[i] == pop[j] ? 1.0 : 0.0;
}

// -----------------------------------------------------------------------------------------
// Diagonal block: Cholesky of the 128x128 tile A_kk (lower, in place) and its inverse
// D_k = L_kk^-1 written as a full tile (zeros above the diagonal) into bufB's diagonal tile.
// Also emits sum(log L_ii) of this block and flags the first non-positive pivot.
// One CTA; the tile lives in shared memory with stride 129.
constexpr int DIAG_LD = TB + 1;
constexpr int DIAG_SMEM_BYTES = (TB * DIAG_LD + 2 * TB) * (int)sizeof(double);

__global__ void __launch_bounds__(256) diag_potf2_inv_kernel(double* __restrict__ Akk, size_t lda,
                                                              double* __restrict__ Dkk, size_t ldd, int kblock,
                                                              int n_valid, double* __restrict__ logdet_part,
                                                              int* __restrict__ info) {
  extern __shared__ double sm[];
  double* s = sm;                     // [128][129] column-major: s[r + c*129]
  double* dvec = sm + TB * DIAG_LD;   // pivots
  double* xv = dvec + TB;             // column scratch for the inverse
  const int tid = threadIdx.x;
  for (int idx = tid; idx < TB * TB; idx += 256) {
    int r = idx & 127, c = idx >> 7;
    s[r + c * DIAG_LD] = Akk[(size_t)r + (size_t)c * lda];
  }
  __syncthreads();
  // ---- left-looking Cholesky, one column per step ----
  for (int j = 0; j < TB; ++j) {
    double sum = 0.0;
    if (tid >= j && tid < TB) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int c = 0;
      for (; c + 3 < j; c += 4) {
        s0 = fma(s[tid + c * DIAG_LD], s[j + c * DIAG_LD], s0);
        s1 = fma(s[tid + (c + 1) * DIAG_LD], s[j + (c + 1) * DIAG_LD], s1);
        s2 = fma(s[tid + (c + 2) * DIAG_LD], s[j + (c + 2) * DIAG_LD], s2);
        s3 = fma(s[tid + (c + 3) * DIAG_LD], s[j + (c + 3) * DIAG_LD], s3);
      }
      for (; c < j; ++c) s0 = fma(s[tid + c * DIAG_LD], s[j + c * DIAG_LD], s0);
      sum = s[tid + j * DIAG_LD] - ((s0 + s1) + (s2 + s3));
      if (tid == j) {
        if (!(sum > 0.0)) {
          if (kblock * TB + j < n_valid) atomicCAS(info, 0, kblock * TB + j + 1);
        }
        dvec[j] = sqrt(sum);
      }
    }
    __syncthreads();
    if (tid >= j && tid < TB) {
      double dj = dvec[j];
      s[tid + j * DIAG_LD] = (tid == j) ? dj : sum / dj;
    }
    __syncthreads();
  }
  // write L back (lower part; upper part of the tile zeroed so later full-tile reads are clean)
  for (int idx = tid; idx < TB * TB; idx += 256) {
    int r = idx & 127, c = idx >> 7;
    Akk[(size_t)r + (size_t)c * lda] = (r >= c) ? s[r + c * DIAG_LD] : 0.0;
  }
  // log-determinant contribution (fixed-order reduction -> deterministic)
  if (tid < TB) xv[tid] = (kblock * TB + tid < n_valid) ? log(dvec[tid]) : 0.0;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if (tid < off) xv[tid] += xv[tid + off];
    __syncthreads();
  }
  if (tid == 0) logdet_part[kblock] = xv[0];
  __syncthreads();
  // ---- in-place inverse of the lower-triangular block (unblocked, columns right to left) ----
  for (int j = TB - 1; j >= 0; --j) {
    if (tid > j && tid < TB) xv[tid] = s[tid + j * DIAG_LD];  // L[r][j], r > j
    __syncthreads();
    double val = 0.0;
    if (tid > j && tid < TB) {
      double s0 = 0.0, s1 = 0.0;
      int c = j + 1;
      for (; c + 1 <= tid; c += 2) {
        s0 = fma(s[tid + c * DIAG_LD], xv[c], s0);
        s1 = fma(s[tid + (c + 1) * DIAG_LD], xv[c + 1], s1);
      }
      for (; c <= tid; ++c) s0 = fma(s[tid + c * DIAG_LD], xv[c], s0);
      val = -(s0 + s1) / dvec[j];
    }
    __syncthreads();
    if (tid > j && tid < TB) s[tid + j * DIAG_LD] = val;
    if (tid == j) s[j + j * DIAG_LD] = 1.0 / dvec[j];
    __syncthreads();
  }
  for (int idx = tid; idx < TB * TB; idx += 256) {
    int r = idx & 127, c = idx >> 7;
    Dkk[(size_t)r + (size_t)c * ldd] = (r >= c) ? s[r + c * DIAG_LD] : 0.0;
  }
}

// -----------------------------------------------------------------------------------------
// Vector work on the lower-triangular X = L^-1:  z = X Y,  a = X^T z  (so a = Sigma^-1 Y,
// reference R/GPfunctions.R:647, and Y'a = ||z||^2).  Tile partials + fixed-order reduction
// keep the result deterministic.

// zpart[tj][row] = sum_{c in tile tj} X[row, c] * y[c]      for lower tiles (ti >= tj)
__device__ __forceinline__ void trmv_n_part_body(const double* __restrict__ X, size_t ld,
                                                           const double* __restrict__ y,
                                                           double* __restrict__ part, size_t np_ld, const int bid) {
  __shared__ double ys[TB];
  __shared__ double red[TB];
  int ti, tj;
  tri_decode(bid, ti, tj);
  const int tid = threadIdx.x, r = tid & 127, half = tid >> 7;
  if (tid < TB) ys[tid] = y[(size_t)tj * TB + tid];
  __syncthreads();
  const double* xp = X + (size_t)ti * TB + r + (size_t)tj * TB * ld;
  double s0 = 0.0, s1 = 0.0;
  for (int c = half * 64; c < half * 64 + 64; c += 2) {
    s0 = fma(xp[(size_t)c * ld], ys[c], s0);
    s1 = fma(xp[(size_t)(c + 1) * ld], ys[c + 1], s1);
  }
  double s = s0 + s1;
  if (half == 1) red[r] = s;
  __syncthreads();
  if (half == 0) part[(size_t)tj * np_ld + (size_t)ti * TB + r] = s + red[r];
}
__global__ void __launch_bounds__(256) trmv_n_part_kernel(const double* __restrict__ X, size_t ld,
                                                           const double* __restrict__ y,
                                                           double* __restrict__ part, size_t np_ld) {
  trmv_n_part_body(X, ld, y, part, np_ld, (int)blockIdx.x);
}

// apart[ti][col] = sum_{r in tile ti} X[r, col] * z[r]       for lower tiles (ti >= tj)
__device__ __forceinline__ void trmv_t_part_body(const double* __restrict__ X, size_t ld,
                                                           const double* __restrict__ z,
                                                           double* __restrict__ part, size_t np_ld, const int bid) {
  __shared__ double zs[TB];
  int ti, tj;
  tri_decode(bid, ti, tj);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < TB) zs[tid] = z[(size_t)ti * TB + tid];
  __syncthreads();
  for (int c = warp; c < TB; c += 8) {
    const double* xp = X + (size_t)ti * TB + (size_t)(tj * TB + c) * ld;
    double s = xp[lane] * zs[lane];
    s = fma(xp[lane + 32], zs[lane + 32], s);
    s = fma(xp[lane + 64], zs[lane + 64], s);
    s = fma(xp[lane + 96], zs[lane + 96], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) part[(size_t)ti * np_ld + (size_t)tj * TB + c] = s;
  }
}
__global__ void __launch_bounds__(256) trmv_t_part_kernel(const double* __restrict__ X, size_t ld,
                                                           const double* __restrict__ z,
                                                           double* __restrict__ part, size_t np_ld) {
  trmv_t_part_body(X, ld, z, part, np_ld, (int)blockIdx.x);
}

// out[i] = sum over the tile partials that exist for i, in fixed order.
// mode 0: partial index tj runs 0..tile(i)   (z);  mode 1: ti runs tile(i)..nb-1   (a)
__device__ __forceinline__ void reduce_parts_body(const double* __restrict__ part, size_t np_ld, int nb, int mode,
                                    double* __restrict__ out, int n_alloc, const int bid) {
  int i = bid * blockDim.x + threadIdx.x;
  if (i >= n_alloc) return;
  int t = i / TB;
  double s = 0.0;
  if (mode == 0)
    for (int q = 0; q <= t; ++q) s += part[(size_t)q * np_ld + i];
  else
    for (int q = t; q < nb; ++q) s += part[(size_t)q * np_ld + i];
  out[i] = s;
}
__global__ void reduce_parts_kernel(const double* __restrict__ part, size_t np_ld, int nb, int mode,
                                    double* __restrict__ out, int n_alloc) {
  reduce_parts_body(part, np_ld, nb, mode, out, n_alloc, (int)blockIdx.x);
}

__device__ __forceinline__ void extract_diag_body(const double* __restrict__ S, size_t ld, int n, double* __restrict__ out, const int bid) {
  int i = bid * blockDim.x + threadIdx.x;
  if (i < n) out[i] = S[(size_t)i + (size_t)i * ld];
}
__global__ void extract_diag_kernel(const double* __restrict__ S, size_t ld, int n, double* __restrict__ out) {
  extract_diag_body(S, ld, n, out, (int)blockIdx.x);
}

// -----------------------------------------------------------------------------------------
// Gradient trace terms (reference R/GPfunctions.R:652-709):  with vQ = 0.5 (a a' - Sigma^-1),
//   T_k   = <vQ, -D_k o Cd>          (dl_k = 0.5 T_k, the reference's extra one half, :663)
//   T_ve  = <vQ, I>, T_s2 = <vQ, Cd/sigma2>, T_rho = <vQ, Cd o (1-popsame) / rho>
// Reads Sigma^-1 (lower tiles) and the kept copy of Sigma (Cd = Sigma off the diagonal, sigma2 on it: exactly the
// entries the likelihood was built from) once each; the per-coordinate squared distances D_k are recomputed from
// the staged predictor tiles instead of being stored (the reference keeps d dense N x N matrices, :539-542).
// Off-diagonal entries count twice (symmetry).  Round 1 also recomputed Cd here (distances + a software exp per
// entry: ~80 FP64 instructions, FP64-issue-bound at 0.35 ms for N = 8182); with Cd read back it is ~3 per coordinate.
// tracepart[cta][q], q = 0..d-1 (phi), d (ve), d+1 (sigma2), d+2 (rho).
// A thread owns two consecutive rows (16-byte loads of both matrices and of the staged predictors) and every
// fourth column of the CTA's column range.  The 1/sigma2 and 1/rho factors are applied once per CTA, not per entry.
template <int DC>
__device__ __forceinline__ void grad_trace_body(const double* __restrict__ S, const double* __restrict__ Sig, size_t ld, int N,
                                                          const double* __restrict__ X,
                                                          const int* __restrict__ pop,
                                                          const double* __restrict__ avec,
                                                          const EvalParams* __restrict__ P,
                                                          double* __restrict__ tracepart, int csplit, const int bid) {
  extern __shared__ double sm[];
  const int d = P->d;
  double* ri = sm;                  // raw predictors of the tile's rows  [d][128]
  double* rj = ri + d * TB;         // ... of its columns
  double* ai = rj + d * TB;         // [128]
  double* aj = ai + TB;
  double* red = aj + TB;            // [8][DC+3]
  int* pi_s = (int*)(red + 8 * (DC + 3));
  int* pj_s = pi_s + TB;
  int ti, tj;
  tri_decode(bid / csplit, ti, tj);
  const int ncols = TB / csplit, cbeg = (bid % csplit) * ncols;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < d * TB; idx += 256) {
    int k = idx >> 7, r = idx & 127;
    int gi = ti * TB + r, gj = tj * TB + r;
    ri[idx] = gi < N ? X[(size_t)gi + (size_t)k * N] : 0.0;
    rj[idx] = gj < N ? X[(size_t)gj + (size_t)k * N] : 0.0;
  }
  if (tid < TB) {
    int gi = ti * TB + tid, gj = tj * TB + tid;
    pi_s[tid] = gi < N ? pop[gi] : -1;
    pj_s[tid] = gj < N ? pop[gj] : -1;
    ai[tid] = gi < N ? avec[gi] : 0.0;
    aj[tid] = gj < N ? avec[gj] : 0.0;
  }
  __syncthreads();
  const double sigma2 = P->sigma2, rho = P->rho;
  const int r0 = 2 * (tid & 63), cg = tid >> 6;
  const int gi0 = ti * TB + r0;
  double acc[DC + 3];
#pragma unroll
  for (int q = 0; q < DC + 3; ++q) acc[q] = 0.0;
  const double a0 = ai[r0], a1 = ai[r0 + 1];
  const int p0 = pi_s[r0], p1 = pi_s[r0 + 1];
  const size_t off0 = (size_t)gi0 + (size_t)tj * TB * ld;
  const double* Sp = S + off0;
  const double* Cp = Sig + off0;
  // the pairs of the NEXT column are requested before the current one is processed (the loop is otherwise bound by
  // the latency of these loads)
  double2 s2n = *reinterpret_cast<const double2*>(Sp + (size_t)(cbeg + cg) * ld);
  double2 c2n = *reinterpret_cast<const double2*>(Cp + (size_t)(cbeg + cg) * ld);
  for (int c = cbeg + cg; c < cbeg + ncols; c += 4) {
    const int gj = tj * TB + c;
    const double2 s2 = s2n, c2 = c2n;
    if (c + 4 < cbeg + ncols) {
      s2n = *reinterpret_cast<const double2*>(Sp + (size_t)(c + 4) * ld);
      c2n = *reinterpret_cast<const double2*>(Cp + (size_t)(c + 4) * ld);
    }
    if (gj >= N || gj > gi0 + 1 || gi0 >= N) continue;  // nothing of this column pair lies in the lower triangle
    const int pj = pj_s[c];
    const double ajc = aj[c];
    const bool ok0 = gj <= gi0, ok1 = gi0 + 1 < N;  // (gj <= gi0 + 1 and gi0 < N hold here)
    const double q0 = 0.5 * (a0 * ajc - s2.x), q1 = 0.5 * (a1 * ajc - s2.y);
    // Cd: Sigma off the diagonal, sigma2 on it (Sigma_ii = sigma2 + ve); off-diagonal entries count twice
    const double qc0 = ok0 ? ((gi0 == gj) ? q0 * sigma2 : 2.0 * q0 * c2.x) : 0.0;
    const double qc1 = ok1 ? ((gi0 + 1 == gj) ? q1 * sigma2 : 2.0 * q1 * c2.y) : 0.0;
#pragma unroll
    for (int k = 0; k < DC; ++k) {
      if (k < d) {
        const double2 xr = *reinterpret_cast<const double2*>(ri + k * TB + r0);
        const double xc = rj[k * TB + c];
        const double f0 = xr.x - xc, f1 = xr.y - xc;
        acc[k] = fma(-qc0, f0 * f0, fma(-qc1, f1 * f1, acc[k]));
      }
    }
    if (gi0 == gj) acc[DC] += q0;
    if (gi0 + 1 == gj && ok1) acc[DC] += q1;
    acc[DC + 1] += qc0 + qc1;
    acc[DC + 2] += (p0 != pj ? qc0 : 0.0) + (p1 != pj ? qc1 : 0.0);
  }
  // block reduction in fixed order: warp shuffle tree, then 8 warp partials summed by warp 0
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int q = 0; q < DC + 3; ++q) {
    double v = acc[q];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[warp * (DC + 3) + q] = v;
  }
  __syncthreads();
  if (tid < DC + 3) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w * (DC + 3) + tid];
    if (tid == DC + 1) v /= sigma2;
    if (tid == DC + 2) v /= rho;
    int qo = tid < DC ? tid : d + (tid - DC);
    if (tid >= DC || tid < d) tracepart[(size_t)bid * (MAXD + 3) + qo] = v;
  }
}
template <int DC>
__global__ void __launch_bounds__(256, DC <= 16 ? 3 : 1) grad_trace_kernel(const double* __restrict__ S, const double* __restrict__ Sig,
                                                          size_t ld, int N,
                                                          const double* __restrict__ X,
                                                          const int* __restrict__ pop,
                                                          const double* __restrict__ avec,
                                                          const EvalParams* __restrict__ P,
                                                          double* __restrict__ tracepart, int csplit) {
  grad_trace_body<DC>(S, Sig, ld, N, X, pop, avec, P, tracepart, csplit, (int)blockIdx.x);
}

// Final scalar reductions of one evaluation (single CTA, fixed order => deterministic):
//  scal[0]        = logdet = sum log L_ii                (:648, :729)
//  scal[1]        = SS = ||z||^2 = Y' a                   (:728)
//  scal[2]        = sum log diag(S)      \  lnL_LOO (:742)
//  scal[3]        = sum a_i^2 / S_ii     /
//  scal[4]        = tr(S)                 -> df = N - ve tr(S)  (:743)
//  scal[8 + q]    = trace terms q = 0..d+2
__device__ __forceinline__ void final_reduce_body(const double* __restrict__ logdet_part, int nb,
                                                            const double* __restrict__ z,
                                                            const double* __restrict__ avec,
                                                            const double* __restrict__ diagS, int N,
                                                            const double* __restrict__ tracepart, int ntiles,
                                                            int nq, int have_S, double* __restrict__ scal, const int bid) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  auto block_sum = [&](double v) -> double {
    red[tid] = v;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (tid < off) red[tid] += red[tid + off];
      __syncthreads();
    }
    double r = red[0];
    __syncthreads();
    return r;
  };
  double v = 0.0;
  for (int i = tid; i < nb; i += 256) v += logdet_part[i];
  double logdet = block_sum(v);
  v = 0.0;
  for (int i = tid; i < N; i += 256) v = fma(z[i], z[i], v);
  double ss = block_sum(v);
  double sl = 0.0, sq = 0.0, tr = 0.0;
  if (have_S) {
    double v1 = 0.0, v2 = 0.0, v3 = 0.0;
    for (int i = tid; i < N; i += 256) {
      double sii = diagS[i];
      v1 += log(sii);
      v2 += avec[i] * avec[i] / sii;
      v3 += sii;
    }
    sl = block_sum(v1);
    sq = block_sum(v2);
    tr = block_sum(v3);
  }
  if (tid == 0) {
    scal[0] = logdet;
    scal[1] = ss;
    scal[2] = sl;
    scal[3] = sq;
    scal[4] = tr;
  }
  for (int q = 0; q < nq; ++q) {
    v = 0.0;
    for (int t = tid; t < ntiles; t += 256) v += tracepart[(size_t)t * (MAXD + 3) + q];
    double s = block_sum(v);
    if (tid == 0) scal[8 + q] = s;
  }
}
__global__ void __launch_bounds__(256) final_reduce_kernel(const double* __restrict__ logdet_part, int nb,
                                                            const double* __restrict__ z,
                                                            const double* __restrict__ avec,
                                                            const double* __restrict__ diagS, int N,
                                                            const double* __restrict__ tracepart, int ntiles,
                                                            int nq, int have_S, double* __restrict__ scal) {
  final_reduce_body(logdet_part, nb, z, avec, diagS, N, tracepart, ntiles, nq, have_S, scal, (int)blockIdx.x);
}

// -----------------------------------------------------------------------------------------
// Prediction helpers.  V = Cs^T (Np x Mp, zero padded), U = L^-1 V.
// mean[m] = sum_n V[n,m] a[n]  (reference :1017);  gpgrad[m,k] = -2 phi_k sum_n (xnew_mk - x_nk) V[n,m] a[n]
// (getGPgrad :1863-1865, unscaled coordinates);  var[m] = sigma2 - sum_n U[n,m]^2 (:1020, evaluated as
// sigma2 - ||L^-1 c||^2).  One CTA per new point.
__global__ void __launch_bounds__(256) predict_reduce_kernel(const double* __restrict__ V,
                                                              const double* __restrict__ U, size_t ldv, int N,
                                                              int M, const double* __restrict__ avec,
                                                              const double* __restrict__ X,
                                                              const double* __restrict__ Xnew,
                                                              const EvalParams* __restrict__ P,
                                                              double* __restrict__ mean, double* __restrict__ var,
                                                              double* __restrict__ gpgrad) {
  __shared__ double red[256];
  const int m = blockIdx.x, tid = threadIdx.x;
  const int d = P->d;
  auto block_sum = [&](double v) -> double {
    red[tid] = v;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (tid < off) red[tid] += red[tid + off];
      __syncthreads();
    }
    double r = red[0];
    __syncthreads();
    return r;
  };
  const double* v = V + (size_t)m * ldv;
  double s = 0.0;
  for (int n = tid; n < N; n += 256) s = fma(v[n], avec[n], s);
  s = block_sum(s);
  if (tid == 0) mean[m] = s;
  if (U != nullptr) {
    const double* u = U + (size_t)m * ldv;
    double q = 0.0;
    for (int n = tid; n < N; n += 256) q = fma(u[n], u[n], q);
    q = block_sum(q);
    if (tid == 0) var[m] = P->sigma2 - q;
  }
  if (gpgrad != nullptr) {
    for (int k = 0; k < d; ++k) {
      double xk = Xnew[(size_t)m + (size_t)k * M];
      double g = 0.0;
      for (int n = tid; n < N; n += 256) g = fma((xk - X[(size_t)n + (size_t)k * N]) * v[n], avec[n], g);
      g = block_sum(g);
      if (tid == 0) gpgrad[(size_t)m + (size_t)k * M] = -2.0 * P->phi[k] * g;
    }
  }
}

// Block leave-out solves (replaces the per-point refactorisations of reference
// R/GPfunctions.R:1056-1104 and :1140-1194).  One CTA per excluded set B_g: gather the symmetric
// block S_BB = Sigma^-1[B_g, B_g], factor it (right-looking Cholesky in the CTA's workspace) and return
//   u = S_BB^-1 a_B        (leave-out mean of row i in B:  Y_i - u_i)
//   dinv = diag(S_BB^-1)   (leave-out variance:            dinv_i - ve)
// idx: concatenated member rows; goff[g]: start of group g in idx; boff[g]: start of its |B|x|B|
// block in `work` (two column-major copies per group: the factor, then the columns of its inverse).
// `bad` (initialised to INT_MAX) receives the smallest 1-based group whose block is not positive definite.
__global__ void __launch_bounds__(128)
    leaveout_blocks_kernel(const double* __restrict__ S, size_t ld, const int* __restrict__ idx,
                           const int* __restrict__ goff, const long long* __restrict__ boff, int ngroups,
                           const double* __restrict__ a, double* __restrict__ work, double* __restrict__ u,
                           double* __restrict__ dinv, int* __restrict__ bad) {
  const int g = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  if (g >= ngroups) return;
  const int b0 = goff[g], k = goff[g + 1] - b0;
  if (k == 0) return;
  if (k == 1) {  // plain leave-one-out: scalar closed form
    if (tid == 0) {
      const int m = idx[b0];
      const double sii = S[(size_t)m + (size_t)m * ld];
      if (!(sii > 0.0)) atomicMin(bad, g + 1);
      u[b0] = a[m] / sii;
      dinv[b0] = 1.0 / sii;
    }
    return;
  }
  double* L = work + 2 * boff[g];
  double* Wi = L + (size_t)k * k;
  for (int e = tid; e < k * k; e += T) {
    const int r = e % k, c = e / k;
    const int ir = idx[b0 + r], ic = idx[b0 + c];
    const int hi = max(ir, ic), lo = min(ir, ic);
    L[e] = S[(size_t)hi + (size_t)lo * ld];
  }
  for (int r = tid; r < k; r += T) u[b0 + r] = a[idx[b0 + r]];
  __syncthreads();
  // ---- Cholesky (lower), column by column
  for (int j = 0; j < k; ++j) {
    const double piv = L[j + (size_t)j * k];
    __syncthreads();
    if (!(piv > 0.0)) {  // uniform across the CTA
      if (tid == 0) atomicMin(bad, g + 1);
      return;
    }
    const double dj = sqrt(piv);
    if (tid == 0) L[j + (size_t)j * k] = dj;
    for (int i = j + 1 + tid; i < k; i += T) L[i + (size_t)j * k] /= dj;
    __syncthreads();
    const int m = k - j - 1;
    for (int e = tid; e < m * m; e += T) {
      const int i = j + 1 + e % m, c = j + 1 + e / m;
      if (i >= c) L[i + (size_t)c * k] -= L[i + (size_t)j * k] * L[c + (size_t)j * k];
    }
    __syncthreads();
  }
  // ---- u = L^-T L^-1 a_B
  for (int j = 0; j < k; ++j) {
    if (tid == 0) u[b0 + j] /= L[j + (size_t)j * k];
    __syncthreads();
    const double uj = u[b0 + j];
    for (int i = j + 1 + tid; i < k; i += T) u[b0 + i] -= L[i + (size_t)j * k] * uj;
    __syncthreads();
  }
  for (int j = k - 1; j >= 0; --j) {
    if (tid == 0) u[b0 + j] /= L[j + (size_t)j * k];
    __syncthreads();
    const double uj = u[b0 + j];
    for (int i = tid; i < j; i += T) u[b0 + i] -= L[j + (size_t)i * k] * uj;
    __syncthreads();
  }
  // ---- diag(S_BB^-1)_r = || L^-1 e_r ||^2 : one forward solve per column, a thread each
  for (int r = tid; r < k; r += T) {
    double* w = Wi + (size_t)r * k;
    double acc = 0.0;
    for (int i = r; i < k; ++i) {
      double t = (i == r) ? 1.0 : 0.0;
      for (int q = r; q < i; ++q) t -= L[i + (size_t)q * k] * w[q];
      t /= L[i + (size_t)i * k];
      w[i] = t;
      acc += t * t;
    }
    dinv[b0 + r] = acc;
  }
}

// Sequential (leave-future-out) prediction from the Cholesky factor of the time-major ordered
// matrix (reference :1196-1245, identities in SURVEY appendix B):
//   mean_r = sum_{c < bstart(r)} L[r,c] z_c,   sd_r = sqrt( sum_{bstart(r) <= c <= r} L[r,c]^2 )
// One warp per row.
__global__ void __launch_bounds__(256) sequential_kernel(const double* __restrict__ L, size_t ld, int N,
                                                          const double* __restrict__ z,
                                                          const int* __restrict__ bstart,
                                                          double* __restrict__ mean, double* __restrict__ sd) {
  int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (r >= N) return;
  int b0 = bstart[r];
  double m = 0.0, q = 0.0;
  for (int c = lane; c <= r; c += 32) {
    double l = L[(size_t)r + (size_t)c * ld];
    if (c < b0)
      m = fma(l, z[c], m);
    else
      q = fma(l, l, q);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    m += __shfl_xor_sync(0xffffffffu, m, off);
    q += __shfl_xor_sync(0xffffffffu, q, off);
  }
  if (lane == 0) {
    mean[r] = m;
    sd[r] = sqrt(q);
  }
}

// LOO / LTO GP-gradient (getGPgrad with the focal group's rows removed, reference :1081-1085,
// :1168-1175).  For focal row i with excluded set B and u = (S_BB)^-1 a_B:
//   a^(-B)_j = a_j - sum_{b in B} S[j,b] u_b  (j not in B),
//   grad_k = -2 phi_k sum_{j not in B} (x_ik - x_jk) Cd[i,j] a^(-B)_j.
// One CTA per focal row.  members/u are the group's concatenated arrays.
__global__ void __launch_bounds__(256) loo_gpgrad_kernel(const double* __restrict__ S, size_t ld, int N,
                                                          const double* __restrict__ X,
                                                          const int* __restrict__ pop,
                                                          const double* __restrict__ rhotab,
                                                          const double* __restrict__ avec,
                                                          const EvalParams* __restrict__ P,
                                                          const int* __restrict__ focal,
                                                          const int* __restrict__ fgroup,
                                                          const int* __restrict__ idx,
                                                          const int* __restrict__ goff,
                                                          const double* __restrict__ uvec, int nfocal,
                                                          double* __restrict__ gpgrad) {
  __shared__ double red[256];
  __shared__ double xis[MAXD], xir[MAXD];
  const int f = blockIdx.x, tid = threadIdx.x;
  if (f >= nfocal) return;
  const int i = focal[f], g = fgroup[f];
  const int b0 = goff[g], nbk = goff[g + 1] - b0;
  const int d = P->d;
  if (tid < d) {
    xir[tid] = X[(size_t)i + (size_t)tid * N];
    xis[tid] = __dmul_rn(xir[tid], P->sqphi[tid]);
  }
  __syncthreads();
  double acc[MAXD];
  for (int k = 0; k < d; ++k) acc[k] = 0.0;
  const double* rt = P->use_rhotab ? rhotab : nullptr;
  for (int j = tid; j < N; j += 256) {
    bool inB = false;
    double corr = 0.0;
    for (int b = 0; b < nbk; ++b) {
      int ib = idx[b0 + b];
      if (ib == j) inB = true;
      int hi = max(j, ib), lo = min(j, ib);
      corr = fma(S[(size_t)hi + (size_t)lo * ld], uvec[b0 + b], corr);
    }
    if (inB) continue;
    double aj = avec[j] - corr;
    double dist = 0.0;
    for (int k = 0; k < d; ++k) {
      double diff = __dsub_rn(xis[k], __dmul_rn(X[(size_t)j + (size_t)k * N], P->sqphi[k]));
      dist = __dadd_rn(dist, __dmul_rn(diff, diff));
    }
    double v = __dmul_rn(P->sigma2, exp(-dist));
    if (rt != nullptr)
      v = __dmul_rn(v, rt[pop[i] + pop[j] * P->np]);
    else if (pop[i] != pop[j])
      v = __dmul_rn(v, P->rho);
    double w = v * aj;
    for (int k = 0; k < d; ++k) acc[k] = fma(xir[k] - X[(size_t)j + (size_t)k * N], w, acc[k]);
  }
  for (int k = 0; k < d; ++k) {
    red[tid] = acc[k];
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (tid < off) red[tid] += red[tid + off];
      __syncthreads();
    }
    if (tid == 0) gpgrad[(size_t)f + (size_t)k * nfocal] = -2.0 * P->phi[k] * red[0];
    __syncthreads();
  }
}

// FP64 issue-rate probes (roofline denominators).
__global__ void __launch_bounds__(256) probe_dmma_kernel(double* out, int iters) {
  double c[32][2];
#pragma unroll
  for (int i = 0; i < 32; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) probe_dfma_kernel(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace gpedm
