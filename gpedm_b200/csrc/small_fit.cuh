// Batched small fits: one CTA evaluates the whole log-posterior + gradient of one model with
// N <= 128 rows (the E/tau grids of short series, per-population models, the pairwise fits of
// getrhomatrix, reference R/rhomatrix.R:169-182) entirely in shared memory, so hundreds of
// independent fits advance together, one launch per Rprop iteration for the whole batch.
// Same arithmetic as the tiled pipeline (assemble -> Cholesky -> L^-1 -> a = Sigma^-1 Y ->
// Sigma^-1 = L^-T L^-1 -> trace terms, reference R/GPfunctions.R:633-709), with Sigma^-1 never
// written out: each 8x8 DMMA tile of it is consumed for the trace terms while still in registers.
#pragma once
#include "diag_block.cuh"
#include "kernels.cuh"

namespace gpedm {

struct SmallFit {
  const double* X;       // N x d column-major
  const double* Y;       // N
  const int* pop;        // N
  const double* rhotab;  // np x np or nullptr
  double* scal;          // 8 + MAXD + 3 reductions, same layout as final_reduce_kernel
  double* a_out;         // N: Sigma^-1 Y      (written when `full`)
  double* diag_out;      // N: diag(Sigma^-1)  (written when `full`)
  int* info;             // first non-positive pivot (1-based) or 0
  int N, d, np, pad;
};

template <int DC>
constexpr int small_eval_smem_bytes() {
  return (TB * DG_LD + 64 * DG_WLD + 2 * TB + 64 + DC * TB + 3 * TB + 16 * (DC + 8)) * (int)sizeof(double) +
         TB * (int)sizeof(int);
}

template <int DC>
__global__ void __launch_bounds__(DG_THREADS, 1)
    small_eval_kernel(const SmallFit* __restrict__ fits, const EvalParams* __restrict__ params,
                      const int* __restrict__ active, int full) {
  extern __shared__ double sm[];
  double* s = sm;                      // [128][132]
  double* w = s + TB * DG_LD;          // [64][68]
  double* rdiag = w + 64 * DG_WLD;
  double* dvec = rdiag + TB;
  double* colbuf = dvec + TB;          // [2][32]
  double* xr = colbuf + 64;            // raw predictors [DC][128]
  double* ys = xr + DC * TB;
  double* zs = ys + TB;
  double* as = zs + TB;
  double* red = as + TB;               // [16][DC + 8]
  int* pops = (int*)(red + 16 * (DC + 8));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int f = active[blockIdx.x];
  const SmallFit F = fits[f];
  const EvalParams* P = params + f;
  const int N = F.N, d = F.d, np = F.np;
  const double sigma2 = P->sigma2, rho = P->rho, ve = P->ve;
  const double* rt = P->use_rhotab ? F.rhotab : nullptr;

  for (int idx = tid; idx < DC * TB; idx += DG_THREADS) {
    const int k = idx >> 7, r = idx & 127;
    xr[idx] = (k < d && r < N) ? F.X[(size_t)r + (size_t)k * N] : 0.0;
  }
  if (tid < TB) {
    ys[tid] = tid < N ? F.Y[tid] : 0.0;
    pops[tid] = tid < N ? F.pop[tid] : -1;
  }
  if (tid == 0) *F.info = 0;
  __syncthreads();

  // covariance entry with the arithmetic order of getcov (see cov_entry in kernels.cuh)
  auto cov = [&](int i, int j) -> double {
    double dist = 0.0;
    for (int k = 0; k < d; ++k) {
      const double sq = P->sqphi[k];
      const double diff = __dsub_rn(__dmul_rn(xr[k * TB + i], sq), __dmul_rn(xr[k * TB + j], sq));
      dist = __dadd_rn(dist, __dmul_rn(diff, diff));
    }
    double v = __dmul_rn(sigma2, exp(-dist));
    if (rt != nullptr)
      v = __dmul_rn(v, rt[pops[i] + pops[j] * np]);
    else if (pops[i] != pops[j])
      v = __dmul_rn(v, rho);
    return v;
  };

  // ---- Sigma = Cd + ve I (full tile; identity in the padding)
  {
    const int r = tid & 127, q4 = tid >> 7;
    for (int c = q4 * 32; c < q4 * 32 + 32; ++c) {
      double v;
      if (r < N && c < N) {
        v = cov(r, c);
        if (r == c) v = __dadd_rn(v, ve);
      } else {
        v = (r == c) ? 1.0 : 0.0;
      }
      s[r + c * DG_LD] = v;
    }
  }
  __syncthreads();

  // ---- Cholesky, log-determinant, L^-1
  diag_cholesky_smem(s, rdiag, dvec, colbuf, 0, N, F.info);
  double logdet;
  {
    double v = (tid < N) ? log(dvec[tid]) : 0.0;
    __syncthreads();
    if (tid < TB) dvec[tid] = v;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
      if (tid < off) dvec[tid] += dvec[tid + off];
      __syncthreads();
    }
    logdet = dvec[0];
  }
  diag_invert_smem(s, w, rdiag);  // ends with a CTA barrier

  // ---- z = X Y, a = X^T z  (X = L^-1 lower triangular)
  if (tid < TB) {
    double z0 = 0.0, z1 = 0.0;
    int c = 0;
    for (; c + 1 <= tid; c += 2) {
      z0 = fma(s[tid + c * DG_LD], ys[c], z0);
      z1 = fma(s[tid + (c + 1) * DG_LD], ys[c + 1], z1);
    }
    if (c <= tid) z0 = fma(s[tid + c * DG_LD], ys[c], z0);
    zs[tid] = z0 + z1;
  }
  __syncthreads();
  for (int c = warp; c < TB; c += DG_THREADS / 32) {
    double v = 0.0;
    for (int r = c + lane; r < TB; r += 32) v = fma(s[r + c * DG_LD], zs[r], v);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) as[c] = v;
  }
  __syncthreads();

  // ---- Sigma^-1 = X^T X tile by tile (8x8 DMMA tiles, lower triangle), consumed in registers
  double acc[DC + 3];
#pragma unroll
  for (int q = 0; q < DC + 3; ++q) acc[q] = 0.0;
  double sl = 0.0, sqa = 0.0, tr = 0.0;
  for (int q = warp; q < 136; q += DG_THREADS / 32) {
    int ti, tj;
    tri_decode(q, ti, tj);
    const int kbeg = 8 * ti;
    const double* ap = s + (kbeg + t) + (8 * ti + g) * DG_LD;
    const double* bp = s + (kbeg + t) + (8 * tj + g) * DG_LD;
    double c0 = 0.0, c1 = 0.0;
    for (int k = kbeg; k < TB; k += 4) {
      dmma884(c0, c1, ap[0], bp[0]);
      ap += 4;
      bp += 4;
    }
    const int gi = 8 * ti + g;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int gj = 8 * tj + 2 * t + e;
      if (gi >= N || gj > gi) continue;
      const double sij = e ? c1 : c0;
      const double qv = 0.5 * (as[gi] * as[gj] - sij);
      const double cd = cov(gi, gj);
      const double qc = ((gi == gj) ? 1.0 : 2.0) * qv * cd;
#pragma unroll
      for (int k = 0; k < DC; ++k)
        if (k < d) {
          const double df = xr[k * TB + gi] - xr[k * TB + gj];
          acc[k] -= qc * (df * df);
        }
      acc[DC + 1] += qc / sigma2;
      if (pops[gi] != pops[gj]) acc[DC + 2] += qc / rho;
      if (gi == gj) {
        acc[DC] += qv;
        sl += log(sij);
        sqa += as[gi] * as[gi] / sij;
        tr += sij;
        if (full) F.diag_out[gi] = sij;
      }
    }
  }
  // ---- deterministic CTA reductions: shuffle tree per warp, then the 16 warp partials in order
  const int NR = DC + 3 + 3;
  auto warp_sum = [&](double v) -> double {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  };
#pragma unroll
  for (int q = 0; q < DC + 3; ++q) {
    const double v = warp_sum(acc[q]);
    if (lane == 0) red[warp * (DC + 8) + q] = v;
  }
  {
    const double v0 = warp_sum(sl), v1 = warp_sum(sqa), v2 = warp_sum(tr);
    if (lane == 0) {
      red[warp * (DC + 8) + DC + 3] = v0;
      red[warp * (DC + 8) + DC + 4] = v1;
      red[warp * (DC + 8) + DC + 5] = v2;
    }
  }
  __syncthreads();
  if (tid < NR) {
    double v = 0.0;
    for (int wq = 0; wq < DG_THREADS / 32; ++wq) v += red[wq * (DC + 8) + tid];
    if (tid < DC) {
      if (tid < d) F.scal[8 + tid] = v;
    } else if (tid < DC + 3) {
      F.scal[8 + d + (tid - DC)] = v;
    } else {
      F.scal[2 + (tid - DC - 3)] = v;  // sum log S_ii, sum a_i^2 / S_ii, tr S
    }
  }
  if (tid == 0) {
    double ss = 0.0;
    for (int i = 0; i < N; ++i) ss = fma(zs[i], zs[i], ss);
    F.scal[0] = logdet;
    F.scal[1] = ss;
  }
  if (full && tid < N) F.a_out[tid] = as[tid];
}

}  // namespace gpedm
