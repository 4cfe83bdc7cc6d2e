// Lockstep batch of MID-SIZE fits (129 <= N <= ~1024: per-population models, pairwise fits of getrhomatrix,
// multistart, b-profiles of fitGP_fish, grids on short series - reference clients R/rhomatrix.R:169-182,
// vignettes/hierarchical.Rmd:47, :89-97, R/GPfisheries.R:114-134).  One such fit cannot fill a B200: its
// Cholesky chain leaves 140 SMs idle and every evaluation costs a dozen launches.  Here ALL fits of the batch
// advance one Rprop iteration per round of launches: every kernel of the evaluation pipeline runs once over
// the tiles of every active fit (a per-CTA map says which fit and which tile), and the persistent
// factorisation kernel (dag_factor.cuh) gets ONE task list in which the fits' tasks are interleaved, so a CTA
// that would wait for one fit's chain picks up another fit's tile.
#pragma once
#include "kernels.cuh"
#include "dag_factor.cuh"

namespace gpedm {

struct MidFit {  // device-side description of one fit of the batch
  double *A, *X, *S, *D;  // Sigma -> L, L^-1, Sigma^-1, kept copy of Sigma (Np x Np, lower tiles)
  size_t ld;
  int N, nb, ntiles, split, d, pad;
  const double* dX;
  const int* dpop;
  const double* drhotab;
  const EvalParams* P;
  const double* dY;
  double *dz, *da, *ddiagS, *dpart, *dlogdet, *dtracepart, *dscal;
};

// map[b] = (fit, block index inside that fit's grid)
__global__ void __launch_bounds__(256) mid_assemble_kernel(const MidFit* __restrict__ fits, const int2* __restrict__ map) {
  const int2 m = map[blockIdx.x];
  const MidFit f = fits[m.x];
  assemble_sigma_body(f.A, f.D, f.ld, f.N, f.dX, f.dpop, f.drhotab, f.P, f.split, m.y);
}
__global__ void __launch_bounds__(256) mid_trmv_n_kernel(const MidFit* __restrict__ fits, const int2* __restrict__ map) {
  const int2 m = map[blockIdx.x];
  const MidFit f = fits[m.x];
  trmv_n_part_body(f.X, f.ld, f.dY, f.dpart, f.ld, m.y);
}
__global__ void __launch_bounds__(256) mid_trmv_t_kernel(const MidFit* __restrict__ fits, const int2* __restrict__ map) {
  const int2 m = map[blockIdx.x];
  const MidFit f = fits[m.x];
  trmv_t_part_body(f.X, f.ld, f.dz, f.dpart, f.ld, m.y);
}
__global__ void __launch_bounds__(256) mid_reduce_kernel(const MidFit* __restrict__ fits, const int2* __restrict__ map, int mode) {
  const int2 m = map[blockIdx.x];
  const MidFit f = fits[m.x];
  reduce_parts_body(f.dpart, f.ld, f.nb, mode, mode == 0 ? f.dz : f.da, f.nb * TB, m.y);
}
__global__ void __launch_bounds__(256) mid_extract_diag_kernel(const MidFit* __restrict__ fits, const int2* __restrict__ map) {
  const int2 m = map[blockIdx.x];
  const MidFit f = fits[m.x];
  extract_diag_body(f.S, f.ld, f.N, f.ddiagS, m.y);
}
template <int DC>
__global__ void __launch_bounds__(256, DC <= 16 ? 3 : 1) mid_trace_kernel(const MidFit* __restrict__ fits, const int2* __restrict__ map) {
  const int2 m = map[blockIdx.x];
  const MidFit f = fits[m.x];
  grad_trace_body<DC>(f.S, f.D, f.ld, f.N, f.dX, f.dpop, f.da, f.P, f.dtracepart, f.split, m.y);
}
// one CTA per active fit
__global__ void __launch_bounds__(256) mid_final_reduce_kernel(const MidFit* __restrict__ fits, const int* __restrict__ active, int have_S) {
  const MidFit f = fits[active[blockIdx.x]];
  final_reduce_body(f.dlogdet, f.nb, f.dz, f.da, f.ddiagS, f.N, f.dtracepart, f.ntiles * f.split, f.d + 3, have_S, f.dscal, 0);
}

}  // namespace gpedm
