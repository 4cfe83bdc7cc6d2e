// On-device data preparation of fitGP(data, y, E, tau, scaling): delay embedding
// inside each population (reference makelags, R/GPfunctions.R:1824-1857), centring / scaling
// (:359-405, statistics over all non-missing entries of a column BEFORE incomplete rows are dropped)
// and removal of incomplete rows (:445-451).  A model-selection grid uploads the raw series once per
// GPU; every (E, tau) cell builds its training set here, straight into the handle's device buffers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gpedm {

struct LagPrep {
  const double* y;     // T raw values (NaN = missing)
  const int* pop;      // T population codes 0..np-1
  const int* rank;     // T: position of row t inside its population (row order)
  const int* poprows;  // T: rows of each population in order, concatenated
  const int* popoff;   // np + 1 offsets into poprows
  double* center;      // (E+1) x G column statistics, index col + (E+1) g
  double* scale;
  int T, E, tau, np;
  int G;               // scaling groups: 1 (none / global) or np (local)
  int mode;            // 0 none, 1 global, 2 local
};

// col 0 = y itself, col k >= 1 = y lagged by k*tau inside the population of row t
__device__ __forceinline__ double lag_raw(const LagPrep& P, int t, int col) {
  if (col == 0) return P.y[t];
  const int r = P.rank[t] - col * P.tau;
  if (r < 0) return nan("");
  return P.y[P.poprows[P.popoff[P.pop[t]] + r]];
}

__device__ __forceinline__ double lag_scaled(const LagPrep& P, int t, int col) {
  const double v = lag_raw(P, t, col);
  if (P.mode == 0) return v;
  const int g = (P.mode == 2) ? P.pop[t] : 0;
  const int q = col + (P.E + 1) * g;
  return (v - P.center[q]) / P.scale[q];  // same two roundings as (x - mean) / sd on the host
}

// Deterministic CTA sum (fixed-order shared-memory tree); every thread returns the total.
__device__ __forceinline__ double cta_sum_256(double v, double* red) {
  const int tid = threadIdx.x;
  __syncthreads();
  red[tid] = v;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) red[tid] += red[tid + off];
    __syncthreads();
  }
  return red[0];
}

// mean and sample standard deviation (two-pass, n-1) of every lag column per scaling group.
// grid (E+1, G), 256 threads.
__global__ void __launch_bounds__(256) lag_stats_kernel(LagPrep P) {
  __shared__ double red[256];
  const int col = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
  const int r0 = (P.mode == 2) ? P.popoff[g] : 0, r1 = (P.mode == 2) ? P.popoff[g + 1] : P.T;
  double s = 0.0, n = 0.0;
  for (int r = r0 + tid; r < r1; r += 256) {
    const int t = (P.mode == 2) ? P.poprows[r] : r;
    const double v = lag_raw(P, t, col);
    if (!isnan(v)) {
      s += v;
      n += 1.0;
    }
  }
  s = cta_sum_256(s, red);
  n = cta_sum_256(n, red);
  const double mean = n > 0 ? s / n : nan("");
  double ss = 0.0;
  for (int r = r0 + tid; r < r1; r += 256) {
    const int t = (P.mode == 2) ? P.poprows[r] : r;
    const double v = lag_raw(P, t, col);
    if (!isnan(v)) ss += (v - mean) * (v - mean);
  }
  ss = cta_sum_256(ss, red);
  if (tid == 0) {
    const int q = col + (P.E + 1) * g;
    P.center[q] = mean;
    P.scale[q] = n > 1 ? sqrt(ss / (n - 1.0)) : nan("");
  }
}

// Complete-row compaction: rowmap[0..N) = rows whose scaled y and E lags are all non-NaN, in
// order; *count = N.  One CTA of 1024 threads walks the series in chunks (ordered, deterministic).
__global__ void __launch_bounds__(1024) lag_compact_kernel(LagPrep P, int* __restrict__ rowmap, int* __restrict__ count) {
  __shared__ int wsum[32];
  __shared__ int base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < P.T; c0 += 1024) {
    const int t = c0 + tid;
    int keep = 0;
    if (t < P.T) {
      keep = 1;
      for (int col = 0; col <= P.E && keep; ++col) keep = !isnan(lag_scaled(P, t, col));
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int inwarp = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wsum[warp] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    if (keep) rowmap[base + before + inwarp] = t;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += wsum[w];
      base += tot;
    }
    __syncthreads();
  }
  if (tid == 0) *count = base;
}

// Write the training set of the kept rows: Y (N), X (N x E column-major), pop (N).
__global__ void __launch_bounds__(256)
    lag_fill_kernel(LagPrep P, const int* __restrict__ rowmap, int N, double* __restrict__ Y, double* __restrict__ X,
                    int* __restrict__ popout) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int col = blockIdx.y;
  if (i >= N) return;
  const int t = rowmap[i];
  const double v = lag_scaled(P, t, col);
  if (col == 0) {
    Y[i] = v;
    popout[i] = P.pop[t];
  } else {
    X[(size_t)i + (size_t)(col - 1) * N] = v;
  }
}

// In-sample and leave-one-out fit statistics in the ORIGINAL units of y (reference getR2,
// R/GPfunctions.R:1366-1370, applied to the back-transformed predictions :469-482):
//   obs_i = Y_i s_g + c_g,  in-sample residual = ve a_i s_g,  LOO residual = (a_i / S_ii) s_g.
// out = [n, mean(obs), sum (obs-mean)^2, sum res_in^2, sum res_loo^2].  One CTA, two passes.
__global__ void __launch_bounds__(256)
    fit_stats_kernel(const double* __restrict__ Y, const double* __restrict__ a, const double* __restrict__ diagS,
                     const int* __restrict__ pop, int N, double ve, const double* __restrict__ center,
                     const double* __restrict__ scale, int stride, int mode, double* __restrict__ out) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  auto sc = [&](int i, double& c, double& s) {
    if (mode == 0) {
      c = 0.0;
      s = 1.0;
    } else {
      const int g = (mode == 2) ? pop[i] : 0;
      c = center[(size_t)stride * g];
      s = scale[(size_t)stride * g];
    }
  };
  double sum = 0.0;
  for (int i = tid; i < N; i += 256) {
    double c, s;
    sc(i, c, s);
    sum += Y[i] * s + c;
  }
  sum = cta_sum_256(sum, red);
  const double mean = sum / N;
  double sst = 0.0, sin = 0.0, sloo = 0.0;
  for (int i = tid; i < N; i += 256) {
    double c, s;
    sc(i, c, s);
    const double o = Y[i] * s + c - mean;
    const double ri = ve * a[i] * s, rl = a[i] / diagS[i] * s;
    sst += o * o;
    sin += ri * ri;
    sloo += rl * rl;
  }
  sst = cta_sum_256(sst, red);
  sin = cta_sum_256(sin, red);
  sloo = cta_sum_256(sloo, red);
  if (tid == 0) {
    out[0] = (double)N;
    out[1] = mean;
    out[2] = sst;
    out[3] = sin;
    out[4] = sloo;
  }
}

}  // namespace gpedm
