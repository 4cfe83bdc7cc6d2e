// Driver around the DMMA tile contraction: times several kernel shapes (GemmCfgT) on random data,
// for both the long-k contraction and the K=128 trailing-update shape; also used for ncu captures.
// usage: gemm_probe [n=4096] [reps=3] [cfg index or -1 for all]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../gpedm_b200/csrc/dgemm_tile.cuh"
using namespace gpedm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_nt_kernel(const double* A, const double* B, double* C, size_t ld, int K) {
  extern __shared__ double smem[];
  const double* At = A + (size_t)blockIdx.x * TB;
  const double* Bt = B + (size_t)blockIdx.y * TB;
  double* Ct = C + (size_t)blockIdx.x * TB + (size_t)blockIdx.y * TB * ld;
  gemm_tile<false, false, Cfg>(At, ld, Bt, ld, 0, K, false, nullptr, 0, Ct, ld, smem);
}
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_nt_mb_kernel(const double* A, const double* B, double* C, size_t ld, int K) {
  extern __shared__ double smem[];
  const double* At = A + (size_t)blockIdx.x * TB;
  const double* Bt = B + (size_t)blockIdx.y * TB;
  double* Ct = C + (size_t)blockIdx.x * TB + (size_t)blockIdx.y * TB * ld;
  gemm_tile_mb<false, false, Cfg>(At, ld, Bt, ld, 0, K, false, nullptr, 0, Ct, ld, smem);
}
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
syrk_trail_mb_kernel(double* Amat, size_t ld, const double* P, size_t ldp) {
  extern __shared__ double smem[];
  int i, j;
  tri_decode(blockIdx.x, i, j);
  double* Ct = Amat + (size_t)i * TB + (size_t)j * TB * ld;
  gemm_tile_mb<false, false, Cfg>(P + (size_t)i * TB, ldp, P + (size_t)j * TB, ldp, 0, TB, true, Ct, ld, Ct, ld, smem);
}
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
syrk_trail_kernel(double* Amat, size_t ld, const double* P, size_t ldp) {
  extern __shared__ double smem[];
  int i, j;
  tri_decode(blockIdx.x, i, j);
  double* Ct = Amat + (size_t)i * TB + (size_t)j * TB * ld;
  gemm_tile<false, false, Cfg>(P + (size_t)i * TB, ldp, P + (size_t)j * TB, ldp, 0, TB, true, Ct, ld, Ct, ld, smem);
}

static double *dA, *dB, *dC;
static cudaEvent_t e0, e1;

template <class Cfg>
void run(const char* name, int n, int reps) {
  CK(cudaFuncSetAttribute(gemm_nt_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(syrk_trail_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  float ms = 0, best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    gemm_nt_kernel<Cfg><<<dim3(n / TB, n / TB), Cfg::THREADS, Cfg::SMEM_BYTES>>>(dA, dB, dC, n, n);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 || reps == 1) best = ms < best ? ms : best;
  }
  double fl = 2.0 * n * (double)n * n;
  printf("{\"probe\": \"gemm_nt\", \"cfg\": \"%s\", \"threads\": %d, \"smem\": %d, \"n\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", name, Cfg::THREADS, Cfg::SMEM_BYTES, n, best, fl / best * 1e-9);
  int nb = n / TB;
  best = 1e30f;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaEventRecord(e0));
    syrk_trail_kernel<Cfg><<<nb * (nb + 1) / 2, Cfg::THREADS, Cfg::SMEM_BYTES>>>(dC, n, dA, n);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0) best = ms < best ? ms : best;
  }
  fl = (double)nb * (nb + 1) / 2 * 2.0 * TB * TB * TB;
  printf("{\"probe\": \"syrk_trail_k128\", \"cfg\": \"%s\", \"n\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", name, n, best, fl / best * 1e-9);
}

template <class Cfg>
void run_mb(const char* name, int n, int reps, double* hC, double* hR, size_t nn) {
  CK(cudaFuncSetAttribute(gemm_nt_mb_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(syrk_trail_mb_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(gemm_nt_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  float ms = 0, best = 1e30f;
  gemm_nt_kernel<Cfg><<<dim3(n / TB, n / TB), Cfg::THREADS, Cfg::SMEM_BYTES>>>(dA, dB, dC, n, n);
  CK(cudaMemcpy(hR, dC, nn * 8, cudaMemcpyDeviceToHost));
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    gemm_nt_mb_kernel<Cfg><<<dim3(n / TB, n / TB), Cfg::THREADS, Cfg::SMEM_BYTES>>>(dA, dB, dC, n, n);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 || reps == 1) best = ms < best ? ms : best;
  }
  CK(cudaMemcpy(hC, dC, nn * 8, cudaMemcpyDeviceToHost));
  double md = 0;
  for (size_t i = 0; i < nn; ++i) { double d = hC[i] - hR[i]; if (d < 0) d = -d; if (d > md) md = d; }
  double fl = 2.0 * n * (double)n * n;
  printf("{\"probe\": \"gemm_nt_mbarrier\", \"cfg\": \"%s\", \"n\": %d, \"ms\": %.4f, \"tflops\": %.3f, \"maxdiff_vs_sync_version\": %.3e}\n", name, n, best, fl / best * 1e-9, md);
  int nb = n / TB;
  best = 1e30f;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaEventRecord(e0));
    syrk_trail_mb_kernel<Cfg><<<nb * (nb + 1) / 2, Cfg::THREADS, Cfg::SMEM_BYTES>>>(dC, n, dA, n);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0) best = ms < best ? ms : best;
  }
  fl = (double)nb * (nb + 1) / 2 * 2.0 * TB * TB * TB;
  printf("{\"probe\": \"syrk_trail_k128_mbarrier\", \"cfg\": \"%s\", \"n\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", name, n, best, fl / best * 1e-9);
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 4096, reps = argc > 2 ? atoi(argv[2]) : 3, which = argc > 3 ? atoi(argv[3]) : -1;
  size_t nn = (size_t)n * n;
  CK(cudaMalloc(&dA, nn * 8)); CK(cudaMalloc(&dB, nn * 8)); CK(cudaMalloc(&dC, nn * 8));
  std::vector<double> h(nn);
  srand(3);
  for (size_t i = 0; i < nn; ++i) h[i] = rand() / (double)RAND_MAX - 0.5;
  CK(cudaMemcpy(dA, h.data(), nn * 8, cudaMemcpyHostToDevice));
  for (size_t i = 0; i < nn; ++i) h[i] = rand() / (double)RAND_MAX - 0.5;
  CK(cudaMemcpy(dB, h.data(), nn * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0, nn * 8));
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<double> hR(nn);
  if (which < 0 || which == 10) run_mb<GemmCfgT<2, 4, 16, 4>>("2x4_bk16_s4", n, reps, h.data(), hR.data(), nn);
  if (which < 0 || which == 11) run_mb<GemmCfgT<2, 4, 16, 5>>("2x4_bk16_s5", n, reps, h.data(), hR.data(), nn);
  if (which < 0 || which == 12) run_mb<GemmCfgT<2, 4, 32, 3>>("2x4_bk32_s3", n, reps, h.data(), hR.data(), nn);
  if (which < 0 || which == 13) run_mb<GemmCfgT<4, 4, 16, 5>>("4x4_bk16_s5", n, reps, h.data(), hR.data(), nn);
  if (which < 0 || which == 0) run<GemmCfgT<2, 4, 16, 4>>("2x4_bk16_s4", n, reps);
  if (which < 0 || which == 1) run<GemmCfgT<4, 4, 16, 4>>("4x4_bk16_s4", n, reps);
  if (which < 0 || which == 2) run<GemmCfgT<4, 4, 32, 3>>("4x4_bk32_s3", n, reps);
  if (which < 0 || which == 3) run<GemmCfgT<2, 4, 32, 3>>("2x4_bk32_s3", n, reps);
  if (which < 0 || which == 4) run<GemmCfgT<2, 8, 16, 4>>("2x8_bk16_s4", n, reps);
  if (which < 0 || which == 5) run<GemmCfgT<4, 2, 16, 4>>("4x2_bk16_s4", n, reps);
  if (which < 0 || which == 6) run<GemmCfgT<4, 4, 16, 5>>("4x4_bk16_s5", n, reps);
  if (which < 0 || which == 7) run<GemmCfgT<4, 8, 16, 4>>("4x8_bk16_s4", n, reps);
  return 0;
}
