// FP64 probe for B200: measures the denominators the roofline fractions need
// (DMMA issue peak, DFMA issue peak, cuBLAS DGEMM as an external comparator) and
// checks/times the hand-written DMMA tile contraction in gpedm_b200/csrc/dgemm_tile.cuh.
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o build/fp64_probe tools/fp64_probe.cu -lcublas
// This is a measurement tool, not part of the product library.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../gpedm_b200/csrc/dgemm_tile.cuh"

using namespace gpedm;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
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

__global__ void __launch_bounds__(256) dmma1688_peak_kernel(double* out, int iters) {
  double c[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 * 1.1, a2 = a0 * 1.2, a3 = a0 * 1.3;
  double b0 = 1.0 - threadIdx.x * 1e-9, b1 = b0 * 0.9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile(
          "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
          : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
          : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
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

// Generic full GEMM over 128x128 tiles for the three operand-layout combinations in use.
template <bool AK, bool BK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_full_kernel(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, int K,
                     int accumulate, int negate) {
  extern __shared__ double smem[];
  int ti = blockIdx.x, tj = blockIdx.y;
  const double* At = AK ? A + (size_t)ti * TB * lda : A + (size_t)ti * TB;
  const double* Bt = BK ? B + (size_t)tj * TB * ldb : B + (size_t)tj * TB;
  double* Ct = C + (size_t)ti * TB + (size_t)tj * TB * ldc;
  gemm_tile<AK, BK>(At, lda, Bt, ldb, 0, K, negate != 0, accumulate ? Ct : nullptr, ldc, Ct, ldc, smem);
}

// Trailing-update shaped launch: lower tiles (i >= j) of an nb x nb tile grid, K = 128.
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    syrk_trail_kernel(double* Amat, size_t ld, const double* P, size_t ldp) {
  extern __shared__ double smem[];
  int i, j;
  tri_decode(blockIdx.x, i, j);
  const double* At = P + (size_t)i * TB;
  const double* Bt = P + (size_t)j * TB;
  double* Ct = Amat + (size_t)i * TB + (size_t)j * TB * ld;
  gemm_tile<false, false>(At, ldp, Bt, ldp, 0, TB, true, Ct, ld, Ct, ld, smem);
}

static double maxdiff(const std::vector<double>& a, const std::vector<double>& b) {
  double m = 0;
  for (size_t i = 0; i < a.size(); ++i) m = fmax(m, fabs(a[i] - b[i]));
  return m;
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  int nsm = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"}\n", prop.name, nsm, prop.major, prop.minor);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms;
  double* dout;
  CK(cudaMalloc(&dout, sizeof(double) * nsm * 8 * 256));

  // ---- issue peaks -------------------------------------------------------------
  for (int cta_per_sm = 1; cta_per_sm <= 2; ++cta_per_sm) {
    int iters = 20000;
    dmma_peak_kernel<<<nsm * cta_per_sm, 256>>>(dout, 100);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dmma_peak_kernel<<<nsm * cta_per_sm, 256>>>(dout, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = (double)nsm * cta_per_sm * 8 * 32.0 * iters * 512.0;
    printf("{\"probe\": \"dmma_m8n8k4_peak\", \"cta_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.3f}\n", cta_per_sm, ms,
           fl / ms * 1e-9);
  }
  {
    int iters = 20000;
    dmma1688_peak_kernel<<<nsm, 256>>>(dout, 100);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dmma1688_peak_kernel<<<nsm, 256>>>(dout, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = (double)nsm * 8 * 16.0 * iters * 2048.0;
    printf("{\"probe\": \"dmma_m16n8k8_peak\", \"ms\": %.3f, \"tflops\": %.3f}\n", ms, fl / ms * 1e-9);
  }
  for (int cta_per_sm = 1; cta_per_sm <= 4; cta_per_sm *= 2) {
    int iters = 20000;
    dfma_peak_kernel<<<nsm * cta_per_sm, 256>>>(dout, 100);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dfma_peak_kernel<<<nsm * cta_per_sm, 256>>>(dout, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = (double)nsm * cta_per_sm * 256 * 16.0 * iters * 2.0;
    printf("{\"probe\": \"dfma_peak\", \"cta_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.3f}\n", cta_per_sm, ms,
           fl / ms * 1e-9);
  }

  // ---- tile GEMM vs cuBLAS -----------------------------------------------------
  cublasHandle_t hb;
  cublasCreate(&hb);
  const double one = 1.0, zero = 0.0, mone = -1.0;
  CK(cudaFuncSetAttribute(gemm_full_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CK(cudaFuncSetAttribute(gemm_full_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CK(cudaFuncSetAttribute(gemm_full_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CK(cudaFuncSetAttribute(syrk_trail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));

  for (int n : {512, 4096, 8192}) {
    size_t nn = (size_t)n * n;
    std::vector<double> hA(nn), hB(nn), hC(nn), hR(nn);
    srand(1);
    for (size_t i = 0; i < nn; ++i) {
      hA[i] = (rand() / (double)RAND_MAX) - 0.5;
      hB[i] = (rand() / (double)RAND_MAX) - 0.5;
    }
    double *dA, *dB, *dC, *dR;
    CK(cudaMalloc(&dA, nn * 8));
    CK(cudaMalloc(&dB, nn * 8));
    CK(cudaMalloc(&dC, nn * 8));
    CK(cudaMalloc(&dR, nn * 8));
    CK(cudaMemcpy(dA, hA.data(), nn * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), nn * 8, cudaMemcpyHostToDevice));
    dim3 grid(n / TB, n / TB);
    double flops = 2.0 * n * (double)n * n;
    for (int combo = 0; combo < 3; ++combo) {
      // combo 0: C = A * B^T (A MN-major, B MN-major)   cuBLAS N,T
      // combo 1: C = A * B   (A MN-major, B K-major)    cuBLAS N,N
      // combo 2: C = A^T * B (A K-major,  B K-major)    cuBLAS T,N
      cublasOperation_t opa = combo == 2 ? CUBLAS_OP_T : CUBLAS_OP_N;
      cublasOperation_t opb = combo == 0 ? CUBLAS_OP_T : CUBLAS_OP_N;
      cublasDgemm(hb, opa, opb, n, n, n, &one, dA, n, dB, n, &zero, dR, n);
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        if (combo == 0)
          gemm_full_kernel<false, false><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES>>>(dA, n, dB, n, dC, n, n, 0, 0);
        else if (combo == 1)
          gemm_full_kernel<false, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES>>>(dA, n, dB, n, dC, n, n, 0, 0);
        else
          gemm_full_kernel<true, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES>>>(dA, n, dB, n, dC, n, n, 0, 0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      CK(cudaMemcpy(hC.data(), dC, nn * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hR.data(), dR, nn * 8, cudaMemcpyDeviceToHost));
      printf("{\"probe\": \"tile_gemm\", \"n\": %d, \"combo\": %d, \"ms\": %.3f, \"tflops\": %.3f, \"maxdiff_vs_cublas\": %.3e}\n",
             n, combo, ms, flops / ms * 1e-9, maxdiff(hC, hR));
    }
    // cuBLAS timing (N,T)
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, dA, n, dB, n, &zero, dR, n);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("{\"probe\": \"cublas_dgemm\", \"n\": %d, \"ms\": %.3f, \"tflops\": %.3f}\n", n, ms, flops / ms * 1e-9);
    // negate + accumulate path check: C = C - A*B^T  vs cuBLAS
    {
      CK(cudaMemcpy(dC, hB.data(), nn * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dR, hB.data(), nn * 8, cudaMemcpyHostToDevice));
      gemm_full_kernel<false, false><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES>>>(dA, n, dB, n, dC, n, n, 1, 1);
      cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &mone, dA, n, dB, n, &one, dR, n);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hC.data(), dC, nn * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hR.data(), dR, nn * 8, cudaMemcpyDeviceToHost));
      printf("{\"probe\": \"tile_gemm_negacc\", \"n\": %d, \"maxdiff_vs_cublas\": %.3e}\n", n, maxdiff(hC, hR));
    }
    // trailing-update shape: K=128, lower tiles only
    if (n >= 4096) {
      int nb = n / TB;
      int ntiles = nb * (nb + 1) / 2;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        syrk_trail_kernel<<<ntiles, GEMM_THREADS, GEMM_SMEM_BYTES>>>(dC, n, dA, n);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      double fl = (double)ntiles * 2.0 * TB * TB * TB;
      printf("{\"probe\": \"syrk_trail_k128\", \"n\": %d, \"tiles\": %d, \"ms\": %.3f, \"tflops\": %.3f}\n", n, ntiles, ms,
             fl / ms * 1e-9);
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    cudaFree(dR);
  }
  // sustained DGEMM for ~2 s to see the power-capped rate
  {
    int n = 8192;
    size_t nn = (size_t)n * n;
    double *dA, *dB, *dC;
    CK(cudaMalloc(&dA, nn * 8));
    CK(cudaMalloc(&dB, nn * 8));
    CK(cudaMalloc(&dC, nn * 8));
    CK(cudaMemset(dA, 0, nn * 8));
    CK(cudaMemset(dB, 0, nn * 8));
    dim3 grid(n / TB, n / TB);
    CK(cudaEventRecord(e0));
    int reps = 40;
    for (int r = 0; r < reps; ++r)
      gemm_full_kernel<false, false><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES>>>(dA, n, dB, n, dC, n, n, 0, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"probe\": \"tile_gemm_sustained\", \"n\": %d, \"reps\": %d, \"ms\": %.3f, \"tflops\": %.3f}\n", n, reps, ms,
           reps * 2.0 * n * (double)n * n / ms * 1e-9);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, dA, n, dB, n, &zero, dC, n);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"probe\": \"cublas_dgemm_sustained\", \"n\": %d, \"reps\": %d, \"ms\": %.3f, \"tflops\": %.3f}\n", n, reps, ms,
           reps * 2.0 * n * (double)n * n / ms * 1e-9);
  }
  return 0;
}
