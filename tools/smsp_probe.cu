// Which warps of a 256-thread CTA share a DMMA (FP64 tensor) pipe?  Runs a fixed number of independent
// mma.sync.m8n8k4.f64 per active warp for several warp masks and prints cycles; masks whose warps share a
// sub-partition take proportionally longer.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/smsp_probe tools/smsp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void probe(unsigned mask, int iters, long long* out, double* sink) {
  const int warp = threadIdx.x >> 5;
  double c[16][2];
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  __syncthreads();
  long long t0 = clock64();
  if ((mask >> warp) & 1u) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a, b);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
int main() {
  long long* d;
  double* sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 8);
  const unsigned masks[] = {0x01, 0x03, 0x05, 0x11, 0x09, 0x0f, 0x55, 0x33, 0xf0, 0xff, 0x81, 0x21, 0x41};
  const int iters = 2000;
  for (unsigned m : masks) {
    long long h = 0;
    probe<<<1, 256>>>(m, iters, d, sink);
    probe<<<1, 256>>>(m, iters, d, sink);
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    int nw = __builtin_popcount(m);
    printf("{\"mask\": \"0x%02x\", \"warps\": %d, \"cycles\": %lld, \"cycles_per_dmma_per_warp\": %.2f, \"sm_cycles_per_dmma\": %.2f}\n", m, nw, h,
           (double)h / (iters * 16), (double)h / (iters * 16 * nw));
  }
  return 0;
}
