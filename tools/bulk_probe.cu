// How should the k slabs of the DMMA tile contraction reach shared memory?  A long-k 128 x 128 tile kernel
// (8 warps of 64 x 32, 4-stage ring of 16-wide slabs, mbarrier hand-off - the k loop of dag_factor.cuh) with
// three loaders:
//   mode 0  every thread issues its 16 LDGSTS (cp.async 16 B) per slab + cp.async.mbarrier.arrive.noinc   (the tree)
//   mode 1  bulk copies (cp.async.bulk, the TMA unit, no tensor map): one elected lane per row of the slab, the
//           padded smem layout kept; MN-major operands = 16 rows x 1 KB per operand and slab, K-major = 128 rows x 128 B
// for both operand layouts (NT = both MN-major as in the Cholesky tasks, TN = both K-major as in X^T X).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o build/bulk_probe tools/bulk_probe.cu
//   build/bulk_probe [k=8192] [tiles_per_sm=4]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../gpedm_b200/csrc/dgemm_tile.cuh"
using namespace gpedm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
using Cfg = GemmCfgT<2, 4, 16, 4>;

__device__ __forceinline__ void mbar_wait_a(uint32_t a, unsigned parity) {
  asm volatile("{\n .reg .pred p;\nWAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra WAIT_%=;\n}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t a) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t a, unsigned bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, unsigned bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int MODE, bool KMAJ>
__global__ void __launch_bounds__(256, 1) probe_kernel(const double* A, const double* B, double* C, size_t ld, int K, int mt) {
  extern __shared__ double smem_raw[];
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_raw);
  asm volatile("" : "+r"(base));
  double* const smem = (double*)__cvta_shared_to_generic(base);
  constexpr int MT = Cfg::MT, NT = Cfg::NT, BK = Cfg::BK, NS = Cfg::NSTAGE, PREF = NS - 2;
  constexpr int SLAB_A = Cfg::SLAB_A, SLAB_B = Cfg::SLAB_B, LD_A = Cfg::LD_MN_A, LD_B = Cfg::LD_MN_B, LD_K = Cfg::LD_K;
  constexpr uint32_t BAR = (uint32_t)(NS * (SLAB_A + SLAB_B) * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp % Cfg::WM) * (MT * 8), n0 = (warp / Cfg::WM) * (NT * 8);
  const int bi = blockIdx.x % mt, bj = blockIdx.x / mt;
  // KMAJ: element (m, k) of the operand tile at P[k + m * ld]; else at P[m + k * ld]
  const double* Ap = KMAJ ? A + (size_t)bi * TB * ld : A + (size_t)bi * TB;
  const double* Bp = KMAJ ? B + (size_t)bj * TB * ld : B + (size_t)bj * TB;
  double* sA = smem;
  double* sB = smem + NS * SLAB_A;
  auto full = [&](unsigned s) { return base + BAR + 8u * s; };
  auto empty = [&](unsigned s) { return base + BAR + 8u * (NS + s); };
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(full(s)), "r"(MODE == 0 ? 256 : 8) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(empty(s)), "r"(8) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const int nk = K / BK;
  auto issue = [&](int j) {
    const unsigned s = j % NS, r = j / NS;
    const uint32_t dA = base + s * (uint32_t)(SLAB_A * 8), dB = base + (uint32_t)(NS * SLAB_A * 8) + s * (uint32_t)(SLAB_B * 8);
    const int k0 = j * BK;
    if (MODE == 0) {
      if (r > 0) mbar_wait_a(empty(s), (r - 1) & 1u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = tid + q * 256;
        uint32_t da, db;
        const double *sa, *sb;
        if (!KMAJ) {
          const int row = chunk >> 6, c2 = (chunk & 63) << 1;
          da = dA + (uint32_t)(row * LD_A + c2) * 8u;
          db = dB + (uint32_t)(row * LD_B + c2) * 8u;
          sa = Ap + (size_t)(k0 + row) * ld + c2;
          sb = Bp + (size_t)(k0 + row) * ld + c2;
        } else {
          const int row = chunk >> 3, c2 = (chunk & 7) << 1;
          da = dA + (uint32_t)(row * LD_K + c2) * 8u;
          db = dB + (uint32_t)(row * LD_K + c2) * 8u;
          sa = Ap + (size_t)row * ld + k0 + c2;
          sb = Bp + (size_t)row * ld + k0 + c2;
        }
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(da), "l"(sa) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(db), "l"(sb) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(full(s)) : "memory");
    } else if (!KMAJ) {
      // warp w copies k rows 2w, 2w+1 of both operands: lanes 0..3, 1 KB each
      if (lane < 4) {
        if (r > 0) mbar_wait_a(empty(s), (r - 1) & 1u);
        if (lane == 0) mbar_expect_tx_a(full(s), 4096u);
        const int row = 2 * warp + (lane & 1);
        const bool isB = lane >> 1;
        bulk_g2s((isB ? dB : dA) + (uint32_t)(row * (isB ? LD_B : LD_A)) * 8u, (isB ? Bp : Ap) + (size_t)(k0 + row) * ld, 1024u, full(s));
      }
    } else {
      // warp w copies rows 16w .. 16w+15 of A (lanes 0..15) and of B (lanes 16..31), 128 B each
      if (r > 0) mbar_wait_a(empty(s), (r - 1) & 1u);
      if (lane == 0) mbar_expect_tx_a(full(s), 4096u);
      const int row = 16 * warp + (lane & 15);
      const bool isB = lane >> 4;
      bulk_g2s((isB ? dB : dA) + (uint32_t)(row * LD_K) * 8u, (isB ? Bp : Ap) + (size_t)row * ld + k0, 128u, full(s));
    }
  };
  for (int j = 0; j < PREF; ++j)
    if (j < nk) issue(j);
  double acc[MT][NT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + PREF < nk) issue(kt + PREF);
    const unsigned s = kt % NS;
    mbar_wait_a(full(s), (kt / NS) & 1u);
    const double* a_s = sA + s * SLAB_A;
    const double* b_s = sB + s * SLAB_B;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double a[MT], b[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) a[i] = KMAJ ? a_s[(m0 + 8 * i + g) * LD_K + k4 * 4 + t] : a_s[(k4 * 4 + t) * LD_A + m0 + 8 * i + g];
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = KMAJ ? b_s[(n0 + 8 * j + g) * LD_K + k4 * 4 + t] : b_s[(k4 * 4 + t) * LD_B + n0 + 8 * j + g];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_a(empty(s));
  }
  double* Ct = C + (size_t)bi * TB + (size_t)bj * TB * ld;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) Ct[(size_t)(m0 + 8 * i + g) + (size_t)(n0 + 8 * j + 2 * t + e) * ld] = acc[i][j][e];
}

static double *dA, *dB, *dC;
template <int MODE, bool KMAJ>
double run(const char* name, size_t ld, int K, int mt, int nt, std::vector<double>& out) {
  const int smem = Cfg::SMEM_BYTES + 64;
  CK(cudaFuncSetAttribute(probe_kernel<MODE, KMAJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f, ms;
  for (int r = 0; r < 4; ++r) {
    CK(cudaEventRecord(e0));
    probe_kernel<MODE, KMAJ><<<mt * nt, 256, smem>>>(dA, dB, dC, ld, K, mt);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best) best = ms;
  }
  out.resize((size_t)TB * mt * TB);
  CK(cudaMemcpy(out.data(), dC, out.size() * 8, cudaMemcpyDeviceToHost));  // first block column of C
  const double fl = 2.0 * TB * TB * (double)K * mt * nt;
  printf("{\"probe\": \"%s\", \"layout\": \"%s\", \"K\": %d, \"tiles\": %d, \"ms\": %.4f, \"tflops\": %.3f, \"us_per_k_tile\": %.3f}\n", name,
         KMAJ ? "TN (K-major)" : "NT (MN-major)", K, mt * nt, best, fl / best * 1e-9, best * 1e3 / (K / TB) / ((mt * nt + 147) / 148));
  return fl / best * 1e-9;
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 8192, per_sm = argc > 2 ? atoi(argv[2]) : 4;
  const int mt = 37, nt = 4 * per_sm;  // 148 x per_sm tiles: whole waves
  const size_t ld = (size_t)std::max(K, std::max(mt, nt) * TB);
  const size_t nn = ld * ld;
  CK(cudaMalloc(&dA, nn * 8));
  CK(cudaMalloc(&dB, nn * 8));
  CK(cudaMalloc(&dC, nn * 8));
  std::vector<double> h(nn);
  srand(3);
  for (size_t i = 0; i < nn; ++i) h[i] = rand() / (double)RAND_MAX - 0.5;
  CK(cudaMemcpy(dA, h.data(), nn * 8, cudaMemcpyHostToDevice));
  for (size_t i = 0; i < nn; ++i) h[i] = rand() / (double)RAND_MAX - 0.5;
  CK(cudaMemcpy(dB, h.data(), nn * 8, cudaMemcpyHostToDevice));
  std::vector<double> r0, r1;
  run<0, false>("ldgsts", ld, K, mt, nt, r0);
  run<1, false>("bulk", ld, K, mt, nt, r1);
  double md = 0;
  for (size_t i = 0; i < r0.size(); ++i) md = std::max(md, fabs(r0[i] - r1[i]));
  printf("{\"check\": \"NT bulk vs ldgsts\", \"maxdiff\": %.3e}\n", md);
  run<0, true>("ldgsts", ld, K, mt, nt, r0);
  run<1, true>("bulk", ld, K, mt, nt, r1);
  md = 0;
  for (size_t i = 0; i < r0.size(); ++i) md = std::max(md, fabs(r0[i] - r1[i]));
  printf("{\"check\": \"TN bulk vs ldgsts\", \"maxdiff\": %.3e}\n", md);
  return 0;
}
