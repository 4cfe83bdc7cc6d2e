// libgpedm_b200 host orchestration - errors, environment knobs and RAII helpers shared by the host orchestration.
// Part of the single translation unit gpedm_core.cu (included in order; not a standalone header).
#pragma once

// ------------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return set_err(GPEDM_ERR_CUDA, "CUDA error '%s' at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, \
                     __LINE__, #call);                                                               \
  } while (0)

// Reference bounds (R/GPfunctions.R:591-593)
static const double VEMIN = 0.0001, SIGMA2MIN = 0.0001, VEMAX = 4.9999, SIGMA2MAX = 4.9999, RHOMIN = 0.0001,
                    RHOMAX = 1.0;
static inline double logit3(double x, double lo, double hi) { return log((x - lo) / (hi - x)); }

// ------------------------------------------------------------------------------- RAII buffers
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
};
// Stream-ordered allocation from the device's (cached) default pool: no implicit device
// synchronisation, unlike cudaMalloc / cudaFree.  The stream must outlive the buffer.
struct PoolBuf {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  explicit PoolBuf(cudaStream_t s = nullptr) : st(s) {}
  ~PoolBuf() { if (p) cudaFreeAsync(p, st); }
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, std::max<size_t>(bytes, 16), st); }
};
struct StreamGuard {
  cudaStream_t s = nullptr;
  cudaError_t create() { return cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); }
  ~StreamGuard() {
    if (s) {
      cudaStreamSynchronize(s);
      cudaStreamDestroy(s);
    }
  }
};
struct PinnedBuf {
  void* p = nullptr;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  cudaError_t alloc(size_t bytes) { return cudaMallocHost(&p, std::max<size_t>(bytes, 16)); }
};
struct EventGuard {
  cudaEvent_t e = nullptr;
  cudaError_t create() { return cudaEventCreate(&e); }
  ~EventGuard() { if (e) cudaEventDestroy(e); }
};
// NVTX phase ranges: host-side ranges around the enqueue of each phase of an
// evaluation and around the public entry points; profilers attribute the launches inside to them.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  static void next(const char* name) {  // close the current phase, open the next
    nvtxRangePop();
    nvtxRangePushA(name);
  }
};
