// libgpedm_b200: host orchestration + C ABI (include/gpedm.h) for the B200-native GP
// evaluation pipeline.  The scalar parts of getlikegrad (parameter transforms, priors,
// chain rule; reference R/GPfunctions.R:590-630, 719-725, 752-777) and the Rprop loop
// (:522-584) run on the host; everything that touches N x N data runs in the kernels of
// kernels.cuh.  There is no CPU fallback for any N^2 / N^3 step.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gpedm.h"
#include "kernels.cuh"
#include "diag_block.cuh"
#include "small_fit.cuh"
#include "lag_prep.cuh"

using namespace gpedm;

static_assert(MAXD == GPEDM_MAX_D, "MAXD mismatch");

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

// ------------------------------------------------------------------------------- handle
struct gpedm_fit_s {
  int device = 0;
  int64_t N = 0;
  int Np = 0, nb = 0, ntiles = 0;
  int d = 0, np = 1, p = 0;
  bool has_rhomat = false;
  bool single_pop = true;
  double *dX = nullptr, *dY = nullptr, *drhotab = nullptr;
  int* dpop = nullptr;
  double *bufA = nullptr, *bufB = nullptr, *bufC = nullptr;
  EvalParams* dparams = nullptr;
  void* pinned = nullptr;         // one recycled pinned block holding hparams | hscal | hinfo
  EvalParams* hparams = nullptr;  // pinned
  double *dz = nullptr, *da = nullptr, *ddiagS = nullptr, *dpart = nullptr, *dlogdet = nullptr,
         *dtracepart = nullptr, *dscal = nullptr;
  double* hscal = nullptr;  // pinned
  int* dinfo = nullptr;
  int* hinfo = nullptr;  // pinned
  int num_sms = 148;
  cudaGraphExec_t graph_exec[3] = {nullptr, nullptr, nullptr};  // gradient-only / full / through-z variants
  int64_t graph_launches[3] = {0, 0, 0};
  const double* dsigma_in = nullptr;  // gpedm_covinv: factor this N x N matrix instead of assembling Sigma
  int trtri_w_done[16] = {}, trtri_x_done[16] = {};  // per level: full pairs already launched this evaluation
  cudaStream_t stream = nullptr;
  cudaStream_t aux = nullptr;  // high-priority stream for the look-ahead chain (diag, panel, next column)
  cudaStream_t fill = nullptr; // low-priority stream: early triangular-inverse levels filling the Cholesky tail
  cudaEvent_t evFill = nullptr;
  std::vector<cudaEvent_t> evCol;  // "block columns < c are final" checkpoints for the fill stream
  std::vector<cudaEvent_t> evPanel, evTrail;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  cudaEvent_t ev[GPEDM_NPHASE + 1] = {};
  float phase_ms[GPEDM_NPHASE] = {};
  int64_t launches = 0;
  bool have_full = false;  // L (bufA), L^-1 (bufB), Sigma^-1 (bufC), a, diagS valid for `cur`
  double cur_ve = 0, cur_sigma2 = 0, cur_rho = 0;
  std::vector<double> hY;  // host copy of Y (N)
  std::vector<int> hpop;
  // set by gpedm_create_lagged: column statistics (E+1) x G used for the scaling, kept rows
  double *dcenter = nullptr, *dscale = nullptr;
  int sc_mode = 0, sc_groups = 1;
  std::vector<int> rows;  // original row of each training row
};

static const int NSCAL = 8 + MAXD + 3;

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

static bool g_attr_set[64] = {};
static int g_num_sms[64] = {};
// GPEDM_DIAG_REF=1 selects the simple (slow) reference diagonal-block kernel, for A/B debugging.
static const bool g_no_lookahead = getenv("GPEDM_NO_LOOKAHEAD") != nullptr && atoi(getenv("GPEDM_NO_LOOKAHEAD")) != 0;
static const int g_group = getenv("GPEDM_POTRF_GROUP") ? std::max(1, atoi(getenv("GPEDM_POTRF_GROUP"))) : 0;
static const int g_group_min = getenv("GPEDM_POTRF_GROUP_MIN") ? atoi(getenv("GPEDM_POTRF_GROUP_MIN")) : 16;
static const bool g_no_half_tiles = getenv("GPEDM_NO_HALF_TILES") != nullptr && atoi(getenv("GPEDM_NO_HALF_TILES")) != 0;
static const int g_fill_tail = getenv("GPEDM_FILL_TAIL") ? atoi(getenv("GPEDM_FILL_TAIL")) : 40;
static const bool g_no_small_batch = getenv("GPEDM_NO_SMALL_BATCH") != nullptr && atoi(getenv("GPEDM_NO_SMALL_BATCH")) != 0;
static const bool g_no_fill = getenv("GPEDM_NO_FILL") != nullptr && atoi(getenv("GPEDM_NO_FILL")) != 0;
static const bool g_no_split = getenv("GPEDM_NO_SPLIT") != nullptr && atoi(getenv("GPEDM_NO_SPLIT")) != 0;
static const int g_graph_max_nb = getenv("GPEDM_GRAPH_MAX_NB") ? atoi(getenv("GPEDM_GRAPH_MAX_NB")) : 48;
static const int g_batch_workers = getenv("GPEDM_BATCH_WORKERS_PER_GPU") ? atoi(getenv("GPEDM_BATCH_WORKERS_PER_GPU")) : 0;
static const bool g_no_graph = getenv("GPEDM_NO_GRAPH") != nullptr && atoi(getenv("GPEDM_NO_GRAPH")) != 0;
static const bool g_diag_ref = getenv("GPEDM_DIAG_REF") != nullptr && atoi(getenv("GPEDM_DIAG_REF")) != 0;
static std::mutex g_attr_mutex;
static int ensure_kernel_attrs(int device) {
  std::lock_guard<std::mutex> lock(g_attr_mutex);  // batch workers create handles concurrently
  if (device < 64 && g_attr_set[device]) return 0;
  CU(cudaFuncSetAttribute(tile_once_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CU(cudaFuncSetAttribute(tile_once_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CU(cudaFuncSetAttribute(tile_once_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CU(cudaFuncSetAttribute(tile_once_kernel<false, false, GemmCfgSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          GemmCfgSplit::SMEM_BYTES));
  CU(cudaFuncSetAttribute(tile_once_kernel<false, false, GemmCfgHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          GemmCfgHalf::SMEM_BYTES));
  CU(cudaFuncSetAttribute(tile_once_kernel<false, true, GemmCfgHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          GemmCfgHalf::SMEM_BYTES));
  CU(cudaFuncSetAttribute(tile_once_kernel<true, true, GemmCfgHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          GemmCfgHalf::SMEM_BYTES));
  {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (device < 64) g_num_sms[device] = prop.multiProcessorCount;
    // Handles are created and destroyed per fit (40 per model-selection grid): keep freed device
    // memory cached in the stream-ordered pool instead of returning 1.6 GB to the driver each time.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
  }
  CU(cudaFuncSetAttribute(diag_potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES));
  CU(cudaFuncSetAttribute(diag_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM_BYTES));
  CU(cudaFuncSetAttribute(small_eval_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_eval_smem_bytes<8>()));
  CU(cudaFuncSetAttribute(small_eval_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_eval_smem_bytes<16>()));
  CU(cudaFuncSetAttribute(small_eval_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_eval_smem_bytes<32>()));
  int trace_smem = (4 * MAXD * TB + 2 * TB + 8 * (MAXD + 3)) * 8 + 2 * TB * 4;
  CU(cudaFuncSetAttribute(grad_trace_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, trace_smem));
  CU(cudaFuncSetAttribute(grad_trace_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, trace_smem));
  CU(cudaFuncSetAttribute(grad_trace_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, trace_smem));
  int asm_smem = 2 * MAXD * TB * 8 + 2 * TB * 4;
  CU(cudaFuncSetAttribute(assemble_sigma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, asm_smem));
  CU(cudaFuncSetAttribute(cov_rect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, asm_smem));
  if (device < 64) g_attr_set[device] = true;
  return 0;
}

extern "C" int gpedm_abi_version(void) { return GPEDM_ABI_VERSION; }
extern "C" const char* gpedm_last_error(void) { return g_err; }
extern "C" int gpedm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
extern "C" void gpedm_default_rprop_options(gpedm_rprop_options* o) {
  o->delta0 = 0.1;
  o->deltamin = 1e-6;
  o->deltamax = 50.0;
  o->eta_minus = 0.5;
  o->eta_plus = 1.2;
  o->maxiter = 200;
  o->tol_grad = 1e-4;
  o->tol_df = 1e-7;
}

// Pinned host blocks (parameters up, scalars / status down) are recycled through a process-wide
// free list: cudaMallocHost / cudaFreeHost cost up to 200 ms each (measured) when handles are created
// and destroyed per fit, as a model-selection grid does.  Blocks are a few KB and never returned.
static const size_t PINNED_BLOCK = 4096;
static_assert(sizeof(EvalParams) + (8 + MAXD + 3) * 8 + 16 <= PINNED_BLOCK, "pinned block too small");
static std::mutex g_pinned_mutex;
static std::vector<void*> g_pinned_free;
static void* pinned_acquire() {
  {
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    if (!g_pinned_free.empty()) {
      void* p = g_pinned_free.back();
      g_pinned_free.pop_back();
      return p;
    }
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, PINNED_BLOCK, cudaHostAllocPortable) != cudaSuccess) return nullptr;
  return p;
}
static void pinned_release(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_pinned_mutex);
  g_pinned_free.push_back(p);
}

static void free_handle(gpedm_fit_s* h) {
  if (!h) return;
  static const bool trace = getenv("GPEDM_TRACE_FREE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = now();
  cudaSetDevice(h->device);
  if (h->stream) {
    cudaStreamSynchronize(h->stream);
    if (h->aux) cudaStreamSynchronize(h->aux);
    void* bufs[] = {h->dX, h->dY, h->drhotab, h->dpop, h->bufA, h->bufB, h->bufC, h->dparams, h->dz, h->da,
                    h->ddiagS, h->dpart, h->dlogdet, h->dtracepart, h->dscal, h->dinfo, h->dcenter, h->dscale};
    for (void* b : bufs)
      if (b) cudaFreeAsync(b, h->stream);
    cudaStreamSynchronize(h->stream);
  }
  double t1 = now();
  pinned_release(h->pinned);
  double t2 = now();
  for (auto& g : h->graph_exec)
    if (g) cudaGraphExecDestroy(g);
  double t3 = now();
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->evPanel) cudaEventDestroy(e);
  for (auto& e : h->evTrail) cudaEventDestroy(e);
  for (auto& e : h->evCol) cudaEventDestroy(e);
  if (h->evFill) cudaEventDestroy(h->evFill);
  if (h->evFork) cudaEventDestroy(h->evFork);
  if (h->evJoin) cudaEventDestroy(h->evJoin);
  double t4 = now();
  if (h->fill) cudaStreamDestroy(h->fill);
  if (h->aux) cudaStreamDestroy(h->aux);
  if (h->stream) cudaStreamDestroy(h->stream);
  double t5 = now();
  if (trace)
    fprintf(stderr, "[gpedm free] device bufs %.2f ms, pinned %.2f, graphs %.2f, events %.2f, streams %.2f\n", t1 - t0,
            t2 - t1, t3 - t2, t4 - t3, t5 - t4);
  delete h;
}

// Allocate a handle for an N x d problem with np populations: device buffers, streams, events.
// The training data (dX, dY, dpop, drhotab, hY, hpop, single_pop) are filled by the caller.
static int alloc_handle(int device, int64_t N, int32_t d, int32_t np, bool has_rhomat, gpedm_handle* out) {
  *out = nullptr;
  if (N < 1 || N > (1 << 20)) return set_err(GPEDM_ERR_ARG, "N=%lld out of range", (long long)N);
  if (d < 1 || d > GPEDM_MAX_D) return set_err(GPEDM_ERR_ARG, "d=%d out of range 1..%d", d, GPEDM_MAX_D);
  if (np < 1) return set_err(GPEDM_ERR_ARG, "np=%d must be >= 1", np);
  int ndev = gpedm_device_count();
  if (ndev == 0) return set_err(GPEDM_ERR_CUDA, "no CUDA device available (libgpedm_b200 has no CPU fallback)");
  if (device < 0 || device >= ndev) return set_err(GPEDM_ERR_ARG, "device %d out of range (have %d)", device, ndev);
  CU(cudaSetDevice(device));
  int rc = ensure_kernel_attrs(device);
  if (rc) return rc;
  gpedm_fit_s* h = new gpedm_fit_s();
  h->device = device;
  h->N = N;
  h->nb = ceil_div((int)N, TB);
  h->Np = h->nb * TB;
  h->ntiles = h->nb * (h->nb + 1) / 2;
  h->d = d;
  h->np = np;
  h->p = d + 3;
  h->has_rhomat = has_rhomat;
  const size_t Np = h->Np, mat = Np * Np * sizeof(double);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    free_handle(h);
    return set_err(GPEDM_ERR_CUDA, "stream creation failed");
  }
#define ALLOC(ptr, bytes)                                                                         \
  do {                                                                                            \
    cudaError_t e_ = cudaMallocAsync((void**)&(ptr), (bytes), h->stream);                         \
    if (e_ != cudaSuccess) {                                                                      \
      free_handle(h);                                                                             \
      return set_err(GPEDM_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", (size_t)(bytes),       \
                     cudaGetErrorString(e_));                                                     \
    }                                                                                             \
  } while (0)
  ALLOC(h->dX, (size_t)N * d * 8);
  ALLOC(h->dY, Np * 8);
  ALLOC(h->dpop, (size_t)N * 4);
  if (has_rhomat) ALLOC(h->drhotab, (size_t)np * np * 8);
  ALLOC(h->bufA, mat);
  ALLOC(h->bufB, mat);
  ALLOC(h->bufC, mat);
  ALLOC(h->dparams, sizeof(EvalParams));
  ALLOC(h->dz, Np * 8);
  ALLOC(h->da, Np * 8);
  ALLOC(h->ddiagS, Np * 8);
  ALLOC(h->dpart, (size_t)h->nb * Np * 8);
  ALLOC(h->dlogdet, (size_t)h->nb * 8);
  ALLOC(h->dtracepart, (size_t)h->ntiles * (MAXD + 3) * 8);
  ALLOC(h->dscal, NSCAL * 8);
  ALLOC(h->dinfo, 4);
  h->num_sms = (device < 64 && g_num_sms[device] > 0) ? g_num_sms[device] : 148;
#undef ALLOC
  h->pinned = pinned_acquire();
  if (!h->pinned) {
    free_handle(h);
    return set_err(GPEDM_ERR_CUDA, "pinned host allocation failed");
  }
  h->hparams = (EvalParams*)h->pinned;
  h->hscal = (double*)((char*)h->pinned + ((sizeof(EvalParams) + 15) / 16) * 16);
  h->hinfo = (int*)(h->hscal + NSCAL);
  cudaError_t e = cudaMemsetAsync(h->dY, 0, Np * 8, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    e = cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking, hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->fill, cudaStreamNonBlocking, lo);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->evFill, cudaEventDisableTiming);
  h->evCol.resize(h->nb / 8 + 1, nullptr);
  for (size_t i = 0; i < h->evCol.size() && e == cudaSuccess; ++i)
    e = cudaEventCreateWithFlags(&h->evCol[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming);
  h->evPanel.resize(h->nb, nullptr);
  h->evTrail.resize(h->nb, nullptr);
  for (int i = 0; i < h->nb && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&h->evPanel[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->evTrail[i], cudaEventDisableTiming);
  }
  for (int i = 0; i <= GPEDM_NPHASE && e == cudaSuccess; ++i) e = cudaEventCreate(&h->ev[i]);
  if (e != cudaSuccess) {
    free_handle(h);
    return set_err(GPEDM_ERR_CUDA, "handle setup failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return 0;
}

static int gpedm_create_impl(int device, int64_t N, int32_t d, int32_t np, const double* X, const double* Y,
                            const int32_t* pop, const double* rhomat, gpedm_handle* out) {
  if (!out) return set_err(GPEDM_ERR_ARG, "out handle is NULL");
  *out = nullptr;
  if (N < 1 || N > (1 << 20)) return set_err(GPEDM_ERR_ARG, "N=%lld out of range", (long long)N);
  if (np < 1) return set_err(GPEDM_ERR_ARG, "np=%d must be >= 1", np);
  if (!X || !Y || !pop) return set_err(GPEDM_ERR_ARG, "X, Y and pop must be non-NULL");
  for (int64_t i = 0; i < N; ++i)
    if (pop[i] < 0 || pop[i] >= np) return set_err(GPEDM_ERR_ARG, "pop[%lld]=%d outside 0..np-1", (long long)i, pop[i]);
  gpedm_fit_s* h = nullptr;
  int rc = alloc_handle(device, N, d, np, rhomat != nullptr, &h);
  if (rc) return rc;
  h->hY.assign(Y, Y + N);
  h->hpop.assign(pop, pop + N);
  h->single_pop = true;
  for (int64_t i = 1; i < N; ++i)
    if (pop[i] != pop[0]) h->single_pop = false;
  cudaError_t e = cudaMemcpy(h->dX, X, (size_t)N * d * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->dY, Y, (size_t)N * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->dpop, pop, (size_t)N * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && rhomat) e = cudaMemcpy(h->drhotab, rhomat, (size_t)np * np * 8, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    free_handle(h);
    return set_err(GPEDM_ERR_CUDA, "handle setup failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return 0;
}

static int gpedm_set_data_impl(gpedm_handle h, const double* X, const double* Y, const int32_t* pop) {
  if (!h || !X || !Y || !pop) return set_err(GPEDM_ERR_ARG, "NULL argument");
  for (int64_t i = 0; i < h->N; ++i)
    if (pop[i] < 0 || pop[i] >= h->np) return set_err(GPEDM_ERR_ARG, "pop[%lld]=%d outside 0..np-1", (long long)i, pop[i]);
  CU(cudaSetDevice(h->device));
  h->have_full = false;
  CU(cudaMemcpyAsync(h->dX, X, (size_t)h->N * h->d * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->dY, Y, (size_t)h->N * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->dpop, pop, (size_t)h->N * 4, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->hY.assign(Y, Y + h->N);
  h->hpop.assign(pop, pop + h->N);
  h->single_pop = true;
  for (int64_t i = 1; i < h->N; ++i)
    if (pop[i] != pop[0]) h->single_pop = false;
  return 0;
}

extern "C" int gpedm_destroy(gpedm_handle h) {
  free_handle(h);
  return 0;
}
extern "C" int64_t gpedm_n(gpedm_handle h) { return h ? h->N : -1; }
extern "C" int64_t gpedm_launch_count(gpedm_handle h) { return h ? h->launches : -1; }
extern "C" int gpedm_last_phase_ms(gpedm_handle h, float* ms) {
  if (!h || !ms) return set_err(GPEDM_ERR_ARG, "NULL argument");
  memcpy(ms, h->phase_ms, sizeof(h->phase_ms));
  return 0;
}


// ------------------------------------------------------------------------------- tile launches
// layout: 0 = (MN, MN), 1 = (MN, K), 2 = (K, K) operand layouts (see dgemm_tile.cuh).
static int launch_tiles(gpedm_fit_s* h, cudaStream_t st, const TileOp& op_in, int layout, bool split = false) {
  if (op_in.ntiles <= 0) return 0;
  TileOp op = op_in;
  op.tag = st == h->aux ? 1 : (st == h->fill ? 2 : 0);
  if (split && layout == 0)  // latency-critical launch with fewer tiles than SMs: 4 row strips per tile
    tile_once_kernel<false, false, GemmCfgSplit>
        <<<op.ntiles * (TB / GemmCfgSplit::TM), GemmCfgSplit::THREADS, GemmCfgSplit::SMEM_BYTES, st>>>(op);
  else if (!g_no_half_tiles && op.type != OP_PANEL && op.ntiles * 2 <= h->num_sms) {
    // launches with fewer tiles than half the SMs (low levels of the triangular inverse, small N):
    // 128 x 64 half tiles double the number of busy SMs.  (Measured slower than full tiles once the
    // grid fills the chip, so only used below that point.  The panel stays on row strips / full
    // tiles: it is updated in place and a column half would read columns another CTA overwrites.)
    const int grid = op.ntiles * 2;
    if (layout == 0)
      tile_once_kernel<false, false, GemmCfgHalf><<<grid, GemmCfgHalf::THREADS, GemmCfgHalf::SMEM_BYTES, st>>>(op);
    else if (layout == 1)
      tile_once_kernel<false, true, GemmCfgHalf><<<grid, GemmCfgHalf::THREADS, GemmCfgHalf::SMEM_BYTES, st>>>(op);
    else
      tile_once_kernel<true, true, GemmCfgHalf><<<grid, GemmCfgHalf::THREADS, GemmCfgHalf::SMEM_BYTES, st>>>(op);
  } else if (layout == 0)
    tile_once_kernel<false, false><<<op.ntiles, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(op);
  else if (layout == 1)
    tile_once_kernel<false, true><<<op.ntiles, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(op);
  else
    tile_once_kernel<true, true><<<op.ntiles, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(op);
  h->launches++;
  return 0;
}

static TileOp make_op(int type, int ntiles, int nb) {
  TileOp op;
  memset(&op, 0, sizeof(op));
  op.type = type;
  op.ntiles = ntiles;
  op.nb = nb;
  return op;
}


// Triangular inverse X = L^-1 by recursive doubling, restricted to what is computable once the
// first `ncols` block columns of L are final.  Level s (block size in tiles), pair p (tile rows/cols
// [2ps, 2ps+2s)): X21 = -X22 (L21 X11) in two launches,
//   W = L21 X11   needs X11 (level s/2, pair 2p) and block columns < 2ps+s of L,
//   X21 = -X22 W  needs X22 (level s/2, pair 2p+1), i.e. block columns < 2ps+2s.
// h->trtri_w_done / trtri_x_done count the full pairs already launched per level, so repeated calls
// with growing `ncols` launch every product exactly once and in dependency order (ascending s within a
// call, calls in stream order).  In particular the W product of the TOP level - a quarter of all
// the flops of the inverse - becomes available when the Cholesky is only half way.
// `final` also launches the ragged last pair of each level.
static int enqueue_trtri_levels(gpedm_fit_s* h, cudaStream_t st, int ncols, bool final) {
  const int nb = h->nb;
  const size_t ld = h->Np;
  int lv = 0;
  for (int s = 1; s < nb; s *= 2, ++lv) {
    const int pf = nb / (2 * s);  // full pairs of this level
    const int rem = nb - 2 * s * pf;
    const bool ragged = final && rem > s;
    // ---- W phase
    int wready = h->trtri_w_done[lv];
    while (wready < pf && 2 * wready * s + s <= ncols && (lv == 0 || h->trtri_x_done[lv - 1] >= 2 * wready + 1)) ++wready;
    {
      const int p0 = h->trtri_w_done[lv], nfull = wready - p0;
      const int ntl = nfull * s * s + (ragged ? (rem - s) * s : 0);
      if (ntl > 0) {
        TileOp w = make_op(OP_TRTRI_W, ntl, nb);
        w.s = s;
        w.p0 = p0;
        w.nfull = nfull;
        w.P0 = h->bufA;
        w.P1 = h->bufB;
        w.P2 = h->bufC;
        w.ld0 = w.ld1 = w.ld2 = ld;
        int rc = launch_tiles(h, st, w, 1);
        if (rc) return rc;
        h->trtri_w_done[lv] = wready;
      }
    }
    // ---- X phase
    {
      const int xready = std::min(h->trtri_w_done[lv], std::min(pf, ncols / (2 * s)));
      const int p0 = h->trtri_x_done[lv], nfull = std::max(0, xready - p0);
      const int ntl = nfull * s * s + (ragged ? (rem - s) * s : 0);
      if (ntl > 0) {
        TileOp x = make_op(OP_TRTRI_X, ntl, nb);
        x.s = s;
        x.p0 = p0;
        x.nfull = nfull;
        x.P0 = h->bufB;
        x.P1 = h->bufC;
        x.P2 = h->bufB;
        x.ld0 = x.ld1 = x.ld2 = ld;
        int rc = launch_tiles(h, st, x, 1);
        if (rc) return rc;
        h->trtri_x_done[lv] = p0 + nfull;
      }
    }
  }
  return 0;
}

// Phase-timing events must stay usable from outside when the launch sequence is captured into a
// CUDA graph: record them as "external" event nodes during capture.
static thread_local bool g_capturing = false;  // set only inside enqueue_eval while capturing
static cudaError_t record_phase_event(cudaEvent_t ev, cudaStream_t st) {
  return g_capturing ? cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal) : cudaEventRecord(ev, st);
}

// ------------------------------------------------------------------------------- pipeline
// Enqueue the device work of one evaluation on h->stream.  h->hparams must be filled.
// stages: through z/a (need_inverse_products=false stops after z for the sequential helper).
static int enqueue_eval_raw(gpedm_fit_s* h, bool full, bool through_z_only) {
  cudaStream_t st = h->stream;
  const int nb = h->nb, Np = h->Np, N = (int)h->N, d = h->d;
  const size_t ld = Np;
  CU(cudaMemcpyAsync(h->dparams, h->hparams, sizeof(EvalParams), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(h->dinfo, 0, 4, st));
  CU(record_phase_event(h->ev[0], st));
  // --- assemble Sigma (lower tiles) into bufA (or take the caller's matrix: gpedm_covinv)
  if (h->dsigma_in) {
    load_sigma_kernel<<<h->ntiles, 256, 0, st>>>(h->bufA, ld, N, h->dsigma_in);
    h->launches++;
  } else {
    int smem = 2 * d * TB * 8 + 2 * TB * 4;
    assemble_sigma_kernel<<<h->ntiles, 256, smem, st>>>(h->bufA, ld, N, h->dX, h->dpop, h->drhotab, h->dparams);
    h->launches++;
  }
  CU(record_phase_event(h->ev[1], st));
  // --- Cholesky, right-looking over 128-wide block columns, with look-ahead and grouped updates.
  // Steps are processed in groups [g0, g1) of G block columns (G = GPEDM_POTRF_GROUP while the
  // trailing matrix is large, 1 near the end); [g1, g2) is the next group.
  //   aux (high priority): for q in [g0, g1):
  //        column q <- pending updates of the group's earlier steps (rank 128 (q-g0))
  //        diag(q) -> panel(q)
  //      then columns [g1, g2) <- all G steps of the group (rank 128 G), so the next group's whole
  //      chain depends only on this stream;
  //   main: columns >= g2 <- all G steps in ONE rank-128G update: each trailing tile is read and
  //        written once per group instead of once per step and the k loop is G times longer.
  // The chain of group g+1 overlaps the bulk update of group g.
  {
    cudaStream_t ax = g_no_lookahead ? st : h->aux;
    CU(cudaEventRecord(h->evFork, st));
    CU(cudaStreamWaitEvent(ax, h->evFork, 0));
    auto trail = [&](cudaStream_t s_, int k0, int kc, int jlo, int ntl) -> int {
      TileOp t = make_op(OP_TRAIL, ntl, nb);
      t.k = k0;
      t.kcount = kc;
      t.jlo = jlo;
      t.P2 = h->bufA;
      t.ld2 = ld;
      return launch_tiles(h, s_, t, 0, s_ == ax && !g_no_split && ntl * 4 <= 2 * h->num_sms);
    };
    // group size: GPEDM_POTRF_GROUP if set, else 4 from 64 block columns up (bulk-dominated; measured
    // 20.46 -> 20.15 ms at N=8192 together with group_min 16 / fill_tail 40), 2 below (4.07 vs 4.17 ms at N=4096)
    const int gsz = g_group > 0 ? g_group : (nb >= 64 ? 4 : 2);
    auto group_size = [&](int start) { return (nb - start - 1 >= g_group_min) ? gsz : 1; };
    // Early triangular inverse: level s, pair p (tile rows/cols [2ps, 2ps+2s)) only needs block
    // columns < 2s(p+1) of L and the diagonal-block inverses, so it can run as soon as the Cholesky
    // has passed that column - on a low-priority stream, filling the SMs the chain-bound tail of the
    // factorisation leaves idle.  Checkpoints every 8 block columns.
    cudaStream_t fl = (g_no_lookahead || g_no_fill) ? nullptr : h->fill;
    if (fl) CU(cudaStreamWaitEvent(fl, h->evFork, 0));
    for (int lv = 0; lv < 16; ++lv) h->trtri_w_done[lv] = h->trtri_x_done[lv] = 0;
    int prev_bulk = -1;  // index of the event recorded after the previous group's bulk update
    int g0 = 0, G = group_size(0);
    while (g0 < nb) {
      const int g1 = std::min(nb, g0 + G);
      for (int q = g0; q < g1; ++q) {
        if (q > g0) {
          int rc = trail(ax, g0, q - g0, q, nb - q);  // column q only
          if (rc) return rc;
        }
        if (g_diag_ref)
          diag_potf2_inv_kernel<<<1, 256, DIAG_SMEM_BYTES, ax>>>(h->bufA + (size_t)q * TB * (ld + 1), ld,
                                                                  h->bufB + (size_t)q * TB * (ld + 1), ld, q, N,
                                                                  h->dlogdet, h->dinfo);
        else
          diag_block_kernel<<<1, DG_THREADS, DG_SMEM_BYTES, ax>>>(h->bufA + (size_t)q * TB * (ld + 1), ld,
                                                                  h->bufB + (size_t)q * TB * (ld + 1), ld, q, N,
                                                                  h->dlogdet, h->dinfo);
        h->launches++;
        if (nb - q - 1 > 0) {
          TileOp pan = make_op(OP_PANEL, nb - q - 1, nb);
          pan.k = q;
          pan.P1 = h->bufB;
          pan.P2 = h->bufA;
          pan.ld1 = pan.ld2 = ld;
          int rc = launch_tiles(h, ax, pan, 0, !g_no_split && pan.ntiles * 4 <= 2 * h->num_sms);
          if (rc) return rc;
        }
      }
      if (fl && g1 % 8 == 0 && g1 < nb && nb - g1 <= g_fill_tail) {  // only once the factorisation is chain-bound
        CU(cudaEventRecord(h->evCol[g1 / 8], ax));
        CU(cudaStreamWaitEvent(fl, h->evCol[g1 / 8], 0));
        int rc = enqueue_trtri_levels(h, fl, g1, false);
        if (rc) return rc;
      }
      if (g1 >= nb) break;
      const int Gn = group_size(g1);
      const int g2 = std::min(nb, g1 + Gn);
      CU(cudaEventRecord(h->evPanel[g1 - 1], ax));
      if (prev_bulk >= 0) CU(cudaStreamWaitEvent(ax, h->evTrail[prev_bulk], 0));
      int next_cols_tiles = 0;
      for (int c = g1; c < g2; ++c) next_cols_tiles += nb - c;
      int rc = trail(ax, g0, g1 - g0, g1, next_cols_tiles);  // columns [g1, g2)
      if (rc) return rc;
      CU(cudaStreamWaitEvent(st, h->evPanel[g1 - 1], 0));
      if (g2 < nb) {
        const int m = nb - g2;
        rc = trail(st, g0, g1 - g0, g2, m * (m + 1) / 2);  // columns >= g2
        if (rc) return rc;
      }
      CU(cudaEventRecord(h->evTrail[g1 - 1], st));
      prev_bulk = g1 - 1;
      g0 = g1;
      G = Gn;
    }
    CU(cudaEventRecord(h->evJoin, ax));
    CU(cudaStreamWaitEvent(st, h->evJoin, 0));
    if (fl) {
      CU(cudaEventRecord(h->evFill, fl));
      CU(cudaStreamWaitEvent(st, h->evFill, 0));
    }
  }
  CU(record_phase_event(h->ev[2], st));
  // --- triangular inverse X = L^-1 by recursive doubling (diagonal blocks already in bufB): whatever the
  // fill stream has not done yet, level by level
  {
    int rc = enqueue_trtri_levels(h, st, nb, true);
    if (rc) return rc;
  }
  CU(record_phase_event(h->ev[3], st));
  // --- z = X Y
  trmv_n_part_kernel<<<h->ntiles, 256, 0, st>>>(h->bufB, ld, h->dY, h->dpart, ld);
  reduce_parts_kernel<<<ceil_div(Np, 256), 256, 0, st>>>(h->dpart, ld, nb, 0, h->dz, Np);
  h->launches += 2;
  if (through_z_only) {
    CU(record_phase_event(h->ev[4], st));
    CU(cudaMemcpyAsync(h->hinfo, h->dinfo, 4, cudaMemcpyDeviceToHost, st));
    return 0;
  }
  // --- a = X^T z
  trmv_t_part_kernel<<<h->ntiles, 256, 0, st>>>(h->bufB, ld, h->dz, h->dpart, ld);
  reduce_parts_kernel<<<ceil_div(Np, 256), 256, 0, st>>>(h->dpart, ld, nb, 1, h->da, Np);
  h->launches += 2;
  CU(record_phase_event(h->ev[4], st));
  // --- S = Sigma^-1 = X^T X (lower tiles) into bufC
  {
    TileOp lu = make_op(OP_LAUUM, h->ntiles, nb);
    lu.P0 = h->bufB;
    lu.P2 = h->bufC;
    lu.ld0 = lu.ld2 = ld;
    int rc = launch_tiles(h, st, lu, 2);
    if (rc) return rc;
  }
  CU(record_phase_event(h->ev[5], st));
  // --- gradient trace terms + scalar reductions
  {
    int DC = d <= 8 ? 8 : (d <= 16 ? 16 : 32);
    int smem = (4 * d * TB + 2 * TB + 8 * (DC + 3)) * 8 + 2 * TB * 4;
    if (DC <= 16) {
      // The kernel is compiled for 3 CTAs/SM (latency-bound: more resident warps help), but a CTA then
      // runs ~1.26x longer than at 2 CTAs/SM (measured), so when the tile count quantises into the same
      // number of waves either way, cap residency at 2 by padding the dynamic shared memory request.
      const int w3 = ceil_div(h->ntiles, 3 * h->num_sms), w2 = ceil_div(h->ntiles, 2 * h->num_sms);
      if (1.26 * w3 > w2) smem = std::max(smem, 80 * 1024);
    }
    if (DC == 8)
      grad_trace_kernel<8><<<h->ntiles, 256, smem, st>>>(h->bufC, ld, N, h->dX, h->dpop, h->drhotab, h->da,
                                                         h->dparams, h->dtracepart);
    else if (DC == 16)
      grad_trace_kernel<16><<<h->ntiles, 256, smem, st>>>(h->bufC, ld, N, h->dX, h->dpop, h->drhotab, h->da,
                                                          h->dparams, h->dtracepart);
    else
      grad_trace_kernel<32><<<h->ntiles, 256, smem, st>>>(h->bufC, ld, N, h->dX, h->dpop, h->drhotab, h->da,
                                                          h->dparams, h->dtracepart);
    h->launches++;
    if (full) {
      extract_diag_kernel<<<ceil_div(N, 256), 256, 0, st>>>(h->bufC, ld, N, h->ddiagS);
      h->launches++;
    }
    final_reduce_kernel<<<1, 256, 0, st>>>(h->dlogdet, nb, h->dz, h->da, h->ddiagS, N, h->dtracepart, h->ntiles,
                                           d + 3, full ? 1 : 0, h->dscal);
    h->launches++;
  }
  CU(record_phase_event(h->ev[6], st));
  CU(cudaMemcpyAsync(h->hscal, h->dscal, NSCAL * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->hinfo, h->dinfo, 4, cudaMemcpyDeviceToHost, st));
  return 0;
}


// One evaluation = a fixed launch sequence on two streams (parameters live in device memory), so
// it is captured once per handle and variant into a CUDA graph and replayed: one graph launch
// instead of hundreds of kernel launches / event waits per evaluation.
static int enqueue_eval(gpedm_fit_s* h, bool full, bool through_z_only = false) {
  // Above ~48 block columns the bulk kernels dominate and the stream-priority look-ahead (which the
  // graph replay does not honour as well as live streams do, measured) matters more than launch cost.
  if (g_no_graph || h->nb > g_graph_max_nb) return enqueue_eval_raw(h, full, through_z_only);
  const int v = through_z_only ? 2 : (full ? 1 : 0);
  if (!h->graph_exec[v]) {
    const int64_t l0 = h->launches;
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    g_capturing = true;
    int rc = enqueue_eval_raw(h, full, through_z_only);
    g_capturing = false;
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->graph_exec[v], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(e));
    h->graph_launches[v] = h->launches - l0;
    h->launches = l0;
  }
  CU(cudaGraphLaunch(h->graph_exec[v], h->stream));
  h->launches += h->graph_launches[v];
  return 0;
}

static int collect_phase_times(gpedm_fit_s* h) {
  // ev: 0 start, 1 after assemble, 2 after potrf, 3 after trtri, 4 after vec, 5 after lauum, 6 after trace
  float t;
  CU(cudaEventElapsedTime(&t, h->ev[0], h->ev[1]));
  h->phase_ms[GPEDM_PHASE_ASSEMBLE] = t;
  CU(cudaEventElapsedTime(&t, h->ev[1], h->ev[2]));
  h->phase_ms[GPEDM_PHASE_POTRF] = t;
  CU(cudaEventElapsedTime(&t, h->ev[2], h->ev[3]));
  h->phase_ms[GPEDM_PHASE_TRTRI] = t;
  CU(cudaEventElapsedTime(&t, h->ev[3], h->ev[4]));
  h->phase_ms[GPEDM_PHASE_VEC] = t;
  CU(cudaEventElapsedTime(&t, h->ev[4], h->ev[5]));
  h->phase_ms[GPEDM_PHASE_LAUUM] = t;
  CU(cudaEventElapsedTime(&t, h->ev[5], h->ev[6]));
  h->phase_ms[GPEDM_PHASE_TRACE] = t;
  CU(cudaEventElapsedTime(&t, h->ev[0], h->ev[6]));
  h->phase_ms[GPEDM_PHASE_TOTAL] = t;
  return 0;
}

// Host-side scalar head of getlikegrad (reference R/GPfunctions.R:590-630).
struct HostPars {
  double phi[MAXD], ve, sigma2, rho;
  double parst[GPEDM_MAX_P], dpars[GPEDM_MAX_P], dlp[GPEDM_MAX_P];
  double lp;
  bool rho_is_fixed;
  bool fixed[GPEDM_MAX_P];
};

static void host_transform(const gpedm_fit_s* h, const double* parst_in, double modeprior, const double* fixedpars,
                           double rhofixed, HostPars& o) {
  const int d = h->d;
  for (int i = 0; i < d + 3; ++i) o.parst[i] = parst_in[i];
  for (int i = 0; i < d; ++i) o.phi[i] = exp(o.parst[i]);
  o.ve = (VEMAX - VEMIN) / (1 + exp(-o.parst[d])) + VEMIN;
  o.sigma2 = (SIGMA2MAX - SIGMA2MIN) / (1 + exp(-o.parst[d + 1])) + SIGMA2MIN;
  o.rho = (RHOMAX - RHOMIN) / (1 + exp(-o.parst[d + 2])) + RHOMIN;
  // :600-602  one population or rhomatrix  => rho fixed to 0.5
  if (h->single_pop || h->has_rhomat) rhofixed = 0.5;
  o.rho_is_fixed = !isnan(rhofixed);
  if (o.rho_is_fixed) {
    o.rho = rhofixed;
    o.parst[d + 2] = logit3(o.rho, RHOMIN, RHOMAX);
  }
  for (int i = 0; i < d + 3; ++i) o.fixed[i] = false;
  if (fixedpars) {
    bool anyphi = false;
    for (int i = 0; i < d; ++i)
      if (!isnan(fixedpars[i])) {
        o.phi[i] = fixedpars[i];
        o.fixed[i] = true;
        anyphi = true;
      }
    if (anyphi)
      for (int i = 0; i < d; ++i) o.parst[i] = log(o.phi[i]);  // :611
    if (!isnan(fixedpars[d])) {
      o.ve = fixedpars[d];
      o.parst[d] = logit3(o.ve, VEMIN, VEMAX);
      o.fixed[d] = true;
    }
    if (!isnan(fixedpars[d + 1])) {
      o.sigma2 = fixedpars[d + 1];
      o.parst[d + 1] = logit3(o.sigma2, SIGMA2MIN, SIGMA2MAX);
      o.fixed[d + 1] = true;
    }
  }
  for (int i = 0; i < d; ++i) o.dpars[i] = o.phi[i];
  o.dpars[d] = (o.ve - VEMIN) * (1 - (o.ve - VEMIN) / (VEMAX - VEMIN));
  o.dpars[d + 1] = (o.sigma2 - SIGMA2MIN) * (1 - (o.sigma2 - SIGMA2MIN) / (SIGMA2MAX - SIGMA2MIN));
  o.dpars[d + 2] = (o.rho - RHOMIN) * (1 - (o.rho - RHOMIN) / (RHOMAX - RHOMIN));
  // getpriors :752-777
  const double lam = pow(pow(2.0, modeprior - 1), 2) * M_PI / 2;
  double sphi2 = 0;
  for (int i = 0; i < d; ++i) sphi2 += o.phi[i] * o.phi[i];
  double lp_phi = -0.5 * sphi2 / lam;
  for (int i = 0; i < d; ++i) o.dlp[i] = -o.phi[i] / lam;
  double lp_s2 = log(o.sigma2 / SIGMA2MAX) + log(1 - o.sigma2 / SIGMA2MAX);
  o.dlp[d + 1] = 1.0 / o.sigma2 - 1.0 / (SIGMA2MAX - o.sigma2);
  double lp_ve = log(o.ve / VEMAX) + log(1 - o.ve / VEMAX);
  o.dlp[d] = 1.0 / o.ve - 1.0 / (VEMAX - o.ve);
  double lp_rho = log(o.rho) + log(1 - o.rho);
  o.dlp[d + 2] = 1.0 / o.rho - 1.0 / (1 - o.rho);
  o.lp = lp_phi + lp_ve + lp_s2 + lp_rho;
}

static void fill_params(gpedm_fit_s* h, const HostPars& hp) {
  EvalParams* P = h->hparams;
  memset(P, 0, sizeof(*P));
  for (int i = 0; i < h->d; ++i) {
    P->phi[i] = hp.phi[i];
    P->sqphi[i] = sqrt(hp.phi[i]);
  }
  P->ve = hp.ve;
  P->sigma2 = hp.sigma2;
  P->rho = hp.rho;
  P->d = h->d;
  P->np = h->np;
  P->use_rhotab = h->has_rhomat ? 1 : 0;
}

static void finish_result(const gpedm_fit_s* h, const HostPars& hp, bool full, gpedm_result* out) {
  const int d = h->d;
  const double* sc = h->hscal;
  double logdet = sc[0], SS = sc[1];
  double like = -0.5 * SS - logdet;  // :648
  double dl[GPEDM_MAX_P];
  for (int i = 0; i < d; ++i) dl[i] = hp.fixed[i] ? 0.0 : 0.5 * sc[8 + i];  // :659-665 (note the extra 0.5)
  dl[d] = hp.fixed[d] ? 0.0 : sc[8 + d];                                    // :688-692
  dl[d + 1] = hp.fixed[d + 1] ? 0.0 : sc[8 + d + 1];                        // :694-698
  dl[d + 2] = hp.rho_is_fixed ? 0.0 : sc[8 + d + 2];                        // :701-709
  memset(out, 0, sizeof(*out));
  out->p = d + 3;
  double g2 = 0;
  for (int i = 0; i < d + 3; ++i) {
    double J = dl[i] + hp.dlp[i];
    out->grad[i] = -J * hp.dpars[i];  // :719-720
    g2 += out->grad[i] * out->grad[i];
    out->parst[i] = hp.parst[i];
  }
  for (int i = 0; i < d; ++i) out->pars[i] = hp.phi[i];
  out->pars[d] = hp.ve;
  out->pars[d + 1] = hp.sigma2;
  out->pars[d + 2] = hp.rho;
  out->gradnorm = sqrt(g2);
  out->lp = hp.lp;
  out->like = like;
  out->nllpost = -(like + hp.lp);  // :721
  out->logdet = logdet;
  out->SS = SS;
  if (full) {
    out->lnL_LOO = 0.5 * sc[2] - 0.5 * sc[3];            // :742
    out->df = (double)h->N - hp.ve * sc[4];              // :743 via tr(Cd S) = N - ve tr(S)
  } else {
    out->lnL_LOO = NAN;
    out->df = NAN;
  }
}

static int eval_internal(gpedm_fit_s* h, const double* parst, double modeprior, const double* fixedpars,
                         double rhofixed, bool full, gpedm_result* out) {
  CU(cudaSetDevice(h->device));
  HostPars hp;
  host_transform(h, parst, modeprior, fixedpars, rhofixed, hp);
  fill_params(h, hp);
  h->have_full = false;
  int rc = enqueue_eval(h, full);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaGetLastError());
  rc = collect_phase_times(h);
  if (rc) return rc;
  finish_result(h, hp, full, out);
  out->nevals = 1;
  int info = *h->hinfo;
  out->status = info;
  if (info > 0)
    return set_err(info, "covariance matrix not positive definite: non-positive pivot at row %d", info);
  if (full) {
    h->have_full = true;
    h->cur_ve = hp.ve;
    h->cur_sigma2 = hp.sigma2;
    h->cur_rho = hp.rho;
  }
  return 0;
}

extern "C" int gpedm_likegrad(gpedm_handle h, const double* parst, double modeprior, const double* fixedpars,
                              double rhofixed, int full, gpedm_result* out) {
  if (!h || !parst || !out) return set_err(GPEDM_ERR_ARG, "NULL argument");
  return eval_internal(h, parst, modeprior, fixedpars, rhofixed, full != 0, out);
}

static inline double sgn(double x) { return (x > 0) - (x < 0); }

extern "C" int gpedm_fit_rprop(gpedm_handle h, const double* initparst, double modeprior, const double* fixedpars,
                               double rhofixed, const gpedm_rprop_options* opts, gpedm_result* out) {
  if (!h || !initparst || !out) return set_err(GPEDM_ERR_ARG, "NULL argument");
  gpedm_rprop_options o;
  if (opts)
    o = *opts;
  else
    gpedm_default_rprop_options(&o);
  const int p = h->p;
  const double eta_minus = o.eta_minus - 1, eta_plus = o.eta_plus - 1;  // :534-535
  double pa[GPEDM_MAX_P], del[GPEDM_MAX_P], grad[GPEDM_MAX_P], panew[GPEDM_MAX_P];
  gpedm_result r;
  int rc = eval_internal(h, initparst, modeprior, fixedpars, rhofixed, false, &r);  // :545
  if (rc) {
    *out = r;
    return rc;
  }
  int nevals = 1;
  double nll = r.nllpost;
  for (int i = 0; i < p; ++i) {
    grad[i] = r.grad[i];
    pa[i] = initparst[i];
    del[i] = o.delta0;
  }
  double s = r.gradnorm;
  int iter = 0;
  double df = 10;
  while (s > o.tol_grad && iter < o.maxiter && df > o.tol_df) {  // :556
    for (int i = 0; i < p; ++i) panew[i] = pa[i] - sgn(grad[i]) * del[i];  // :559
    rc = eval_internal(h, panew, modeprior, fixedpars, rhofixed, false, &r);
    ++nevals;
    if (rc) {
      *out = r;
      out->iter = iter;
      out->nevals = nevals;
      return rc;
    }
    s = r.gradnorm;
    df = fabs(r.nllpost / nll - 1);  // :565
    for (int i = 0; i < p; ++i) {
      double gc = grad[i] * r.grad[i];
      double tdel = del[i] * (1 + eta_plus * (gc > 0) + eta_minus * (gc < 0));  // :569
      del[i] = tdel < o.deltamin ? o.deltamin : (tdel > o.deltamax ? o.deltamax : tdel);
      pa[i] = panew[i];
      grad[i] = r.grad[i];
    }
    nll = r.nllpost;
    ++iter;
  }
  rc = eval_internal(h, pa, modeprior, fixedpars, rhofixed, true, &r);  // :580
  ++nevals;
  *out = r;
  out->iter = iter;
  out->nevals = nevals;
  return rc;
}

// ------------------------------------------------------------------------------- exports
static int gpedm_get_vector_impl(gpedm_handle h, int which, double* outv) {
  if (!h || !outv) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (!h->have_full) return set_err(GPEDM_ERR_STATE, "no full evaluation resident (call gpedm_likegrad(full=1) or gpedm_fit_rprop first)");
  CU(cudaSetDevice(h->device));
  const int64_t N = h->N;
  std::vector<double> a(N), ds(N);
  CU(cudaMemcpy(a.data(), h->da, N * 8, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ds.data(), h->ddiagS, N * 8, cudaMemcpyDeviceToHost));
  const double ve = h->cur_ve;
  for (int64_t i = 0; i < N; ++i) {
    switch (which) {
      case GPEDM_VEC_MEAN: outv[i] = h->hY[i] - ve * a[i]; break;            // Cd a = Y - ve a
      case GPEDM_VEC_VAR: outv[i] = ve - ve * ve * ds[i]; break;             // diag(Cd - Cd S Cd)
      case GPEDM_VEC_ALPHA: outv[i] = a[i]; break;
      case GPEDM_VEC_DIAG_IKVS: outv[i] = ds[i]; break;
      case GPEDM_VEC_LOO_MEAN: outv[i] = h->hY[i] - a[i] / ds[i]; break;
      case GPEDM_VEC_LOO_VAR: outv[i] = 1.0 / ds[i] - ve; break;
      default: return set_err(GPEDM_ERR_ARG, "unknown vector id %d", which);
    }
  }
  return 0;
}

static int launch_cov_rect(gpedm_fit_s* h, double* out, size_t ldo, int rows_alloc, int cols_alloc, int transpose,
                           int M, const double* dX2, const int* dpop2, int add_ve) {
  int smem = 2 * h->d * TB * 8 + 2 * TB * 4;
  dim3 grid(ceil_div(rows_alloc, TB), ceil_div(cols_alloc, TB));
  cov_rect_kernel<<<grid, 256, smem, h->stream>>>(out, ldo, rows_alloc, cols_alloc, transpose, (int)h->N, h->dX,
                                                  h->dpop, M, dX2, dpop2, h->drhotab, h->dparams, add_ve);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

static int gpedm_get_matrix_impl(gpedm_handle h, int which, double* outm) {
  if (!h || !outm) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (!h->have_full) return set_err(GPEDM_ERR_STATE, "no full evaluation resident");
  CU(cudaSetDevice(h->device));
  const int64_t N = h->N;
  const size_t ld = h->Np;
  if (which == GPEDM_MAT_CD || which == GPEDM_MAT_SIGMA || which == GPEDM_MAT_POPSAME) {
    double* tmp = nullptr;
    CU(cudaMalloc((void**)&tmp, (size_t)N * N * 8));
    int rc = 0;
    if (which == GPEDM_MAT_POPSAME) {
      size_t tot = (size_t)N * N;
      popsame_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(tmp, (int)N, h->dpop);
      h->launches++;
    } else {
      rc = launch_cov_rect(h, tmp, (size_t)N, (int)N, (int)N, 0, (int)N, h->dX, h->dpop,
                           which == GPEDM_MAT_SIGMA ? 1 : 0);
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(outm, tmp, (size_t)N * N * 8, cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    if (rc) return rc;
    if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "export failed: %s", cudaGetErrorString(e));
    return 0;
  }
  const double* src = which == GPEDM_MAT_IKVS ? h->bufC : (which == GPEDM_MAT_L ? h->bufA : (which == GPEDM_MAT_LINV ? h->bufB : nullptr));
  if (!src) return set_err(GPEDM_ERR_ARG, "unknown matrix id %d", which);
  CU(cudaMemcpy2D(outm, (size_t)N * 8, src, ld * 8, (size_t)N * 8, (size_t)N, cudaMemcpyDeviceToHost));
  for (int64_t j = 0; j < N; ++j)
    for (int64_t i = 0; i < j; ++i)
      outm[i + j * N] = which == GPEDM_MAT_IKVS ? outm[j + i * N] : 0.0;
  return 0;
}

static int gpedm_getcov_impl(int device, int32_t d, const double* phi, double sigma2, double rho, int64_t N,
                            const double* X, const int32_t* pop, int64_t M, const double* X2, const int32_t* pop2,
                            int32_t np, const double* rhomat, double* Cd_out) {
  if (!phi || !X || !pop || !X2 || !pop2 || !Cd_out) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (d < 1 || d > GPEDM_MAX_D || N < 1 || M < 1) return set_err(GPEDM_ERR_ARG, "bad sizes");
  int ndev = gpedm_device_count();
  if (ndev == 0) return set_err(GPEDM_ERR_CUDA, "no CUDA device available (libgpedm_b200 has no CPU fallback)");
  CU(cudaSetDevice(device));
  int rc = ensure_kernel_attrs(device);
  if (rc) return rc;
  double *dX = nullptr, *dX2 = nullptr, *dout = nullptr, *drt = nullptr;
  int *dp = nullptr, *dp2 = nullptr;
  EvalParams* dP = nullptr;
  EvalParams P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < d; ++i) {
    P.phi[i] = phi[i];
    P.sqphi[i] = sqrt(phi[i]);
  }
  P.sigma2 = sigma2;
  P.rho = rho;
  P.d = d;
  P.np = np;
  P.use_rhotab = rhomat ? 1 : 0;
  cudaError_t e = cudaMalloc((void**)&dX, (size_t)N * d * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dX2, (size_t)M * d * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dout, (size_t)M * N * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dp, (size_t)N * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dp2, (size_t)M * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dP, sizeof(EvalParams));
  if (e == cudaSuccess && rhomat) e = cudaMalloc((void**)&drt, (size_t)np * np * 8);
  if (e == cudaSuccess) e = cudaMemcpy(dX, X, (size_t)N * d * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dX2, X2, (size_t)M * d * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dp, pop, (size_t)N * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dp2, pop2, (size_t)M * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dP, &P, sizeof(P), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && rhomat) e = cudaMemcpy(drt, rhomat, (size_t)np * np * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    int smem = 2 * d * TB * 8 + 2 * TB * 4;
    dim3 grid(ceil_div((int)M, TB), ceil_div((int)N, TB));
    cov_rect_kernel<<<grid, 256, smem>>>(dout, (size_t)M, (int)M, (int)N, 0, (int)N, dX, dp, (int)M, dX2, dp2, drt,
                                         dP, 0);
    e = cudaDeviceSynchronize();
  }
  if (e == cudaSuccess) e = cudaMemcpy(Cd_out, dout, (size_t)M * N * 8, cudaMemcpyDeviceToHost);
  cudaFree(dX);
  cudaFree(dX2);
  cudaFree(dout);
  cudaFree(dp);
  cudaFree(dp2);
  cudaFree(dP);
  cudaFree(drt);
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "gpedm_getcov failed: %s", cudaGetErrorString(e));
  return 0;
}


// getcovinv (reference R/GPfunctions.R:818-830): Cholesky factor and full inverse of a caller-supplied
// symmetric positive definite matrix, through the same factorisation pipeline as the likelihood.
static int gpedm_covinv_impl(int device, int64_t n, const double* Sigma, double* L_out, double* iKVs_out) {
  if (!Sigma || (!L_out && !iKVs_out) || n < 1) return set_err(GPEDM_ERR_ARG, "bad arguments");
  std::vector<double> X((size_t)n, 0.0), Y((size_t)n, 0.0);
  std::vector<int32_t> pop((size_t)n, 0);
  gpedm_handle h = nullptr;
  int rc = gpedm_create(device, n, 1, 1, X.data(), Y.data(), pop.data(), nullptr, &h);
  if (rc) return rc;
  double* dS = nullptr;
  cudaError_t e = cudaMalloc((void**)&dS, (size_t)n * n * 8);
  if (e == cudaSuccess) e = cudaMemcpy(dS, Sigma, (size_t)n * n * 8, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(dS);
    gpedm_destroy(h);
    return set_err(GPEDM_ERR_CUDA, "gpedm_covinv: %s", cudaGetErrorString(e));
  }
  h->dsigma_in = dS;
  memset(h->hparams, 0, sizeof(EvalParams));
  h->hparams->d = 1;
  h->hparams->np = 1;
  h->hparams->sigma2 = 1.0;
  h->hparams->rho = 0.5;
  rc = enqueue_eval(h, true);
  if (!rc) {
    e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = set_err(GPEDM_ERR_CUDA, "gpedm_covinv: %s", cudaGetErrorString(e));
  }
  if (!rc && *h->hinfo > 0)
    rc = set_err(*h->hinfo, "matrix not positive definite: non-positive pivot at row %d", *h->hinfo);
  if (!rc) {
    h->have_full = true;
    if (L_out) rc = gpedm_get_matrix(h, GPEDM_MAT_L, L_out);
    if (!rc && iKVs_out) rc = gpedm_get_matrix(h, GPEDM_MAT_IKVS, iKVs_out);
  }
  cudaFree(dS);
  gpedm_destroy(h);
  return rc;
}

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

// ------------------------------------------------------------------------------- prediction
static int gpedm_predict_new_impl(gpedm_handle h, int64_t M, const double* Xnew, const int32_t* popnew, double* mean,
                                 double* var, double* gpgrad) {
  if (!h || !Xnew || !popnew || !mean || !var) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (M < 1) return set_err(GPEDM_ERR_ARG, "M must be >= 1");
  if (!h->have_full) return set_err(GPEDM_ERR_STATE, "no full evaluation resident");
  for (int64_t i = 0; i < M; ++i)
    if (popnew[i] < 0 || popnew[i] >= h->np) return set_err(GPEDM_ERR_ARG, "popnew[%lld] outside 0..np-1", (long long)i);
  CU(cudaSetDevice(h->device));
  const int d = h->d;
  const int64_t CH = 4096;  // new points per chunk
  const int64_t chunk_max = std::min<int64_t>(M, CH);
  const int Mp_max = ceil_div((int)chunk_max, TB) * TB;
  const size_t ldv = h->Np;
  double *dV = nullptr, *dU = nullptr, *dXn = nullptr, *dmean = nullptr, *dvar = nullptr, *dgrad = nullptr;
  int* dpn = nullptr;
  cudaError_t e = cudaMalloc((void**)&dV, ldv * Mp_max * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dU, ldv * Mp_max * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dXn, (size_t)chunk_max * d * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dpn, (size_t)chunk_max * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dmean, (size_t)chunk_max * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dvar, (size_t)chunk_max * 8);
  if (e == cudaSuccess && gpgrad) e = cudaMalloc((void**)&dgrad, (size_t)chunk_max * d * 8);
  int rc = 0;
  std::vector<double> xn, gtmp;
  for (int64_t m0 = 0; m0 < M && e == cudaSuccess && rc == 0; m0 += CH) {
    const int64_t mc = std::min<int64_t>(CH, M - m0);
    const int Mp = ceil_div((int)mc, TB) * TB;
    xn.resize((size_t)mc * d);
    for (int k = 0; k < d; ++k)
      for (int64_t i = 0; i < mc; ++i) xn[i + (size_t)k * mc] = Xnew[m0 + i + (size_t)k * M];
    e = cudaMemcpyAsync(dXn, xn.data(), (size_t)mc * d * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dpn, popnew + m0, (size_t)mc * 4, cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) break;
    rc = launch_cov_rect(h, dV, ldv, h->Np, Mp, 1, (int)mc, dXn, dpn, 0);
    if (rc) break;
    {
      TileOp ta = make_op(OP_TRI_APPLY, h->nb * (Mp / TB), h->nb);
      ta.mt = Mp / TB;
      ta.P0 = h->bufB;
      ta.P1 = dV;
      ta.P2 = dU;
      ta.ld0 = ldv;
      ta.ld1 = ta.ld2 = ldv;
      rc = launch_tiles(h, h->stream, ta, 1);
      if (rc) break;
    }
    predict_reduce_kernel<<<(unsigned)mc, 256, 0, h->stream>>>(dV, dU, ldv, (int)h->N, (int)mc, h->da, h->dX, dXn,
                                                                h->dparams, dmean, dvar, dgrad);
    h->launches += 1;
    e = cudaMemcpyAsync(mean + m0, dmean, (size_t)mc * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(var + m0, dvar, (size_t)mc * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && gpgrad) {
      gtmp.resize((size_t)mc * d);
      e = cudaMemcpyAsync(gtmp.data(), dgrad, (size_t)mc * d * 8, cudaMemcpyDeviceToHost, h->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess && gpgrad)
      for (int k = 0; k < d; ++k)
        for (int64_t i = 0; i < mc; ++i) gpgrad[m0 + i + (size_t)k * M] = gtmp[i + (size_t)k * mc];
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(dV);
  cudaFree(dU);
  cudaFree(dXn);
  cudaFree(dpn);
  cudaFree(dmean);
  cudaFree(dvar);
  cudaFree(dgrad);
  if (rc) return rc;
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "gpedm_predict_new failed: %s", cudaGetErrorString(e));
  return 0;
}

// Shared implementation of loo / lto given the groups (excluded sets) and focal rows.
//   members/goff : concatenated excluded-row indices per group;  focal[f], fgroup[f].
// The block solves run on the device (leaveout_blocks_kernel); the host only maps focal rows to
// their position inside their group.
static int leaveout_core(gpedm_fit_s* h, const std::vector<int>& members, const std::vector<int>& goff,
                         const std::vector<int>& focal, const std::vector<int>& fgroup, double* mean, double* var,
                         double* gpgrad) {
  CU(cudaSetDevice(h->device));
  const int64_t N = h->N;
  const int ng = (int)goff.size() - 1, nf = (int)focal.size();
  const size_t nm = members.size();
  std::vector<long long> boff(ng + 1, 0);
  for (int g = 0; g < ng; ++g) {
    long long nbk = goff[g + 1] - goff[g];
    boff[g + 1] = boff[g] + (nbk > 1 ? nbk * nbk : 0);
  }
  PoolBuf didx(h->stream), dgoff(h->stream), dboff(h->stream), dwork(h->stream), du(h->stream), ddinv(h->stream),
      dbad(h->stream), dfocal(h->stream), dfg(h->stream), dg(h->stream);
  cudaError_t e = didx.alloc(nm * 4);
  if (e == cudaSuccess) e = dgoff.alloc((size_t)(ng + 1) * 4);
  if (e == cudaSuccess) e = dboff.alloc((size_t)(ng + 1) * 8);
  if (e == cudaSuccess) e = dwork.alloc((size_t)boff[ng] * 2 * 8);
  if (e == cudaSuccess) e = du.alloc(nm * 8);
  if (e == cudaSuccess) e = ddinv.alloc(nm * 8);
  if (e == cudaSuccess) e = dbad.alloc(4);
  if (e == cudaSuccess && nm) e = cudaMemcpyAsync(didx.p, members.data(), nm * 4, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dgoff.p, goff.data(), (size_t)(ng + 1) * 4, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dboff.p, boff.data(), (size_t)(ng + 1) * 8, cudaMemcpyHostToDevice, h->stream);
  const int big = 0x7fffffff;
  if (e == cudaSuccess) e = cudaMemcpyAsync(dbad.p, &big, 4, cudaMemcpyHostToDevice, h->stream);
  std::vector<double> u(nm), dinv(nm);
  int bad = big;
  if (e == cudaSuccess && ng > 0) {
    leaveout_blocks_kernel<<<ng, 128, 0, h->stream>>>(h->bufC, (size_t)h->Np, (const int*)didx.p, (const int*)dgoff.p,
                                                      (const long long*)dboff.p, ng, h->da, (double*)dwork.p,
                                                      (double*)du.p, (double*)ddinv.p, (int*)dbad.p);
    h->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess && nm) e = cudaMemcpyAsync(u.data(), du.p, nm * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && nm) e = cudaMemcpyAsync(dinv.data(), ddinv.p, nm * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, dbad.p, 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  }
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "leave-out prediction failed: %s", cudaGetErrorString(e));
  if (bad != big) return set_err(GPEDM_ERR_STATE, "leave-out block %d of Sigma^-1 is not positive definite", bad);
  for (int f = 0; f < nf; ++f) {
    const int g = fgroup[f], b0 = goff[g], nbk = goff[g + 1] - b0;
    int pos = -1;
    for (int r = 0; r < nbk; ++r)
      if (members[b0 + r] == focal[f]) pos = r;
    if (pos < 0) {
      mean[f] = NAN;
      var[f] = NAN;
      continue;
    }
    mean[f] = h->hY[focal[f]] - u[b0 + pos];
    var[f] = dinv[b0 + pos] - h->cur_ve;
  }
  if (gpgrad && nf > 0) {
    e = dfocal.alloc((size_t)nf * 4);
    if (e == cudaSuccess) e = dfg.alloc((size_t)nf * 4);
    if (e == cudaSuccess) e = dg.alloc((size_t)nf * h->d * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dfocal.p, focal.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dfg.p, fgroup.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
      loo_gpgrad_kernel<<<nf, 256, 0, h->stream>>>(h->bufC, (size_t)h->Np, (int)N, h->dX, h->dpop, h->drhotab, h->da,
                                                   h->dparams, (const int*)dfocal.p, (const int*)dfg.p,
                                                   (const int*)didx.p, (const int*)dgoff.p, (const double*)du.p, nf,
                                                   (double*)dg.p);
      h->launches++;
      e = cudaGetLastError();
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(gpgrad, dg.p, (size_t)nf * h->d * 8, cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "leave-out gradient failed: %s", cudaGetErrorString(e));
  }
  return 0;
}

static int gpedm_predict_loo_impl(gpedm_handle h, int32_t exclradius, const double* time, const int32_t* primary,
                                 double* mean, double* var, double* gpgrad) {
  if (!h || !mean || !var) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (!h->have_full) return set_err(GPEDM_ERR_STATE, "no full evaluation resident");
  if (exclradius < 0) return set_err(GPEDM_ERR_ARG, "exclradius must be >= 0");
  const int N = (int)h->N;
  int nd = 0;
  for (int i = 0; i < N; ++i) nd += primary ? (primary[i] != 0) : 1;
  // reference :1061, :1067 - primary rows are assumed to come first
  for (int i = 0; i < nd; ++i)
    if (primary && !primary[i]) return set_err(GPEDM_ERR_ARG, "primary rows must precede augmentation rows");
  if (nd < N && !time) return set_err(GPEDM_ERR_ARG, "time is required when augmentation rows are present");
  std::vector<int> members, goff(1, 0), focal(nd), fgroup(nd);
  for (int i = 0; i < nd; ++i) {
    size_t start = members.size();
    for (int j = i - exclradius; j <= i + exclradius; ++j)  // :1069-1070
      if (j >= 0 && j < N && h->hpop[j] == h->hpop[i]) members.push_back(j);
    size_t nex = members.size();
    for (int q = nd; q < N; ++q) {  // :1072 duplicates among augmentation rows
      bool dup = false;
      for (size_t e = start; e < nex && !dup; ++e)
        dup = (h->hpop[q] == h->hpop[members[e]]) && (time[q] == time[members[e]]);
      if (dup && std::find(members.begin() + start, members.end(), q) == members.end()) members.push_back(q);
    }
    goff.push_back((int)members.size());
    focal[i] = i;
    fgroup[i] = i;
  }
  return leaveout_core(h, members, goff, focal, fgroup, mean, var, gpgrad);
}

static int gpedm_predict_lto_impl(gpedm_handle h, int32_t exclradius, const double* time, const int32_t* primary,
                                 double* mean, double* var, double* gpgrad) {
  if (!h || !mean || !var || !time) return set_err(GPEDM_ERR_ARG, "NULL argument (time is required)");
  if (!h->have_full) return set_err(GPEDM_ERR_STATE, "no full evaluation resident");
  if (exclradius < 0) return set_err(GPEDM_ERR_ARG, "exclradius must be >= 0");
  const int N = (int)h->N;
  int nd = 0;
  for (int i = 0; i < N; ++i) nd += primary ? (primary[i] != 0) : 1;
  for (int i = 0; i < nd; ++i)
    if (primary && !primary[i]) return set_err(GPEDM_ERR_ARG, "primary rows must precede augmentation rows");
  // timevals = unique(Time) in order of first appearance (:1142)
  std::vector<double> timevals;
  std::vector<int> tcode(N);
  for (int i = 0; i < N; ++i) {
    int c = -1;
    for (size_t q = 0; q < timevals.size(); ++q)
      if (timevals[q] == time[i]) {
        c = (int)q;
        break;
      }
    if (c < 0) {
      c = (int)timevals.size();
      timevals.push_back(time[i]);
    }
    tcode[i] = c;
  }
  const int nt = (int)timevals.size();
  std::vector<int> members, goff(1, 0), focal, fgroup;
  for (int i = 0; i < nd; ++i) {
    mean[i] = NAN;
    var[i] = NAN;
  }
  std::vector<int> fpos;
  for (int t = 0; t < nt; ++t) {
    int lo = std::max(0, t - exclradius), hi = std::min(nt - 1, t + exclradius);  // :1155-1156
    for (int j = 0; j < N; ++j)
      if (tcode[j] >= lo && tcode[j] <= hi) members.push_back(j);  // complement of `train`, :1158
    goff.push_back((int)members.size());
    for (int j = 0; j < nd; ++j)
      if (tcode[j] == t) {  // :1157 focal rows among primary
        focal.push_back(j);
        fgroup.push_back(t);
        fpos.push_back(j);
      }
  }
  const int nf = (int)focal.size();
  std::vector<double> m(nf), v(nf), g;
  if (gpgrad) g.resize((size_t)nf * h->d);
  int rc = leaveout_core(h, members, goff, focal, fgroup, m.data(), v.data(), gpgrad ? g.data() : nullptr);
  if (rc) return rc;
  for (int f = 0; f < nf; ++f) {
    mean[fpos[f]] = m[f];
    var[fpos[f]] = v[f];
    if (gpgrad)
      for (int k = 0; k < h->d; ++k) gpgrad[fpos[f] + (size_t)k * nd] = g[f + (size_t)k * nf];
  }
  return 0;
}

static int gpedm_predict_sequential_impl(gpedm_handle h, const double* time, double* ymean, double* ysd2) {
  if (!h || !time || !ymean || !ysd2) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (!h->have_full) return set_err(GPEDM_ERR_STATE, "no full evaluation resident");
  CU(cudaSetDevice(h->device));
  const int N = (int)h->N, d = h->d;
  // time-major stable ordering by order of first appearance of each time value (:1201)
  std::vector<double> timevals;
  std::vector<int> tcode(N);
  for (int i = 0; i < N; ++i) {
    int c = -1;
    for (size_t q = 0; q < timevals.size(); ++q)
      if (timevals[q] == time[i]) {
        c = (int)q;
        break;
      }
    if (c < 0) {
      c = (int)timevals.size();
      timevals.push_back(time[i]);
    }
    tcode[i] = c;
  }
  std::vector<int> perm(N);
  for (int i = 0; i < N; ++i) perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return tcode[a] < tcode[b]; });
  std::vector<double> Xh((size_t)N * d), Xp((size_t)N * d), Yp(N);
  std::vector<int> popp(N), bstart(N);
  CU(cudaMemcpy(Xh.data(), h->dX, (size_t)N * d * 8, cudaMemcpyDeviceToHost));
  for (int r = 0; r < N; ++r) {
    int s = perm[r];
    Yp[r] = h->hY[s];
    popp[r] = h->hpop[s];
    for (int k = 0; k < d; ++k) Xp[r + (size_t)k * N] = Xh[s + (size_t)k * N];
  }
  for (int r = 0; r < N; ++r) bstart[r] = (r > 0 && tcode[perm[r]] == tcode[perm[r - 1]]) ? bstart[r - 1] : r;
  std::vector<double> rhomat;
  if (h->has_rhomat) {
    rhomat.resize((size_t)h->np * h->np);
    CU(cudaMemcpy(rhomat.data(), h->drhotab, rhomat.size() * 8, cudaMemcpyDeviceToHost));
  }
  gpedm_handle t = nullptr;
  int rc = gpedm_create(h->device, N, d, h->np, Xp.data(), Yp.data(), popp.data(), h->has_rhomat ? rhomat.data() : nullptr, &t);
  if (rc) return rc;
  *t->hparams = *h->hparams;
  rc = enqueue_eval(t, false, true);
  int* dbs = nullptr;
  double *dm = nullptr, *ds = nullptr;
  cudaError_t e = cudaSuccess;
  if (!rc) {
    e = cudaMalloc((void**)&dbs, (size_t)N * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dm, (size_t)N * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ds, (size_t)N * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dbs, bstart.data(), (size_t)N * 4, cudaMemcpyHostToDevice, t->stream);
    if (e == cudaSuccess) {
      sequential_kernel<<<ceil_div(N, 8), 256, 0, t->stream>>>(t->bufA, (size_t)t->Np, N, t->dz, dbs, dm, ds);
      t->launches++;
      e = cudaStreamSynchronize(t->stream);
    }
  }
  std::vector<double> mp(N), sp(N);
  if (!rc && e == cudaSuccess) e = cudaMemcpy(mp.data(), dm, (size_t)N * 8, cudaMemcpyDeviceToHost);
  if (!rc && e == cudaSuccess) e = cudaMemcpy(sp.data(), ds, (size_t)N * 8, cudaMemcpyDeviceToHost);
  int info = rc ? 0 : *t->hinfo;
  h->launches += t->launches;
  cudaFree(dbs);
  cudaFree(dm);
  cudaFree(ds);
  gpedm_destroy(t);
  if (rc) return rc;
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "sequential prediction failed: %s", cudaGetErrorString(e));
  if (info > 0) return set_err(info, "time-ordered covariance matrix not positive definite at row %d", info);
  const double s2ve = h->cur_sigma2 + h->cur_ve;
  for (int r = 0; r < N; ++r) {
    int s = perm[r];
    if (tcode[s] == 0) {  // first time value: never updated (:1215-1216, :1236)
      ymean[s] = 0.0;
      ysd2[s] = s2ve;
    } else {
      ymean[s] = mp[r];
      ysd2[s] = sp[r];
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------- batched fits
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static bool load_nccl(NcclApi& n) {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.lib) break;
  }
  if (!n.lib) return false;
  n.CommInitAll = (int (*)(void**, int, const int*))dlsym(n.lib, "ncclCommInitAll");
  n.CommDestroy = (int (*)(void*))dlsym(n.lib, "ncclCommDestroy");
  n.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(n.lib, "ncclAllGather");
  n.GroupStart = (int (*)())dlsym(n.lib, "ncclGroupStart");
  n.GroupEnd = (int (*)())dlsym(n.lib, "ncclGroupEnd");
  n.GetErrorString = (const char* (*)(int))dlsym(n.lib, "ncclGetErrorString");
  return n.CommInitAll && n.CommDestroy && n.AllGather && n.GroupStart && n.GroupEnd;
}


// ------------------------------------------------------------------------------- batched small fits
// All fits of a batch with N <= 128 advance in lockstep: one small_eval_kernel launch per Rprop
// iteration evaluates every still-active fit (one CTA each); the scalar Rprop logic of
// fmingrad_Rprop (reference R/GPfunctions.R:522-584) runs on the host per fit exactly as in
// gpedm_fit_rprop.  A last launch with full = 1 at each fit's final parameters yields a and
// diag(Sigma^-1), from which the in-sample and leave-one-out vectors follow in O(N).
static int fit_batch_small(int device, const std::vector<int>& idx, const gpedm_fit_desc* descs,
                           const gpedm_rprop_options* opts_in, gpedm_result* results) {
  const int nf = (int)idx.size();
  if (nf == 0) return 0;
  CU(cudaSetDevice(device));
  int rc = ensure_kernel_attrs(device);
  if (rc) return rc;
  gpedm_rprop_options o;
  if (opts_in)
    o = *opts_in;
  else
    gpedm_default_rprop_options(&o);
  const double eta_minus = o.eta_minus - 1, eta_plus = o.eta_plus - 1;
  // ---- pack inputs
  std::vector<size_t> offX(nf), offN(nf), offR(nf);
  size_t totX = 0, totN = 0, totR = 0;
  int dmax = 1;
  for (int q = 0; q < nf; ++q) {
    const gpedm_fit_desc& ds = descs[idx[q]];
    if (ds.N < 1 || ds.N > TB || ds.d < 1 || ds.d > GPEDM_MAX_D || ds.np < 1 || !ds.X || !ds.Y || !ds.pop || !ds.initparst)
      return set_err(GPEDM_ERR_ARG, "bad descriptor for fit %d", idx[q]);
    for (int64_t i = 0; i < ds.N; ++i)
      if (ds.pop[i] < 0 || ds.pop[i] >= ds.np) return set_err(GPEDM_ERR_ARG, "fit %d: pop code outside 0..np-1", idx[q]);
    offX[q] = totX;
    offN[q] = totN;
    offR[q] = totR;
    totX += (size_t)ds.N * ds.d;
    totN += (size_t)ds.N;
    totR += ds.rhomat ? (size_t)ds.np * ds.np : 0;
    dmax = std::max(dmax, (int)ds.d);
  }
  std::vector<double> hX(totX), hY(totN), hR(std::max<size_t>(totR, 1));
  std::vector<int> hP(totN);
  for (int q = 0; q < nf; ++q) {
    const gpedm_fit_desc& ds = descs[idx[q]];
    memcpy(&hX[offX[q]], ds.X, (size_t)ds.N * ds.d * 8);
    memcpy(&hY[offN[q]], ds.Y, (size_t)ds.N * 8);
    memcpy(&hP[offN[q]], ds.pop, (size_t)ds.N * 4);
    if (ds.rhomat) memcpy(&hR[offR[q]], ds.rhomat, (size_t)ds.np * ds.np * 8);
  }
  DevBuf dX, dY, dP, dR, dFits, dParams, dScal, dA, dDiag, dInfo, dActive;
  PinnedBuf pParams, pScal, pInfo;
  cudaError_t e = dX.alloc(totX * 8);
  if (e == cudaSuccess) e = dY.alloc(totN * 8);
  if (e == cudaSuccess) e = dP.alloc(totN * 4);
  if (e == cudaSuccess) e = dR.alloc(totR * 8);
  if (e == cudaSuccess) e = dFits.alloc(sizeof(SmallFit) * nf);
  if (e == cudaSuccess) e = dParams.alloc(sizeof(EvalParams) * nf);
  if (e == cudaSuccess) e = dScal.alloc((size_t)NSCAL * 8 * nf);
  if (e == cudaSuccess) e = dA.alloc(totN * 8);
  if (e == cudaSuccess) e = dDiag.alloc(totN * 8);
  if (e == cudaSuccess) e = dInfo.alloc((size_t)nf * 4);
  if (e == cudaSuccess) e = dActive.alloc((size_t)nf * 4);
  if (e == cudaSuccess) e = pParams.alloc(sizeof(EvalParams) * nf);
  if (e == cudaSuccess) e = pScal.alloc((size_t)NSCAL * 8 * nf);
  if (e == cudaSuccess) e = pInfo.alloc((size_t)nf * 4);
  if (e == cudaSuccess) e = cudaMemcpy(dX.p, hX.data(), totX * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dY.p, hY.data(), totN * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dP.p, hP.data(), totN * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && totR) e = cudaMemcpy(dR.p, hR.data(), totR * 8, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "small-fit batch setup failed: %s", cudaGetErrorString(e));
  EvalParams* hparams = (EvalParams*)pParams.p;
  double* hscal = (double*)pScal.p;
  int* hinfo = (int*)pInfo.p;
  std::vector<SmallFit> hfits(nf);
  // light host-side stand-ins for a handle: only the fields the scalar head / tail of an evaluation reads
  std::vector<gpedm_fit_s> meta(nf);
  for (int q = 0; q < nf; ++q) {
    const gpedm_fit_desc& ds = descs[idx[q]];
    SmallFit& F = hfits[q];
    F.X = (const double*)dX.p + offX[q];
    F.Y = (const double*)dY.p + offN[q];
    F.pop = (const int*)dP.p + offN[q];
    F.rhotab = ds.rhomat ? (const double*)dR.p + offR[q] : nullptr;
    F.scal = (double*)dScal.p + (size_t)q * NSCAL;
    F.a_out = (double*)dA.p + offN[q];
    F.diag_out = (double*)dDiag.p + offN[q];
    F.info = (int*)dInfo.p + q;
    F.N = (int)ds.N;
    F.d = ds.d;
    F.np = ds.np;
    F.pad = 0;
    gpedm_fit_s& m = meta[q];
    m.N = ds.N;
    m.d = ds.d;
    m.np = ds.np;
    m.p = ds.d + 3;
    m.has_rhomat = ds.rhomat != nullptr;
    m.single_pop = true;
    for (int64_t i = 1; i < ds.N; ++i)
      if (ds.pop[i] != ds.pop[0]) m.single_pop = false;
    m.hparams = hparams + q;
    m.hscal = hscal + (size_t)q * NSCAL;
  }
  CU(cudaMemcpy(dFits.p, hfits.data(), sizeof(SmallFit) * nf, cudaMemcpyHostToDevice));
  cudaStream_t st;
  CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const int DC = dmax <= 8 ? 8 : (dmax <= 16 ? 16 : 32);
  std::vector<int> active;
  auto launch = [&](int full) -> int {
    const int na = (int)active.size();
    if (na == 0) return 0;
    cudaError_t e2 = cudaMemcpyAsync(dParams.p, hparams, sizeof(EvalParams) * nf, cudaMemcpyHostToDevice, st);
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(dActive.p, active.data(), (size_t)na * 4, cudaMemcpyHostToDevice, st);
    if (e2 != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "small-fit upload failed: %s", cudaGetErrorString(e2));
    if (DC == 8)
      small_eval_kernel<8><<<na, DG_THREADS, small_eval_smem_bytes<8>(), st>>>((const SmallFit*)dFits.p, (const EvalParams*)dParams.p, (const int*)dActive.p, full);
    else if (DC == 16)
      small_eval_kernel<16><<<na, DG_THREADS, small_eval_smem_bytes<16>(), st>>>((const SmallFit*)dFits.p, (const EvalParams*)dParams.p, (const int*)dActive.p, full);
    else
      small_eval_kernel<32><<<na, DG_THREADS, small_eval_smem_bytes<32>(), st>>>((const SmallFit*)dFits.p, (const EvalParams*)dParams.p, (const int*)dActive.p, full);
    e2 = cudaMemcpyAsync(hscal, dScal.p, (size_t)NSCAL * 8 * nf, cudaMemcpyDeviceToHost, st);
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(hinfo, dInfo.p, (size_t)nf * 4, cudaMemcpyDeviceToHost, st);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
    if (e2 == cudaSuccess) e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "small-fit evaluation failed: %s", cudaGetErrorString(e2));
    return 0;
  };
  struct State {
    double pa[GPEDM_MAX_P], del[GPEDM_MAX_P], grad[GPEDM_MAX_P], panew[GPEDM_MAX_P];
    double nll = 0, s = 0, df = 10;
    int iter = 0, nevals = 0, status = 0;
    bool done = false;
  };
  std::vector<State> stt(nf);
  std::vector<HostPars> hp(nf);
  // ---- initial evaluation (:545)
  for (int q = 0; q < nf; ++q) {
    const gpedm_fit_desc& ds = descs[idx[q]];
    for (int i = 0; i < ds.d + 3; ++i) {
      stt[q].pa[i] = stt[q].panew[i] = ds.initparst[i];
      stt[q].del[i] = o.delta0;
    }
    host_transform(&meta[q], stt[q].panew, ds.modeprior, ds.fixedpars, ds.rhofixed, hp[q]);
    fill_params(&meta[q], hp[q]);
    active.push_back(q);
  }
  rc = launch(0);
  bool first = true;
  while (rc == 0) {
    std::vector<int> next_active;
    for (int q : active) {
      State& S = stt[q];
      const gpedm_fit_desc& ds = descs[idx[q]];
      gpedm_result r;
      finish_result(&meta[q], hp[q], false, &r);
      S.nevals++;
      if (hinfo[q] > 0) {
        S.status = hinfo[q];
        S.done = true;
        continue;
      }
      const int p = ds.d + 3;
      if (first) {
        S.nll = r.nllpost;
        for (int i = 0; i < p; ++i) S.grad[i] = r.grad[i];
        S.s = r.gradnorm;
      } else {
        S.s = r.gradnorm;
        S.df = fabs(r.nllpost / S.nll - 1);  // :565
        for (int i = 0; i < p; ++i) {
          const double gc = S.grad[i] * r.grad[i];
          const double tdel = S.del[i] * (1 + eta_plus * (gc > 0) + eta_minus * (gc < 0));  // :569
          S.del[i] = tdel < o.deltamin ? o.deltamin : (tdel > o.deltamax ? o.deltamax : tdel);
          S.pa[i] = S.panew[i];
          S.grad[i] = r.grad[i];
        }
        S.nll = r.nllpost;
        S.iter++;
      }
      if (S.s > o.tol_grad && S.iter < o.maxiter && S.df > o.tol_df) {  // :556
        for (int i = 0; i < p; ++i) S.panew[i] = S.pa[i] - sgn(S.grad[i]) * S.del[i];  // :559
        host_transform(&meta[q], S.panew, ds.modeprior, ds.fixedpars, ds.rhofixed, hp[q]);
        fill_params(&meta[q], hp[q]);
        next_active.push_back(q);
      } else {
        S.done = true;
      }
    }
    first = false;
    active.swap(next_active);
    if (active.empty()) break;
    rc = launch(0);
  }
  // ---- final full evaluation at each fit's last parameters (:580)
  std::vector<double> ha(totN), hd(totN);
  if (rc == 0) {
    active.clear();
    for (int q = 0; q < nf; ++q) {
      if (stt[q].status != 0) continue;
      const gpedm_fit_desc& ds = descs[idx[q]];
      host_transform(&meta[q], stt[q].pa, ds.modeprior, ds.fixedpars, ds.rhofixed, hp[q]);
      fill_params(&meta[q], hp[q]);
      active.push_back(q);
    }
    rc = launch(1);
    if (rc == 0) {
      e = cudaMemcpy(ha.data(), dA.p, totN * 8, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(hd.data(), dDiag.p, totN * 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) rc = set_err(GPEDM_ERR_CUDA, "small-fit download failed: %s", cudaGetErrorString(e));
    }
  }
  cudaStreamDestroy(st);
  if (rc) return rc;
  int first_bad = 0;
  for (int q = 0; q < nf; ++q) {
    const gpedm_fit_desc& ds = descs[idx[q]];
    gpedm_result& R = results[idx[q]];
    State& S = stt[q];
    if (S.status == 0 && hinfo[q] > 0) S.status = hinfo[q];
    if (S.status != 0) {
      memset(&R, 0, sizeof(R));
      R.status = S.status;
      R.iter = S.iter;
      R.nevals = S.nevals;
      R.p = ds.d + 3;
      if (!first_bad) first_bad = S.status;
      continue;
    }
    finish_result(&meta[q], hp[q], true, &R);
    R.iter = S.iter;
    R.nevals = S.nevals + 1;
    R.status = 0;
    const double ve = hp[q].ve;
    for (int64_t i = 0; i < ds.N; ++i) {
      const double a = ha[offN[q] + i], sii = hd[offN[q] + i], y = ds.Y[i];
      if (ds.want_loo && ds.loo_mean) ds.loo_mean[i] = y - a / sii;
      if (ds.want_loo && ds.loo_var) ds.loo_var[i] = 1.0 / sii - ve;
      if (ds.insamp_mean) ds.insamp_mean[i] = y - ve * a;
      if (ds.insamp_var) ds.insamp_var[i] = ve - ve * ve * sii;
    }
  }
  if (first_bad) return set_err(first_bad, "at least one small fit failed: covariance matrix not positive definite (pivot %d)", first_bad);
  return 0;
}

// Fixed-width record exchanged by the final all-gather.
struct BatchRecord {
  int32_t fit_index;
  int32_t valid;
  gpedm_result res;
  gpedm_fit_stats st;  // filled by gpedm_fit_grid only
};

// Every GPU ends up with every fit's fixed-width record.  One GPU: a copy.  Several: ONE NCCL
// all-gather (single process, one communicator per GPU); every GPU contributes `slab` records
// (padded with valid=0).  No other collective exists on the path.
static int exchange_records(const std::vector<int>& devs, const std::vector<std::vector<BatchRecord>>& per_gpu,
                            int nfits, gpedm_result* results, gpedm_fit_stats* stats) {
  const int ngpus = (int)devs.size();
  if (ngpus == 1) {
    for (auto& r : per_gpu[0]) {
      results[r.fit_index] = r.res;
      if (stats) stats[r.fit_index] = r.st;
    }
    return 0;
  }
  NcclApi nc;
  if (!load_nccl(nc)) return set_err(GPEDM_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
  size_t slab = 0;
  for (int g = 0; g < ngpus; ++g) slab = std::max(slab, per_gpu[g].size());
  if (slab == 0) slab = 1;
  const size_t slab_bytes = slab * sizeof(BatchRecord);
  std::vector<void*> comms(ngpus, nullptr);
  int nrc = nc.CommInitAll(comms.data(), ngpus, devs.data());
  if (nrc != 0)
    return set_err(GPEDM_ERR_NCCL, "ncclCommInitAll failed: %s", nc.GetErrorString ? nc.GetErrorString(nrc) : "?");
  std::vector<char*> dsend(ngpus, nullptr), drecv(ngpus, nullptr);
  std::vector<cudaStream_t> sts(ngpus, nullptr);
  cudaError_t e = cudaSuccess;
  for (int g = 0; g < ngpus && e == cudaSuccess; ++g) {
    e = cudaSetDevice(devs[g]);
    if (e == cudaSuccess) e = cudaStreamCreate(&sts[g]);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dsend[g], slab_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&drecv[g], slab_bytes * ngpus);
    std::vector<BatchRecord> pad(slab);
    memset(pad.data(), 0, slab_bytes);
    for (size_t q = 0; q < per_gpu[g].size(); ++q) pad[q] = per_gpu[g][q];
    if (e == cudaSuccess) e = cudaMemcpy(dsend[g], pad.data(), slab_bytes, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    nc.GroupStart();
    for (int g = 0; g < ngpus; ++g) {
      cudaSetDevice(devs[g]);
      nrc = nc.AllGather(dsend[g], drecv[g], slab_bytes, /*ncclChar*/ 0, comms[g], sts[g]);
      if (nrc != 0) break;
    }
    int nrc2 = nc.GroupEnd();
    if (nrc == 0) nrc = nrc2;
    for (int g = 0; g < ngpus && e == cudaSuccess; ++g) {
      cudaSetDevice(devs[g]);
      e = cudaStreamSynchronize(sts[g]);
    }
  }
  std::vector<BatchRecord> all(slab * ngpus);
  if (e == cudaSuccess && nrc == 0) {
    cudaSetDevice(devs[0]);
    e = cudaMemcpy(all.data(), drecv[0], slab_bytes * ngpus, cudaMemcpyDeviceToHost);
  }
  for (int g = 0; g < ngpus; ++g) {
    cudaSetDevice(devs[g]);
    cudaFree(dsend[g]);
    cudaFree(drecv[g]);
    if (sts[g]) cudaStreamDestroy(sts[g]);
    if (comms[g]) nc.CommDestroy(comms[g]);
  }
  if (nrc != 0) return set_err(GPEDM_ERR_NCCL, "ncclAllGather failed: %s", nc.GetErrorString ? nc.GetErrorString(nrc) : "?");
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "batch gather failed: %s", cudaGetErrorString(e));
  int got = 0;
  for (auto& r : all)
    if (r.valid && r.fit_index >= 0 && r.fit_index < nfits) {
      results[r.fit_index] = r.res;
      if (stats) stats[r.fit_index] = r.st;
      ++got;
    }
  if (got != nfits) return set_err(GPEDM_ERR_STATE, "batch gather returned %d of %d records", got, nfits);
  return 0;
}

static int run_one_fit(int device, const gpedm_fit_desc& ds, const gpedm_rprop_options* opts, gpedm_result* res) {
  gpedm_handle h = nullptr;
  int rc = gpedm_create(device, ds.N, ds.d, ds.np, ds.X, ds.Y, ds.pop, ds.rhomat, &h);
  if (rc) {
    memset(res, 0, sizeof(*res));
    res->status = rc;
    return rc;
  }
  rc = gpedm_fit_rprop(h, ds.initparst, ds.modeprior, ds.fixedpars, ds.rhofixed, opts, res);
  if (rc == 0) {
    if (ds.want_loo && ds.loo_mean) rc = gpedm_get_vector(h, GPEDM_VEC_LOO_MEAN, ds.loo_mean);
    if (rc == 0 && ds.want_loo && ds.loo_var) rc = gpedm_get_vector(h, GPEDM_VEC_LOO_VAR, ds.loo_var);
    if (rc == 0 && ds.insamp_mean) rc = gpedm_get_vector(h, GPEDM_VEC_MEAN, ds.insamp_mean);
    if (rc == 0 && ds.insamp_var) rc = gpedm_get_vector(h, GPEDM_VEC_VAR, ds.insamp_var);
  }
  res->status = rc;
  gpedm_destroy(h);
  return rc;
}

static int gpedm_fit_batch_impl(int32_t nfits, const gpedm_fit_desc* descs, int32_t ngpus, const int32_t* devices,
                               const gpedm_rprop_options* opts, gpedm_result* results) {
  if (nfits < 0 || (nfits > 0 && (!descs || !results))) return set_err(GPEDM_ERR_ARG, "bad arguments");
  if (nfits == 0) return 0;
  int ndev = gpedm_device_count();
  if (ndev == 0) return set_err(GPEDM_ERR_CUDA, "no CUDA device available (libgpedm_b200 has no CPU fallback)");
  if (ngpus < 1 || ngpus > ndev) return set_err(GPEDM_ERR_ARG, "ngpus=%d out of range (have %d devices)", ngpus, ndev);
  std::vector<int> devs(ngpus);
  for (int g = 0; g < ngpus; ++g) {
    devs[g] = devices ? devices[g] : g;
    if (devs[g] < 0 || devs[g] >= ndev) return set_err(GPEDM_ERR_ARG, "device %d out of range", devs[g]);
  }
  // queue ordered by descending estimated cost ~ N^3 (+ d N^2)
  std::vector<int> order(nfits);
  for (int i = 0; i < nfits; ++i) order[i] = i;
  auto cost = [&](int i) {
    double n = (double)descs[i].N;
    return n * n * n + 40.0 * descs[i].d * n * n;
  };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
  // Fits with N <= 128 (one tile) advance in lockstep through the batched one-CTA-per-fit kernel,
  // one host thread per GPU over a strided share; everything larger goes through the tiled
  // pipeline, pulled from a shared queue by W worker threads per GPU.  Both kinds run concurrently.
  std::vector<int> small_order, big_order;
  for (int q = 0; q < nfits; ++q) (descs[order[q]].N <= TB && !g_no_small_batch ? small_order : big_order).push_back(order[q]);
  if (small_order.size() < 2) {  // a single small fit gains nothing from the batched kernel
    big_order = order;
    small_order.clear();
  }
  const int nbig = (int)big_order.size();
  std::atomic<int> next(0);
  std::vector<std::vector<BatchRecord>> per_gpu(ngpus);
  std::vector<std::string> errs(ngpus);
  // W worker threads per GPU, each with its own handles / streams: independent fits interleave
  // on one device, so the latency-bound stretches of one evaluation (the serial Cholesky chain)
  // are filled by the bulk kernels of the other.
  int64_t nmax = 0;
  for (int i : big_order) nmax = std::max<int64_t>(nmax, descs[i].N);
  // small problems leave most of the GPU idle -> interleave several fits; large ones already fill it
  const int W = g_batch_workers > 0 ? g_batch_workers : (nmax <= 4096 ? 3 : 1);
  std::vector<std::vector<BatchRecord>> per_worker(ngpus * (W + 1));  // slot W of each GPU: the small-fit thread
  std::vector<std::string> werrs(ngpus * (W + 1));
  auto worker = [&](int g, int w) {
    const int wi = g * (W + 1) + w;
    cudaSetDevice(devs[g]);
    for (;;) {
      int q = next.fetch_add(1);
      if (q >= nbig) break;
      int i = big_order[q];
      BatchRecord rec;
      memset(&rec, 0, sizeof(rec));
      rec.fit_index = i;
      rec.valid = 1;
      int rc = run_one_fit(devs[g], descs[i], opts, &rec.res);
      if (rc != 0 && werrs[wi].empty()) werrs[wi] = gpedm_last_error();
      per_worker[wi].push_back(rec);
    }
  };
  auto small_worker = [&](int g) {
    const int wi = g * (W + 1) + W;
    std::vector<int> mine;
    for (size_t q = g; q < small_order.size(); q += ngpus) mine.push_back(small_order[q]);
    if (mine.empty()) return;
    std::vector<gpedm_result> tmp(nfits);
    int rc = fit_batch_small(devs[g], mine, descs, opts, tmp.data());
    if (rc != 0) werrs[wi] = gpedm_last_error();
    for (int i : mine) {
      BatchRecord rec;
      memset(&rec, 0, sizeof(rec));
      rec.fit_index = i;
      rec.valid = 1;
      rec.res = tmp[i];
      if (rc < 0 && rec.res.status == 0) rec.res.status = rc;
      per_worker[wi].push_back(rec);
    }
  };
  {
    std::vector<std::thread> th;
    try {
      if (!small_order.empty())
        for (int g = 0; g < ngpus; ++g) th.emplace_back(small_worker, g);
      if (nbig > 0)
        for (int g = 0; g < ngpus; ++g)
          for (int w = 0; w < W; ++w) th.emplace_back(worker, g, w);
    } catch (...) {
      next.store(nbig);  // could not start every thread: let the started ones drain and stop
      for (auto& t : th) t.join();
      return set_err(GPEDM_ERR_STATE, "could not start batch worker threads");
    }
    for (auto& t : th) t.join();
  }
  for (int wi = 0; wi < ngpus * (W + 1); ++wi) {
    const int g = wi / (W + 1);
    per_gpu[g].insert(per_gpu[g].end(), per_worker[wi].begin(), per_worker[wi].end());
    if (errs[g].empty() && !werrs[wi].empty()) errs[g] = werrs[wi];
  }
  {
    size_t got = 0;
    for (int g = 0; g < ngpus; ++g) got += per_gpu[g].size();
    if (got != (size_t)nfits) return set_err(GPEDM_ERR_STATE, "batch produced %zu of %d records", got, nfits);
  }

  {
    int rc = exchange_records(devs, per_gpu, nfits, results, nullptr);
    if (rc) return rc;
  }
  for (int g = 0; g < ngpus; ++g)
    if (!errs[g].empty()) {
      int first_bad = 0;
      for (int i = 0; i < nfits; ++i)
        if (results[i].status != 0) {
          first_bad = results[i].status;
          break;
        }
      return set_err(first_bad ? first_bad : GPEDM_ERR_STATE, "at least one fit failed: %s", errs[g].c_str());
    }
  return 0;
}

// ------------------------------------------------------------------------------- lagged series / grids
// Raw series resident on one device, shared by every grid cell fitted there.
struct LagSeries {
  int device = 0;
  int T = 0, np = 1;
  StreamGuard sg;  // declared before the buffers: destroyed after them
  PoolBuf y, pop, rank, poprows, popoff;
  std::vector<int> hpop;
};

static int upload_series(int device, int64_t T, const double* y, const int32_t* pop, int32_t np, LagSeries& S) {
  if (!y || T < 2 || T > (1 << 24)) return set_err(GPEDM_ERR_ARG, "bad series (T=%lld)", (long long)T);
  if (np < 1) return set_err(GPEDM_ERR_ARG, "np=%d must be >= 1", np);
  S.device = device;
  S.T = (int)T;
  S.np = np;
  S.hpop.assign((size_t)T, 0);
  if (pop)
    for (int64_t t = 0; t < T; ++t) {
      if (pop[t] < 0 || pop[t] >= np) return set_err(GPEDM_ERR_ARG, "pop[%lld]=%d outside 0..np-1", (long long)t, pop[t]);
      S.hpop[t] = pop[t];
    }
  // index bookkeeping only (no arithmetic on the data): position of each row inside its population
  std::vector<int> rank((size_t)T), popoff(np + 1, 0), poprows((size_t)T), fillp(np, 0);
  for (int64_t t = 0; t < T; ++t) popoff[S.hpop[t] + 1]++;
  for (int g = 0; g < np; ++g) popoff[g + 1] += popoff[g];
  for (int64_t t = 0; t < T; ++t) {
    const int g = S.hpop[t];
    rank[t] = fillp[g];
    poprows[popoff[g] + fillp[g]] = (int)t;
    fillp[g]++;
  }
  CU(cudaSetDevice(device));
  CU(S.sg.create());
  S.y.st = S.pop.st = S.rank.st = S.poprows.st = S.popoff.st = S.sg.s;
  cudaError_t e = S.y.alloc((size_t)T * 8);
  if (e == cudaSuccess) e = S.pop.alloc((size_t)T * 4);
  if (e == cudaSuccess) e = S.rank.alloc((size_t)T * 4);
  if (e == cudaSuccess) e = S.poprows.alloc((size_t)T * 4);
  if (e == cudaSuccess) e = S.popoff.alloc((size_t)(np + 1) * 4);
  cudaStream_t st = S.sg.s;
  if (e == cudaSuccess) e = cudaMemcpyAsync(S.y.p, y, (size_t)T * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(S.pop.p, S.hpop.data(), (size_t)T * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(S.rank.p, rank.data(), (size_t)T * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(S.poprows.p, poprows.data(), (size_t)T * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(S.popoff.p, popoff.data(), (size_t)(np + 1) * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "series upload failed: %s", cudaGetErrorString(e));
  return 0;
}

// Build the training set of fitGP(y, E, tau, scaling) from the resident series into a new handle.
static int create_from_series(const LagSeries& S, int32_t E, int32_t tau, int32_t scaling, const double* rhomat,
                              gpedm_handle* out) {
  *out = nullptr;
  if (E < 1 || E > GPEDM_MAX_D) return set_err(GPEDM_ERR_ARG, "E=%d out of range 1..%d", E, GPEDM_MAX_D);
  if (tau < 1) return set_err(GPEDM_ERR_ARG, "tau=%d must be >= 1", tau);
  if (scaling < GPEDM_SCALE_NONE || scaling > GPEDM_SCALE_LOCAL) return set_err(GPEDM_ERR_ARG, "unknown scaling %d", scaling);
  CU(cudaSetDevice(S.device));
  const int G = scaling == GPEDM_SCALE_LOCAL ? S.np : 1;
  const size_t nstat = (size_t)(E + 1) * G;
  StreamGuard sg;
  CU(sg.create());
  cudaStream_t st = sg.s;
  PoolBuf dcenter(st), dscale(st), drowmap(st), dcount(st);
  cudaError_t e = dcenter.alloc(nstat * 8);
  if (e == cudaSuccess) e = dscale.alloc(nstat * 8);
  if (e == cudaSuccess) e = drowmap.alloc((size_t)S.T * 4);
  if (e == cudaSuccess) e = dcount.alloc(4);
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "lag preparation failed: %s", cudaGetErrorString(e));
  LagPrep P;
  P.y = (const double*)S.y.p;
  P.pop = (const int*)S.pop.p;
  P.rank = (const int*)S.rank.p;
  P.poprows = (const int*)S.poprows.p;
  P.popoff = (const int*)S.popoff.p;
  P.center = (double*)dcenter.p;
  P.scale = (double*)dscale.p;
  P.T = S.T;
  P.E = E;
  P.tau = tau;
  P.np = S.np;
  P.G = G;
  P.mode = scaling;
  int N = 0;
  lag_stats_kernel<<<dim3(E + 1, G), 256, 0, st>>>(P);
  lag_compact_kernel<<<1, 1024, 0, st>>>(P, (int*)drowmap.p, (int*)dcount.p);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&N, dcount.p, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "lag preparation failed: %s", cudaGetErrorString(e));
  if (N < 1) return set_err(GPEDM_ERR_ARG, "no complete rows for E=%d tau=%d (series length %d)", E, tau, S.T);
  gpedm_fit_s* h = nullptr;
  int rc = alloc_handle(S.device, N, E, S.np, rhomat != nullptr, &h);
  if (rc) return rc;
  lag_fill_kernel<<<dim3(ceil_div(N, 256), E + 1), 256, 0, st>>>(P, (const int*)drowmap.p, N, h->dY, h->dX, h->dpop);
  h->launches += 3;
  h->hY.resize(N);
  h->hpop.resize(N);
  h->rows.resize(N);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->hY.data(), h->dY, (size_t)N * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->hpop.data(), h->dpop, (size_t)N * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->rows.data(), drowmap.p, (size_t)N * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && rhomat) e = cudaMemcpyAsync(h->drhotab, rhomat, (size_t)S.np * S.np * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    free_handle(h);
    return set_err(GPEDM_ERR_CUDA, "lag preparation failed: %s", cudaGetErrorString(e));
  }
  h->single_pop = true;
  for (int i = 1; i < N; ++i)
    if (h->hpop[i] != h->hpop[0]) h->single_pop = false;
  // the handle keeps the column statistics (for gpedm_get_fit_stats / gpedm_lagged_info)
  e = cudaMallocAsync((void**)&h->dcenter, nstat * 8, h->stream);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&h->dscale, nstat * 8, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->dcenter, dcenter.p, nstat * 8, cudaMemcpyDeviceToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->dscale, dscale.p, nstat * 8, cudaMemcpyDeviceToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    free_handle(h);
    return set_err(GPEDM_ERR_CUDA, "lag preparation failed: %s", cudaGetErrorString(e));
  }
  h->sc_mode = scaling;
  h->sc_groups = G;
  *out = h;
  return 0;
}

static int gpedm_create_lagged_impl(int device, int64_t T, const double* y, const int32_t* pop, int32_t np, int32_t E,
                                    int32_t tau, int32_t scaling, const double* rhomat, gpedm_handle* out) {
  if (!out) return set_err(GPEDM_ERR_ARG, "out handle is NULL");
  *out = nullptr;
  int ndev = gpedm_device_count();
  if (ndev == 0) return set_err(GPEDM_ERR_CUDA, "no CUDA device available (libgpedm_b200 has no CPU fallback)");
  if (device < 0 || device >= ndev) return set_err(GPEDM_ERR_ARG, "device %d out of range (have %d)", device, ndev);
  LagSeries S;
  int rc = upload_series(device, T, y, pop, np, S);
  if (rc) return rc;
  return create_from_series(S, E, tau, scaling, rhomat, out);
}

static int gpedm_lagged_info_impl(gpedm_handle h, int64_t* rows, double* center, double* scale, double* Y, double* X,
                                  int32_t* pop) {
  if (!h) return set_err(GPEDM_ERR_ARG, "NULL handle");
  if (!h->dcenter) return set_err(GPEDM_ERR_STATE, "handle was not made by gpedm_create_lagged");
  CU(cudaSetDevice(h->device));
  const size_t nstat = (size_t)(h->d + 1) * h->sc_groups;
  if (rows)
    for (int64_t i = 0; i < h->N; ++i) rows[i] = h->rows[i];
  if (center) CU(cudaMemcpy(center, h->dcenter, nstat * 8, cudaMemcpyDeviceToHost));
  if (scale) CU(cudaMemcpy(scale, h->dscale, nstat * 8, cudaMemcpyDeviceToHost));
  if (Y) memcpy(Y, h->hY.data(), (size_t)h->N * 8);
  if (X) CU(cudaMemcpy(X, h->dX, (size_t)h->N * h->d * 8, cudaMemcpyDeviceToHost));
  if (pop)
    for (int64_t i = 0; i < h->N; ++i) pop[i] = h->hpop[i];
  return 0;
}

static int gpedm_get_fit_stats_impl(gpedm_handle h, gpedm_fit_stats* out) {
  if (!h || !out) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (!h->have_full) return set_err(GPEDM_ERR_STATE, "no full evaluation resident");
  CU(cudaSetDevice(h->device));
  // the handle's scalar scratch (device + pinned host) is free between evaluations
  fit_stats_kernel<<<1, 256, 0, h->stream>>>(h->dY, h->da, h->ddiagS, h->dpop, (int)h->N, h->cur_ve, h->dcenter,
                                             h->dscale, h->d + 1, h->dcenter ? h->sc_mode : 0, h->dscal);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h->hscal, h->dscal, 5 * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  const double* v = h->hscal;
  out->n = v[0];
  out->obs_mean = v[1];
  out->ss_total = v[2];
  out->ss_insample = v[3];
  out->ss_loo = v[4];
  out->R2_insample = 1.0 - v[3] / v[2];
  out->R2_loo = 1.0 - v[4] / v[2];
  out->rmse_insample = sqrt(v[3] / v[0]);
  out->rmse_loo = sqrt(v[4] / v[0]);
  return 0;
}

// Default start of fitGP (reference R/GPfunctions.R:409, :425-429): pars = (0.1 x d, 0.001, 0.999, 0.5),
// transformed with rhomin = 0 (so the rho coordinate starts at 0 for one or many populations).
static void default_initparst(int d, double* t) {
  for (int i = 0; i < d; ++i) t[i] = log(0.1);
  t[d] = logit3(0.001, VEMIN, VEMAX);
  t[d + 1] = logit3(1 - 0.001, SIGMA2MIN, SIGMA2MAX);
  t[d + 2] = logit3(0.5, 0.0, 1.0);
}

static int gpedm_fit_grid_impl(int64_t T, const double* y, const int32_t* pop, int32_t np, int32_t ncells,
                               const int32_t* E, const int32_t* tau, int32_t scaling, double modeprior, int32_t ngpus,
                               const int32_t* devices, const gpedm_rprop_options* opts, gpedm_result* results,
                               gpedm_fit_stats* stats) {
  if (ncells < 0 || (ncells > 0 && (!E || !tau || !results || !y))) return set_err(GPEDM_ERR_ARG, "bad arguments");
  if (ncells == 0) return 0;
  int ndev = gpedm_device_count();
  if (ndev == 0) return set_err(GPEDM_ERR_CUDA, "no CUDA device available (libgpedm_b200 has no CPU fallback)");
  if (ngpus < 1 || ngpus > ndev) return set_err(GPEDM_ERR_ARG, "ngpus=%d out of range (have %d devices)", ngpus, ndev);
  std::vector<int> devs(ngpus);
  for (int g = 0; g < ngpus; ++g) {
    devs[g] = devices ? devices[g] : g;
    if (devs[g] < 0 || devs[g] >= ndev) return set_err(GPEDM_ERR_ARG, "device %d out of range", devs[g]);
  }
  for (int c = 0; c < ncells; ++c)
    if (E[c] < 1 || E[c] > GPEDM_MAX_D || tau[c] < 1) return set_err(GPEDM_ERR_ARG, "cell %d: bad E / tau", c);
  // queue by descending estimated cost: N ~ T - E tau rows, d = E predictors
  std::vector<int> order(ncells);
  for (int i = 0; i < ncells; ++i) order[i] = i;
  auto cost = [&](int c) {
    const double n = std::max<double>(1.0, (double)T - (double)E[c] * tau[c]);
    return n * n * n + 40.0 * E[c] * n * n;
  };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
  const int W = g_batch_workers > 0 ? g_batch_workers : (T <= 4096 ? 3 : 1);
  std::vector<LagSeries> series(ngpus);
  for (int g = 0; g < ngpus; ++g) {  // the raw series goes to every GPU once
    int rc = upload_series(devs[g], T, y, pop, np, series[g]);
    if (rc) return rc;
  }
  std::atomic<int> next(0);
  std::vector<std::vector<BatchRecord>> per_worker(ngpus * W), per_gpu(ngpus);
  std::vector<std::string> werrs(ngpus * W);
  auto worker = [&](int g, int w) {
    const int wi = g * W + w;
    cudaSetDevice(devs[g]);
    for (;;) {
      const int q = next.fetch_add(1);
      if (q >= ncells) break;
      const int c = order[q];
      BatchRecord rec;
      memset(&rec, 0, sizeof(rec));
      rec.fit_index = c;
      rec.valid = 1;
      gpedm_handle h = nullptr;
      int rc = create_from_series(series[g], E[c], tau[c], scaling, nullptr, &h);
      if (rc == 0) {
        double ip[GPEDM_MAX_P];
        default_initparst(E[c], ip);
        rc = gpedm_fit_rprop(h, ip, modeprior, nullptr, NAN, opts, &rec.res);
        if (rc == 0) rc = gpedm_get_fit_stats(h, &rec.st);
      }
      rec.res.status = rc;
      if (rc != 0 && werrs[wi].empty()) werrs[wi] = gpedm_last_error();
      if (h) gpedm_destroy(h);
      per_worker[wi].push_back(rec);
    }
  };
  {
    std::vector<std::thread> th;
    try {
      for (int g = 0; g < ngpus; ++g)
        for (int w = 0; w < W; ++w) th.emplace_back(worker, g, w);
    } catch (...) {
      next.store(ncells);
      for (auto& t : th) t.join();
      return set_err(GPEDM_ERR_STATE, "could not start grid worker threads");
    }
    for (auto& t : th) t.join();
  }
  std::string err;
  for (int wi = 0; wi < ngpus * W; ++wi) {
    per_gpu[wi / W].insert(per_gpu[wi / W].end(), per_worker[wi].begin(), per_worker[wi].end());
    if (err.empty() && !werrs[wi].empty()) err = werrs[wi];
  }
  std::vector<gpedm_fit_stats> st_tmp(ncells);
  int rc = exchange_records(devs, per_gpu, ncells, results, stats ? stats : st_tmp.data());
  if (rc) return rc;
  if (!err.empty()) {
    int first_bad = 0;
    for (int i = 0; i < ncells && !first_bad; ++i) first_bad = results[i].status;
    return set_err(first_bad ? first_bad : GPEDM_ERR_STATE, "at least one grid cell failed: %s", err.c_str());
  }
  return 0;
}

// ------------------------------------------------------------------------------- measurement
extern "C" int gpedm_fp64_peak(int device, int which, double* tflops) {
  if (!tflops) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (gpedm_device_count() == 0) return set_err(GPEDM_ERR_CUDA, "no CUDA device available");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int nsm = prop.multiProcessorCount;
  double* dout = nullptr;
  CU(cudaMalloc((void**)&dout, (size_t)nsm * 4 * 256 * 8));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  float best = 1e30f;
  const int iters = which == 0 ? 4000 : 20000;
  const int ctas = which == 0 ? nsm : nsm * 4;
  for (int rep = 0; rep < 4; ++rep) {
    CU(cudaEventRecord(e0));
    if (which == 0)
      probe_dmma_kernel<<<ctas, 256>>>(dout, iters);
    else
      probe_dfma_kernel<<<ctas, 256>>>(dout, iters);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CU(cudaGetLastError());
  double fl = which == 0 ? (double)ctas * 8 * 32.0 * iters * 512.0 : (double)ctas * 256 * 16.0 * iters * 2.0;
  *tflops = fl / best * 1e-9;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dout);
  return 0;
}

extern "C" int gpedm_bench_evals(gpedm_handle h, const double* parst, double modeprior, int steps, float* ms_per_eval) {
  if (!h || !parst || !ms_per_eval || steps < 1) return set_err(GPEDM_ERR_ARG, "bad arguments");
  CU(cudaSetDevice(h->device));
  HostPars hp;
  host_transform(h, parst, modeprior, nullptr, NAN, hp);
  fill_params(h, hp);
  h->have_full = false;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, h->stream));
  for (int s = 0; s < steps; ++s) {
    int rc = enqueue_eval(h, false);
    if (rc) return rc;
  }
  CU(cudaEventRecord(e1, h->stream));
  CU(cudaEventSynchronize(e1));
  CU(cudaGetLastError());
  float ms;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  *ms_per_eval = ms / steps;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  collect_phase_times(h);
  if (*h->hinfo > 0) return set_err(*h->hinfo, "covariance matrix not positive definite at row %d", *h->hinfo);
  return 0;
}


// ------------------------------------------------------------------------------- guarded entry points
// The implementations above use std::vector / std::thread; no C++ exception may cross the C boundary.
#define GPEDM_GUARDED(call)                                                              \
  try {                                                                                  \
    return (call);                                                                       \
  } catch (const std::bad_alloc&) {                                                      \
    return set_err(GPEDM_ERR_STATE, "out of host memory");                               \
  } catch (const std::exception& ex) {                                                   \
    return set_err(GPEDM_ERR_STATE, "internal error: %s", ex.what());                    \
  } catch (...) {                                                                        \
    return set_err(GPEDM_ERR_STATE, "internal error");                                   \
  }
extern "C" int gpedm_create(int device, int64_t N, int32_t d, int32_t np, const double* X, const double* Y, const int32_t* pop, const double* rhomat, gpedm_handle* out) {
  GPEDM_GUARDED(gpedm_create_impl(device, N, d, np, X, Y, pop, rhomat, out));
}

extern "C" int gpedm_set_data(gpedm_handle h, const double* X, const double* Y, const int32_t* pop) {
  GPEDM_GUARDED(gpedm_set_data_impl(h, X, Y, pop));
}

extern "C" int gpedm_get_vector(gpedm_handle h, int which, double* outv) {
  GPEDM_GUARDED(gpedm_get_vector_impl(h, which, outv));
}

extern "C" int gpedm_get_matrix(gpedm_handle h, int which, double* outm) {
  GPEDM_GUARDED(gpedm_get_matrix_impl(h, which, outm));
}

extern "C" int gpedm_getcov(int device, int32_t d, const double* phi, double sigma2, double rho, int64_t N, const double* X, const int32_t* pop, int64_t M, const double* X2, const int32_t* pop2, int32_t np, const double* rhomat, double* Cd_out) {
  GPEDM_GUARDED(gpedm_getcov_impl(device, d, phi, sigma2, rho, N, X, pop, M, X2, pop2, np, rhomat, Cd_out));
}

extern "C" int gpedm_covinv(int device, int64_t n, const double* Sigma, double* L_out, double* iKVs_out) {
  GPEDM_GUARDED(gpedm_covinv_impl(device, n, Sigma, L_out, iKVs_out));
}

extern "C" int gpedm_predict_new(gpedm_handle h, int64_t M, const double* Xnew, const int32_t* popnew, double* mean, double* var, double* gpgrad) {
  GPEDM_GUARDED(gpedm_predict_new_impl(h, M, Xnew, popnew, mean, var, gpgrad));
}

extern "C" int gpedm_predict_loo(gpedm_handle h, int32_t exclradius, const double* time, const int32_t* primary, double* mean, double* var, double* gpgrad) {
  GPEDM_GUARDED(gpedm_predict_loo_impl(h, exclradius, time, primary, mean, var, gpgrad));
}

extern "C" int gpedm_predict_lto(gpedm_handle h, int32_t exclradius, const double* time, const int32_t* primary, double* mean, double* var, double* gpgrad) {
  GPEDM_GUARDED(gpedm_predict_lto_impl(h, exclradius, time, primary, mean, var, gpgrad));
}

extern "C" int gpedm_predict_sequential(gpedm_handle h, const double* time, double* ymean, double* ysd2) {
  GPEDM_GUARDED(gpedm_predict_sequential_impl(h, time, ymean, ysd2));
}

extern "C" int gpedm_fit_batch(int32_t nfits, const gpedm_fit_desc* descs, int32_t ngpus, const int32_t* devices, const gpedm_rprop_options* opts, gpedm_result* results) {
  GPEDM_GUARDED(gpedm_fit_batch_impl(nfits, descs, ngpus, devices, opts, results));
}

extern "C" int gpedm_create_lagged(int device, int64_t T, const double* y, const int32_t* pop, int32_t np, int32_t E,
                                   int32_t tau, int32_t scaling, const double* rhomat, gpedm_handle* out) {
  GPEDM_GUARDED(gpedm_create_lagged_impl(device, T, y, pop, np, E, tau, scaling, rhomat, out));
}
extern "C" int gpedm_lagged_info(gpedm_handle h, int64_t* rows, double* center, double* scale, double* Y, double* X,
                                 int32_t* pop) {
  GPEDM_GUARDED(gpedm_lagged_info_impl(h, rows, center, scale, Y, X, pop));
}
extern "C" int gpedm_get_fit_stats(gpedm_handle h, gpedm_fit_stats* out) {
  GPEDM_GUARDED(gpedm_get_fit_stats_impl(h, out));
}
extern "C" int gpedm_fit_grid(int64_t T, const double* y, const int32_t* pop, int32_t np, int32_t ncells,
                              const int32_t* E, const int32_t* tau, int32_t scaling, double modeprior, int32_t ngpus,
                              const int32_t* devices, const gpedm_rprop_options* opts, gpedm_result* results,
                              gpedm_fit_stats* stats) {
  GPEDM_GUARDED(gpedm_fit_grid_impl(T, y, pop, np, ncells, E, tau, scaling, modeprior, ngpus, devices, opts, results,
                                    stats));
}

#ifdef GPEDM_TIMELINE
// Debug export: copy the CTA timeline records logged so far (4 x uint64 each) and reset the log.
extern "C" int gpedm_debug_timeline(unsigned long long* out, unsigned int max_records, unsigned int* n_out) {
  unsigned int n = 0;
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(&n, g_tl_count, sizeof(n)));
  if (n > TL_MAX) n = TL_MAX;
  if (n > max_records) n = max_records;
  if (out && n) CU(cudaMemcpyFromSymbol(out, g_tl_rec, (size_t)n * 4 * sizeof(unsigned long long)));
  const unsigned int zero = 0;
  CU(cudaMemcpyToSymbol(g_tl_count, &zero, sizeof(zero)));
  if (n_out) *n_out = n;
  return 0;
}
#endif

#ifdef GPEDM_DIAG_TIMING
extern "C" int gpedm_debug_diag_timing(long long* out) {
  CU(cudaMemcpyFromSymbol(out, g_diag_ts, sizeof(long long) * 32));
  return 0;
}
#endif
