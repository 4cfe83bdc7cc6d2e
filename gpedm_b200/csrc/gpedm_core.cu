// libgpedm_b200: host orchestration + C ABI (include/gpedm.h) for the B200-native GP
// evaluation pipeline.  The scalar parts of getlikegrad (parameter transforms, priors,
// chain rule; reference R/GPfunctions.R:590-630, 719-725, 752-777) and the Rprop loop
// (:522-584) run on the host; everything that touches N x N data runs in the kernels of
// kernels.cuh.  There is no CPU fallback for any N^2 / N^3 step.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler injects itself
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
#include "dag_factor.cuh"
#include "mid_batch.cuh"
#include "small_fit.cuh"
#include "lag_prep.cuh"

using namespace gpedm;

static_assert(MAXD == GPEDM_MAX_D, "MAXD mismatch");

// The host side is one translation unit, split by concern (order matters: static functions):
#include "host_common.cuh"
#include "host_handle.cuh"
#include "host_pipeline.cuh"
#include "host_likelihood.cuh"
#include "host_predict.cuh"
#include "host_batch.cuh"
#include "host_lagged.cuh"

// ------------------------------------------------------------------------------- measurement
extern "C" int gpedm_fp64_peak(int device, int which, double* tflops) {
  if (!tflops) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (gpedm_device_count() == 0) return set_err(GPEDM_ERR_CUDA, "no CUDA device available");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int nsm = prop.multiProcessorCount;
  DevBuf dbuf;
  CU(dbuf.alloc((size_t)nsm * 4 * 256 * 8));
  double* dout = (double*)dbuf.p;
  EventGuard g0, g1;
  CU(g0.create());
  CU(g1.create());
  cudaEvent_t e0 = g0.e, e1 = g1.e;
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
  return 0;
}

extern "C" int gpedm_bench_evals(gpedm_handle h, const double* parst, double modeprior, int steps, float* ms_per_eval) {
  if (!h || !parst || !ms_per_eval || steps < 1) return set_err(GPEDM_ERR_ARG, "bad arguments");
  CU(cudaSetDevice(h->device));
  HostPars hp;
  host_transform(h, parst, modeprior, nullptr, NAN, hp);
  fill_params(h, hp);
  h->have_full = false;
  EventGuard g0, g1;
  CU(g0.create());
  CU(g1.create());
  cudaEvent_t e0 = g0.e, e1 = g1.e;
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

extern "C" int gpedm_last_gather_status(void) { return g_last_gather_status.load(); }

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

// Debug export (host only, no device needed): the task list the persistent factorisation kernel would run for nb
// block columns, with the library's scheduling defaults.  out[4 q .. 4 q + 3] = {type, i, j, kt0 | kt1 << 10 |
// piece << 20 | last << 30}.  tests/test_host_logic.py replays it against the kernel's dependency rules.
extern "C" int gpedm_debug_dag_tasks(int nb, int lauum_chunk, int* out, int max_tasks, int* ntasks, int* nfactor) {
  if (nb < 1 || nb > 1023) return GPEDM_ERR_ARG;
  std::vector<int4> tasks;
  int nf = 0;
  build_dag_tasks(nb, g_dag_lag, g_dag_lead, g_dag_chunk, lauum_chunk < 0 ? dag_lauum_chunk(nb) : lauum_chunk, g_dag_ltail, 0,
                  tasks, &nf);
  if (ntasks) *ntasks = (int)tasks.size();
  if (nfactor) *nfactor = nf;
  if (out)
    for (size_t q = 0; q < tasks.size() && (int)q < max_tasks; ++q) {
      out[4 * q] = tasks[q].x;
      out[4 * q + 1] = tasks[q].y;
      out[4 * q + 2] = tasks[q].z;
      out[4 * q + 3] = tasks[q].w;
    }
  return 0;
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
