// libgpedm_b200 host orchestration - the fit handle: device buffers, streams, events; creation / destruction / data upload.
// Part of the single translation unit gpedm_core.cu (included in order; not a standalone header).
#pragma once

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
  double *bufA = nullptr, *bufB = nullptr, *bufC = nullptr, *bufD = nullptr;  // bufD: kept copy of Sigma (lower tiles)
  EvalParams* dparams = nullptr;
  void* pinned = nullptr;         // one recycled pinned block holding hparams | hscal | hinfo
  EvalParams* hparams = nullptr;  // pinned
  double *dz = nullptr, *da = nullptr, *ddiagS = nullptr, *dpart = nullptr, *dlogdet = nullptr,
         *dtracepart = nullptr, *dscal = nullptr;
  double* hscal = nullptr;  // pinned
  int* dinfo = nullptr;
  int* hinfo = nullptr;  // pinned
  int4* dtasks = nullptr;  // task list of the persistent factorisation kernel (dag_factor.cuh)
  int ntasks = 0;          // all tasks: Cholesky, inverse, Sigma^-1 = X^T X
  int ntasks_factor = 0;   // without the Sigma^-1 tasks (sequential helper: only L and z are needed); that list starts at
  int factor_off = 0;      // dtasks + factor_off
  int* dsync = nullptr;    // its task counter + per-tile ready flags, zeroed at the start of every evaluation
  DagFit* dfit = nullptr;  // its description of this fit's matrices
  int num_sms = 148;
  int tile_split = 1;  // CTAs per 128 x 128 tile in the assembly / trace kernels (column ranges; > 1 below one wave of tiles)
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
// 32-row strips are used on the chain when (tiles x 4 strips) <= (factor / 2) x SMs; factors in halves
static const int g_split_update2 = getenv("GPEDM_SPLIT_UPDATE2") ? atoi(getenv("GPEDM_SPLIT_UPDATE2")) : 4;
static const int g_split_panel2 = getenv("GPEDM_SPLIT_PANEL2") ? atoi(getenv("GPEDM_SPLIT_PANEL2")) : 4;
static const bool g_no_split = getenv("GPEDM_NO_SPLIT") != nullptr && atoi(getenv("GPEDM_NO_SPLIT")) != 0;
static const int g_graph_max_nb = getenv("GPEDM_GRAPH_MAX_NB") ? atoi(getenv("GPEDM_GRAPH_MAX_NB")) : 48;
static const int g_batch_workers = getenv("GPEDM_BATCH_WORKERS_PER_GPU") ? atoi(getenv("GPEDM_BATCH_WORKERS_PER_GPU")) : 0;
static const bool g_no_graph = getenv("GPEDM_NO_GRAPH") != nullptr && atoi(getenv("GPEDM_NO_GRAPH")) != 0;
static const bool g_diag_ref = getenv("GPEDM_DIAG_REF") != nullptr && atoi(getenv("GPEDM_DIAG_REF")) != 0;
// GPEDM_NO_DAG=1: round-1 stream-ordered Cholesky / recursive inverse instead of the persistent task kernel
static const bool g_no_dag = getenv("GPEDM_NO_DAG") != nullptr && atoi(getenv("GPEDM_NO_DAG")) != 0;
// rows of the inverse are queued this many block columns behind the Cholesky
static const int g_dag_lead = getenv("GPEDM_DAG_LEAD") ? std::max(0, atoi(getenv("GPEDM_DAG_LEAD"))) : 1;
// k tiles per piece of a long tile (512 = never split)
static const int g_dag_chunk = getenv("GPEDM_DAG_CHUNK") ? std::min(512, std::max(1, atoi(getenv("GPEDM_DAG_CHUNK")))) : 512;
// chain tile L(c+1, c) handed to DIAG(c+1) in 32-column strips (dag_factor.cuh); GPEDM_DAG_NO_STRIPS=1: whole tile + flag
static const bool g_dag_strips = !(getenv("GPEDM_DAG_NO_STRIPS") != nullptr && atoi(getenv("GPEDM_DAG_NO_STRIPS")) != 0);
// Tiles L(c+1 .. c+t, c) are SOLVED against the block columns of L_cc as DIAG(c) publishes them (dag_trsm_chain) instead of
// multiplied by D_c once the inverse is complete, and handed on strip by strip.  GPEDM_DAG_TRSM=t; 0: product with D_c
// everywhere.  Measured (profiles/r02_trsm_depth.jsonl): the gain saturates at t = 6 .. 8 (N = 2038: 1.074 ms with 0,
// 0.893 with 2, 0.826 with 4, 0.813 with 6 and beyond); N >= 8182 is indifferent to t.
static const int g_dag_trsm = !g_dag_strips ? 0 : (getenv("GPEDM_DAG_TRSM") ? std::max(0, atoi(getenv("GPEDM_DAG_TRSM"))) : 6);
static const int g_dag_lag = getenv("GPEDM_DAG_LAG") ? std::max(1, atoi(getenv("GPEDM_DAG_LAG"))) : 2;
// Sigma^-1 = X^T X tiles summed in pieces of this many block rows, interleaved with the factorisation (0 = whole
// tiles after it, the default: interleaved pieces were measured SLOWER at every size - 4086 rows 3.38 -> 3.66 ms
// with pieces of 8, 4.07 with 4 - because the busier CTAs claim the chain tasks later, profiles/r02_lauum_pieces.txt).
static const int g_dag_lchunk = getenv("GPEDM_DAG_LAUUM_CHUNK") ? std::max(0, atoi(getenv("GPEDM_DAG_LAUUM_CHUNK"))) : 0;
static const int g_dag_lgroup = getenv("GPEDM_DAG_LAUUM_GROUP") ? std::max(1, atoi(getenv("GPEDM_DAG_LAUUM_GROUP"))) : 1024;
static const int g_dag_lmaxnb = getenv("GPEDM_DAG_LAUUM_MAX_NB") ? atoi(getenv("GPEDM_DAG_LAUUM_MAX_NB")) : 40;
static const int g_dag_ltail = getenv("GPEDM_DAG_LAUUM_TAIL") ? std::max(1, atoi(getenv("GPEDM_DAG_LAUUM_TAIL"))) : 1;
static inline int dag_lauum_chunk(int nb) { return (nb >= 3 && nb <= g_dag_lmaxnb) ? g_dag_lchunk : 0; }
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
    // memory cached in the stream-ordered pool instead of returning 2.1 GB to the driver each time.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
  }
  CU(cudaFuncSetAttribute(diag_potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES));
  CU(cudaFuncSetAttribute(diag_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM_BYTES));
  CU(cudaFuncSetAttribute(dag_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DAG_SMEM_BYTES));
  CU(cudaFuncSetAttribute(small_eval_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_eval_smem_bytes<8>()));
  CU(cudaFuncSetAttribute(small_eval_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_eval_smem_bytes<16>()));
  CU(cudaFuncSetAttribute(small_eval_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_eval_smem_bytes<32>()));
  int trace_smem = (2 * MAXD * TB + 2 * TB + MAXD + 8 * (MAXD + 3)) * 8 + 2 * TB * 4;
  CU(cudaFuncSetAttribute(grad_trace_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, trace_smem));
  CU(cudaFuncSetAttribute(grad_trace_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, trace_smem));
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
    void* bufs[] = {h->dX, h->dY, h->drhotab, h->dpop, h->bufA, h->bufB, h->bufC, h->bufD, h->dparams, h->dz, h->da,
                    h->ddiagS, h->dpart, h->dlogdet, h->dtracepart, h->dscal, h->dinfo, h->dcenter, h->dscale,
                    h->dtasks, h->dsync, h->dfit};
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

// ------------------------------------------------------------------------------- factorisation task list
// Task list of the persistent factorisation kernel (dag_factor.cuh).  A tile whose k loop is longer
// than `ch` tiles is computed in PIECES of ch k tiles (running partial result kept in place / in the
// scratch matrix); only the last piece carries the epilogue.  Placement by "column step" s:
//   * final piece of POTRF(i, s), i >= s + 2, and of TRTRI row s - lag: at step s;
//   * the two chain tasks of block column c - DIAG(c) and POTRF(c+1, c) - `lead` steps early: their
//     CTAs are then resident, with everything but the last k tiles accumulated, when the chain arrives;
//   * a non-final piece as soon as its operands are final, staggered over the following ch steps so
//     that the bulk work never queues up in front of the chain.
// Within a step: chain tasks, Cholesky tiles, inverse tiles, then pieces.  Every dependency of a task
// precedes it in the list, except that chain tasks may precede theirs (bounded: 2 (lead + 1) tasks).
static void build_dag_tasks(int nb, int lag, int lead, int ch, int lch, int ltail, int fit, std::vector<int4>& tasks, int* nfactor) {
  struct Ev {
    int step, cls, seq;
    int4 t;
  };
  std::vector<Ev> evs;
  evs.reserve((size_t)nb * nb * 2);
  int seq = 0;
  auto word = [](int kt0, int kt1, int piece, int last) { return kt0 | (kt1 << 10) | (piece << 20) | (last << 30); };
  // tile (type, i, j) with nkt k tiles; koff: TRTRI operand rows start at block row j
  auto add_tile = [&](int type, int i, int j, int nkt, int final_step, int cls, int ready_off, int partial_cap) {
    const int np = std::max(1, (nkt + ch - 1) / ch);
    for (int q = 0; q + 1 < np; ++q) {
      const int e = ready_off + (q + 1) * ch;  // first step at which the operands of piece q are final
      const int st = std::max(e, std::min(e + (i + j) % ch, partial_cap));
      evs.push_back({st, 3, seq++, make_int4(type | (fit << 4), i, j, word(q * ch, (q + 1) * ch, q, 0))});
    }
    evs.push_back({final_step, cls, seq++, make_int4(type | (fit << 4), i, j, word((np - 1) * ch, nkt, np - 1, 1))});
  };
  for (int c = 0; c < nb; ++c) {
    const int fs = std::max(0, c - lead);
    add_tile(DAG_DIAG, c, c, c, fs, 0, 0, std::max(0, fs - 1));
    if (c + 1 < nb) add_tile(DAG_POTRF, c + 1, c, c, fs, 0, 0, std::max(0, fs - 1));
    for (int i = c + 2; i < nb; ++i) add_tile(DAG_POTRF, i, c, c, c, 1, 0, c - 1);
  }
  for (int r = 1; r < nb; ++r)
    for (int j = 0; j < r; ++j) add_tile(DAG_TRTRI, r, j, r - j, r + lag, 2, j + lag, r + lag - 1);
  // Sigma^-1 = X^T X.  S_ij sums over block rows k >= i of X in ascending order - the order in which the inverse
  // completes them.  With lch == 0 every tile is one task after the factorisation (longest k range first): the
  // CTAs that run out of inverse tiles accumulate S behind the inverse instead of idling.  With lch > 0 a tile is
  // summed in pieces - block rows up to the next multiple of lch, ..., and a final piece of the last `ltail`
  // rows - each placed right after the inverse row it ends with: a chain-bound factorisation (nb <~ 40, where
  // most CTAs wait for the Cholesky chain most of the time) then absorbs the X^T X work while it waits, and only
  // the short final pieces remain after the last column.
  int nfac = 0;
  if (lch > 0) {
    const int tail0 = std::max(0, nb - std::max(1, ltail));  // first block row of the final piece
    for (int i = 0; i < nb; ++i)
      for (int j = 0; j <= i; ++j) {
        int k0 = i, piece = 0;
        while (k0 < nb) {
          int k1;
          if (k0 >= tail0)
            k1 = nb;
          else
            k1 = std::min(tail0, (k0 / lch + 1) * lch);
          const int last = k1 == nb;
          evs.push_back({last ? nb + lag + 1 : k1 - 1 + lag + 1, 4, seq++,
                         make_int4(DAG_LAUUM | (fit << 4), i, j, word(k0 - i, k1 - i, piece, last))});
          ++piece;
          k0 = k1;
        }
      }
  }
  std::stable_sort(evs.begin(), evs.end(), [](const Ev& x, const Ev& y) {
    if (x.step != y.step) return x.step < y.step;
    if (x.cls != y.cls) return x.cls < y.cls;
    return x.seq < y.seq;
  });
  tasks.resize(evs.size());
  for (size_t q = 0; q < evs.size(); ++q) {
    tasks[q] = evs[q].t;
    if ((evs[q].t.x & 15) != DAG_LAUUM) nfac = (int)q + 1;
  }
  *nfactor = nfac;
  if (lch <= 0) {
    // Order: longest k ranges first, row by row (GPEDM_DAG_LAUUM_GROUP = G >= nb, the default).  Square groups of
    // G x G tiles share 2 G column panels of X through the L2; measured at N = 8182 with G = 12: HBM reads of the
    // launch 13.3 -> 12.6 GB but 0.2 % SLOWER (the launch is DMMA-bound at 0.8 TB/s of HBM traffic), 2 % slower at
    // N = 4086 - kept as a knob only (profiles/r02_lauum_group.txt).
    const int G = g_dag_lgroup;
    for (int i0 = 0; i0 < nb; i0 += G)
      for (int j0 = 0; j0 <= i0; j0 += G)
        for (int i = i0; i < std::min(nb, i0 + G); ++i)
          for (int j = j0; j <= std::min(i, j0 + G - 1); ++j)
            tasks.push_back(make_int4(DAG_LAUUM | (fit << 4), i, j, word(0, nb - i, 0, 1)));
  }
}

// Allocate a handle for an N x d problem with np populations: device buffers, streams, events.
// The training data (dX, dY, dpop, drhotab, hY, hpop, single_pop) are filled by the caller.
static int alloc_handle(int device, int64_t N, int32_t d, int32_t np, bool has_rhomat, gpedm_handle* out) {
  *out = nullptr;
  if (N < 1 || N > 1023 * TB) return set_err(GPEDM_ERR_ARG, "N=%lld out of range 1..%d", (long long)N, 1023 * TB);
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
  ALLOC(h->bufD, mat);
  ALLOC(h->dparams, sizeof(EvalParams));
  ALLOC(h->dz, Np * 8);
  ALLOC(h->da, Np * 8);
  ALLOC(h->ddiagS, Np * 8);
  ALLOC(h->dpart, (size_t)h->nb * Np * 8);
  ALLOC(h->dlogdet, (size_t)h->nb * 8);
  {
    const int sms = (device < 64 && g_num_sms[device] > 0) ? g_num_sms[device] : 148;
    h->tile_split = h->ntiles * 4 <= 2 * sms ? 4 : (h->ntiles * 2 <= 2 * sms ? 2 : 1);
  }
  ALLOC(h->dtracepart, (size_t)h->ntiles * h->tile_split * (MAXD + 3) * 8);
  ALLOC(h->dscal, NSCAL * 8);
  ALLOC(h->dinfo, 4);
  ALLOC(h->dsync, (size_t)(1 + DAG_NFLAG * h->nb * h->nb) * 4);
  {
    std::vector<int4> tasks;
    const int lch = dag_lauum_chunk(h->nb);
    build_dag_tasks(h->nb, g_dag_lag, g_dag_lead, g_dag_chunk, lch, g_dag_ltail, 0, tasks, &h->ntasks_factor);
    const int nb = h->nb;
    h->ntasks = (int)tasks.size();
    h->factor_off = 0;
    if (lch > 0) {  // Sigma^-1 pieces are interleaved: the factor-only list (sequential helper) is a separate copy
      std::vector<int4> ftasks;
      int nf = 0;
      build_dag_tasks(h->nb, g_dag_lag, g_dag_lead, g_dag_chunk, 0, g_dag_ltail, 0, ftasks, &nf);
      h->factor_off = h->ntasks;
      h->ntasks_factor = nf;
      tasks.insert(tasks.end(), ftasks.begin(), ftasks.begin() + nf);
    }
    ALLOC(h->dtasks, tasks.size() * sizeof(int4));
    ALLOC(h->dfit, sizeof(DagFit));
    DagFit hf;
    hf.A = h->bufA;
    hf.X = h->bufB;
    hf.W = h->bufC;
    hf.ld = Np;
    hf.nb = nb;
    hf.n_valid = (int)N;
    hf.flags = h->dsync + 1;
    hf.logdet_part = h->dlogdet;
    hf.info = h->dinfo;
    cudaError_t e_ = cudaMemcpyAsync(h->dfit, &hf, sizeof(hf), cudaMemcpyHostToDevice, h->stream);
    if (e_ == cudaSuccess)
      e_ = cudaMemcpyAsync(h->dtasks, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream);
    if (e_ == cudaSuccess) e_ = cudaStreamSynchronize(h->stream);  // `tasks` is pageable and goes out of scope
    if (e_ != cudaSuccess) {
      free_handle(h);
      return set_err(GPEDM_ERR_CUDA, "task list upload failed: %s", cudaGetErrorString(e_));
    }
  }
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
  // the X buffer: its diagonal tiles are only ever written on and below the diagonal blocks
  // (diag_block_kernel), the rest of them must read as zero in the tile products
  if (e == cudaSuccess) e = cudaMemsetAsync(h->bufB, 0, mat, h->stream);
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

// A rhomatrix is a between-population correlation matrix: its diagonal must be 1.  The closed forms of the
// leave-out variances (1/S_ii - ve = Cd_ii - c' S^-1 c against the reference's sigma2 - c' S^-1 c, :1080) and the
// diagonal of Cd in the gradient trace terms (sigma2) rely on Cd_ii = sigma2, i.e. on rhomat[p, p] = 1, which is what
// getrhomatrix produces (R/rhomatrix.R:183-186); anything else is refused rather than silently answered differently.
static int check_rhomat(int32_t np, const double* rhomat) {
  if (!rhomat) return 0;
  for (int32_t p = 0; p < np; ++p)
    if (!(fabs(rhomat[(size_t)p * (np + 1)] - 1.0) <= 1e-12))
      return set_err(GPEDM_ERR_ARG, "rhomatrix must have a unit diagonal (entry %d is %.17g)", p, rhomat[(size_t)p * (np + 1)]);
  return 0;
}

static int gpedm_create_impl(int device, int64_t N, int32_t d, int32_t np, const double* X, const double* Y,
                            const int32_t* pop, const double* rhomat, gpedm_handle* out) {
  if (!out) return set_err(GPEDM_ERR_ARG, "out handle is NULL");
  *out = nullptr;
  if (np >= 1 && check_rhomat(np, rhomat)) return GPEDM_ERR_ARG;
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
extern "C" int32_t gpedm_d(gpedm_handle h) { return h ? h->d : -1; }
extern "C" int64_t gpedm_launch_count(gpedm_handle h) { return h ? h->launches : -1; }
extern "C" int gpedm_last_phase_ms(gpedm_handle h, float* ms) {
  if (!h || !ms) return set_err(GPEDM_ERR_ARG, "NULL argument");
  memcpy(ms, h->phase_ms, sizeof(h->phase_ms));
  return 0;
}
