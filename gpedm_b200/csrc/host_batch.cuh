// libgpedm_b200 host orchestration - batched independent fits: NCCL record exchange, lockstep small fits, worker-thread scheduler.
// Part of the single translation unit gpedm_core.cu (included in order; not a standalone header).
#pragma once

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
  // stream-ordered pool allocations (cudaMalloc / cudaFree synchronise the device; see fit_batch_mid)
  StreamGuard stg;  // declared before the buffers: destroyed after them
  CU(stg.create());
  cudaStream_t st = stg.s;
  PoolBuf dX(st), dY(st), dP(st), dR(st), dFits(st), dParams(st), dScal(st), dA(st), dDiag(st), dInfo(st), dActive(st);
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
  if (e == cudaSuccess) e = cudaMemcpyAsync(dX.p, hX.data(), totX * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dY.p, hY.data(), totN * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dP.p, hP.data(), totN * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && totR) e = cudaMemcpyAsync(dR.p, hR.data(), totR * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
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
  CU(cudaMemcpyAsync(dFits.p, hfits.data(), sizeof(SmallFit) * nf, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
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

// ------------------------------------------------------------------------------- batched mid-size fits
// All fits of a batch with 128 < N <= GPEDM_MID_MAX_N advance in lockstep (mid_batch.cuh): per Rprop
// iteration ONE round of launches evaluates every still-active fit - assembly, the persistent
// factorisation kernel over the interleaved task lists of all fits, the two triangular matrix-vector
// products, the trace terms, the scalar reductions - and the scalar Rprop logic of fmingrad_Rprop
// (reference R/GPfunctions.R:522-584) runs on the host per fit exactly as in gpedm_fit_rprop.
// Upper size of the lockstep path.  Measured aggregate evaluations/s of a batch, lockstep vs one fit per worker thread
// (12..48 fits, tools/mid_batch_bench.py): N ~ 1500 3852 vs 1837, 2050 2042 vs 917, 3000 903 vs 560, 4100 403 vs 265,
// 5000 223 vs 173, 6150 127 vs 123, 8190 58 vs 58 - above ~6000 rows one fit fills the GPU and lockstep only costs memory.
static const int g_mid_max_n = getenv("GPEDM_MID_MAX_N") ? atoi(getenv("GPEDM_MID_MAX_N")) : 6144;
// fits per interleaved group of the lockstep task list (0: all active fits, see fit_batch_mid)
static const int g_mid_group = getenv("GPEDM_MID_GROUP") ? atoi(getenv("GPEDM_MID_GROUP")) : 0;
static const bool g_no_mid_batch = getenv("GPEDM_NO_MID_BATCH") != nullptr && atoi(getenv("GPEDM_NO_MID_BATCH")) != 0;

static int fit_batch_mid(int device, const std::vector<int>& idx, const gpedm_fit_desc* descs,
                         const gpedm_rprop_options* opts_in, gpedm_result* results) {
  const int nf = (int)idx.size();
  if (nf == 0) return 0;
  NvtxRange nvtx_entry("gpedm:fit_batch_mid");
  CU(cudaSetDevice(device));
  int rc = ensure_kernel_attrs(device);
  if (rc) return rc;
  gpedm_rprop_options o;
  if (opts_in)
    o = *opts_in;
  else
    gpedm_default_rprop_options(&o);
  const double eta_minus = o.eta_minus - 1, eta_plus = o.eta_plus - 1;
  const int sms = (device < 64 && g_num_sms[device] > 0) ? g_num_sms[device] : 148;
  // ---- layout of one device arena: per fit the three work matrices, vectors, partials, flags
  struct Lay {
    size_t A, X, S, D, dX, dY, dpop, rhotab, z, a, diagS, part, logdet, tracepart, scal, flags, params;
    int N, d, np, nb, Np, ntiles, split, nflags;
  };
  std::vector<Lay> lay(nf);
  size_t tot = 0;
  auto take = [&](size_t bytes) {
    size_t off = tot;
    tot += (bytes + 255) & ~(size_t)255;
    return off;
  };
  int dmax = 1;
  size_t flags_begin = 0, flags_end = 0;
  for (int q = 0; q < nf; ++q) {
    const gpedm_fit_desc& ds = descs[idx[q]];
    if (ds.N < 1 || ds.d < 1 || ds.d > GPEDM_MAX_D || ds.np < 1 || !ds.X || !ds.Y || !ds.pop || !ds.initparst)
      return set_err(GPEDM_ERR_ARG, "bad descriptor for fit %d", idx[q]);
    for (int64_t i = 0; i < ds.N; ++i)
      if (ds.pop[i] < 0 || ds.pop[i] >= ds.np) return set_err(GPEDM_ERR_ARG, "fit %d: pop code outside 0..np-1", idx[q]);
    Lay& L = lay[q];
    L.N = (int)ds.N;
    L.d = ds.d;
    L.np = ds.np;
    L.nb = ceil_div(L.N, TB);
    L.Np = L.nb * TB;
    L.ntiles = L.nb * (L.nb + 1) / 2;
    L.split = 4;  // few tiles per fit: four column ranges per tile in the assembly / trace kernels
    L.nflags = DAG_NFLAG * L.nb * L.nb;
    dmax = std::max(dmax, L.d);
  }
  for (int q = 0; q < nf; ++q) {  // flags of all fits contiguous: one memset per iteration
    if (q == 0) flags_begin = tot;
    lay[q].flags = take((size_t)lay[q].nflags * 4);
    flags_end = tot;
  }
  const size_t counter_off = take(4);
  const size_t info_off = take((size_t)nf * 4);
  size_t x_begin = tot, x_end = tot;
  for (int q = 0; q < nf; ++q) {  // X matrices contiguous: zeroed once (upper part of their diagonal tiles is never written)
    lay[q].X = take((size_t)lay[q].Np * lay[q].Np * 8);
    x_end = tot;
  }
  for (int q = 0; q < nf; ++q) {
    Lay& L = lay[q];
    const size_t mat = (size_t)L.Np * L.Np * 8;
    L.A = take(mat);
    L.S = take(mat);
    L.D = take(mat);
    L.dX = take((size_t)L.N * L.d * 8);
    L.dY = take((size_t)L.Np * 8);
    L.dpop = take((size_t)L.N * 4);
    L.rhotab = descs[idx[q]].rhomat ? take((size_t)L.np * L.np * 8) : 0;
    L.z = take((size_t)L.Np * 8);
    L.a = take((size_t)L.Np * 8);
    L.diagS = take((size_t)L.Np * 8);
    L.part = take((size_t)L.nb * L.Np * 8);
    L.logdet = take((size_t)L.nb * 8);
    L.tracepart = take((size_t)L.ntiles * L.split * (MAXD + 3) * 8);
    L.scal = take((size_t)NSCAL * 8);
    L.params = take(sizeof(EvalParams));
  }
  // device memory from the stream-ordered pool, like the handles: cudaMalloc / cudaFree synchronise the device and were
  // measured at up to hundreds of milliseconds when they follow other work (a 40-cell grid on a 1024-point series took
  // 0.44 .. 0.94 s from call to call with them)
  StreamGuard stg;  // declared before the buffers: destroyed after them
  CU(stg.create());
  cudaStream_t st = stg.s;
  PoolBuf arena(st), dFits(st), dDag(st), dTasks(st), dMapTS(st), dMapT(st), dMapV(st), dActive(st);
  PinnedBuf pParams, pScal, pInfo;
  static const bool test_alloc_fail = getenv("GPEDM_TEST_MID_ALLOC_FAIL") != nullptr;  // exercises the caller's fallback
  cudaError_t e = test_alloc_fail ? cudaErrorMemoryAllocation : arena.alloc(tot);
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "mid-size batch: allocation of %zu bytes failed: %s", tot, cudaGetErrorString(e));
  char* base = (char*)arena.p;
  // ---- per-fit descriptors, static per-fit task lists, uploads
  std::vector<MidFit> hfits(nf);
  std::vector<DagFit> hdag(nf);
  std::vector<std::vector<int4>> ftasks(nf);
  std::vector<gpedm_fit_s> meta(nf);
  e = pParams.alloc(sizeof(EvalParams) * nf);
  if (e == cudaSuccess) e = pScal.alloc((size_t)NSCAL * 8 * nf);
  if (e == cudaSuccess) e = pInfo.alloc((size_t)nf * 4);
  if (e == cudaSuccess) e = cudaMemsetAsync(base + x_begin, 0, x_end - x_begin, st);
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "mid-size batch setup failed: %s", cudaGetErrorString(e));
  EvalParams* hparams = (EvalParams*)pParams.p;
  double* hscal = (double*)pScal.p;
  int* hinfo = (int*)pInfo.p;
  size_t total_tasks = 0, nTS = 0, nT = 0, nV = 0;
  for (int q = 0; q < nf; ++q) {
    const gpedm_fit_desc& ds = descs[idx[q]];
    const Lay& L = lay[q];
    MidFit& F = hfits[q];
    F.A = (double*)(base + L.A);
    F.X = (double*)(base + L.X);
    F.S = (double*)(base + L.S);
    F.D = (double*)(base + L.D);
    F.ld = L.Np;
    F.N = L.N;
    F.nb = L.nb;
    F.ntiles = L.ntiles;
    F.split = L.split;
    F.d = L.d;
    F.pad = 0;
    F.dX = (const double*)(base + L.dX);
    F.dpop = (const int*)(base + L.dpop);
    F.drhotab = ds.rhomat ? (const double*)(base + L.rhotab) : nullptr;
    F.P = (const EvalParams*)(base + L.params);
    F.dY = (const double*)(base + L.dY);
    F.dz = (double*)(base + L.z);
    F.da = (double*)(base + L.a);
    F.ddiagS = (double*)(base + L.diagS);
    F.dpart = (double*)(base + L.part);
    F.dlogdet = (double*)(base + L.logdet);
    F.dtracepart = (double*)(base + L.tracepart);
    F.dscal = (double*)(base + L.scal);
    DagFit& G = hdag[q];
    G.A = F.A;
    G.X = F.X;
    G.W = F.S;
    G.ld = L.Np;
    G.nb = L.nb;
    G.n_valid = L.N;
    G.flags = (int*)(base + L.flags);
    G.logdet_part = F.dlogdet;
    G.info = (int*)(base + info_off) + q;
    int nfac = 0;
    build_dag_tasks(L.nb, 1, 0, 512, 0, 1, q, ftasks[q], &nfac);  // strict dependency order: any number of CTAs makes progress
    total_tasks += ftasks[q].size();
    nTS += (size_t)L.ntiles * L.split;
    nT += L.ntiles;
    nV += ceil_div(L.Np, 256);
    e = cudaMemcpyAsync((void*)F.dX, ds.X, (size_t)L.N * L.d * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync((void*)F.dY, 0, (size_t)L.Np * 8, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync((void*)F.dY, ds.Y, (size_t)L.N * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync((void*)F.dpop, ds.pop, (size_t)L.N * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && ds.rhomat)
      e = cudaMemcpyAsync((void*)F.drhotab, ds.rhomat, (size_t)L.np * L.np * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "mid-size batch upload failed: %s", cudaGetErrorString(e));
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
  e = dFits.alloc(sizeof(MidFit) * nf);
  if (e == cudaSuccess) e = dDag.alloc(sizeof(DagFit) * nf);
  if (e == cudaSuccess) e = dTasks.alloc(total_tasks * sizeof(int4));
  if (e == cudaSuccess) e = dMapTS.alloc(nTS * sizeof(int2));
  if (e == cudaSuccess) e = dMapT.alloc(nT * sizeof(int2));
  if (e == cudaSuccess) e = dMapV.alloc(nV * sizeof(int2));
  if (e == cudaSuccess) e = dActive.alloc((size_t)nf * 4);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dFits.p, hfits.data(), sizeof(MidFit) * nf, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dDag.p, hdag.data(), sizeof(DagFit) * nf, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "mid-size batch setup failed: %s", cudaGetErrorString(e));
  const int DC = dmax <= 8 ? 8 : (dmax <= 12 ? 12 : (dmax <= 16 ? 16 : 32));
  const int asm_smem = 2 * dmax * TB * 8 + 2 * TB * 4;
  const int trace_smem = (2 * dmax * TB + 2 * TB + DC + 8 * (DC + 3)) * 8 + 2 * TB * 4;
  static std::mutex attr_mu;
  {
    std::lock_guard<std::mutex> lock(attr_mu);
    const int big = (2 * MAXD * TB + 2 * TB + MAXD + 8 * (MAXD + 3)) * 8 + 2 * TB * 4;
    cudaFuncSetAttribute(mid_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(mid_trace_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(mid_trace_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(mid_trace_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(mid_trace_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  }
  std::vector<int> active;
  std::vector<int4> tasks;
  std::vector<int2> mapTS, mapT, mapV;
  // one round of launches: evaluate every fit in `active` at the parameters in hparams
  static const bool trace_iters = getenv("GPEDM_MID_TRACE") != nullptr;
  int round_no = 0;
  auto launch = [&](int full) -> int {
    const int na = (int)active.size();
    if (na == 0) return 0;
    const auto t_begin = std::chrono::steady_clock::now();
    // interleave the active fits' task lists position by position; per-CTA maps of the tiled kernels
    tasks.clear();
    mapTS.clear();
    mapT.clear();
    mapV.clear();
    // Task lists are interleaved position by position, so that a CTA which would wait for one fit's Cholesky chain
    // holds another fit's tile instead.  (GPEDM_MID_GROUP = G interleaves only G fits at a time, group after group;
    // measured on 12..24 fits of 3000..6150 rows, G = 1 / 2 / 3 / all: 493 / 714 / 862 / 903 evaluations/s at N ~ 3000,
    // 115 / 125 / 126 / 127 at N ~ 6150 - more fits per group never hurt, so the default is all of them.)
    for (size_t b = 0; b < active.size();) {
      const size_t e3 = g_mid_group > 0 ? std::min(active.size(), b + (size_t)g_mid_group) : active.size();
      size_t maxlen = 0;
      for (size_t a2 = b; a2 < e3; ++a2) maxlen = std::max(maxlen, ftasks[active[a2]].size());
      for (size_t pos = 0; pos < maxlen; ++pos)
        for (size_t a2 = b; a2 < e3; ++a2)
          if (pos < ftasks[active[a2]].size()) tasks.push_back(ftasks[active[a2]][pos]);
      b = e3;
    }
    for (int q : active) {
      const Lay& L = lay[q];
      for (int b = 0; b < L.ntiles * L.split; ++b) mapTS.push_back(make_int2(q, b));
      for (int b = 0; b < L.ntiles; ++b) mapT.push_back(make_int2(q, b));
      for (int b = 0; b < ceil_div(L.Np, 256); ++b) mapV.push_back(make_int2(q, b));
    }
    cudaError_t e2 = cudaSuccess;
    for (int q : active)
      if (e2 == cudaSuccess)
        e2 = cudaMemcpyAsync(base + lay[q].params, hparams + q, sizeof(EvalParams), cudaMemcpyHostToDevice, st);
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(dTasks.p, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice, st);
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(dMapTS.p, mapTS.data(), mapTS.size() * sizeof(int2), cudaMemcpyHostToDevice, st);
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(dMapT.p, mapT.data(), mapT.size() * sizeof(int2), cudaMemcpyHostToDevice, st);
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(dMapV.p, mapV.data(), mapV.size() * sizeof(int2), cudaMemcpyHostToDevice, st);
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(dActive.p, active.data(), (size_t)na * 4, cudaMemcpyHostToDevice, st);
    if (e2 == cudaSuccess) e2 = cudaMemsetAsync(base + flags_begin, 0, flags_end - flags_begin, st);
    if (e2 == cudaSuccess) e2 = cudaMemsetAsync(base + counter_off, 0, 4, st);
    if (e2 == cudaSuccess) e2 = cudaMemsetAsync(base + info_off, 0, (size_t)nf * 4, st);
    if (e2 != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "mid-size batch upload failed: %s", cudaGetErrorString(e2));
    const MidFit* F = (const MidFit*)dFits.p;
    mid_assemble_kernel<<<(unsigned)mapTS.size(), 256, asm_smem, st>>>(F, (const int2*)dMapTS.p);
    DagArgs da;
    da.fits = (const DagFit*)dDag.p;
    da.tasks = (const int4*)dTasks.p;
    da.ntasks = (int)tasks.size();
    da.counter = (int*)(base + counter_off);
    da.strips = g_dag_strips ? 1 : 0;
    da.trsm = g_dag_trsm;
    dag_factor_kernel<<<std::min(sms, da.ntasks), DAG_THREADS, DAG_SMEM_BYTES, st>>>(da);
    mid_trmv_n_kernel<<<(unsigned)mapT.size(), 256, 0, st>>>(F, (const int2*)dMapT.p);
    mid_reduce_kernel<<<(unsigned)mapV.size(), 256, 0, st>>>(F, (const int2*)dMapV.p, 0);
    mid_trmv_t_kernel<<<(unsigned)mapT.size(), 256, 0, st>>>(F, (const int2*)dMapT.p);
    mid_reduce_kernel<<<(unsigned)mapV.size(), 256, 0, st>>>(F, (const int2*)dMapV.p, 1);
    if (DC == 8)
      mid_trace_kernel<8><<<(unsigned)mapTS.size(), 256, trace_smem, st>>>(F, (const int2*)dMapTS.p);
    else if (DC == 12)
      mid_trace_kernel<12><<<(unsigned)mapTS.size(), 256, trace_smem, st>>>(F, (const int2*)dMapTS.p);
    else if (DC == 16)
      mid_trace_kernel<16><<<(unsigned)mapTS.size(), 256, trace_smem, st>>>(F, (const int2*)dMapTS.p);
    else
      mid_trace_kernel<32><<<(unsigned)mapTS.size(), 256, trace_smem, st>>>(F, (const int2*)dMapTS.p);
    if (full) mid_extract_diag_kernel<<<(unsigned)mapV.size(), 256, 0, st>>>(F, (const int2*)dMapV.p);
    mid_final_reduce_kernel<<<na, 256, 0, st>>>(F, (const int*)dActive.p, full);
    for (int q : active)
      if (e2 == cudaSuccess)
        e2 = cudaMemcpyAsync(hscal + (size_t)q * NSCAL, base + lay[q].scal, (size_t)NSCAL * 8, cudaMemcpyDeviceToHost, st);
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(hinfo, base + info_off, (size_t)nf * 4, cudaMemcpyDeviceToHost, st);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
    if (e2 == cudaSuccess) e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "mid-size batch evaluation failed: %s", cudaGetErrorString(e2));
    if (trace_iters)
      fprintf(stderr, "[gpedm mid] round %d: %d active fits, %zu tasks, %.3f ms\n", round_no,
              na, tasks.size(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    ++round_no;
    return 0;
  };
  struct State {
    double pa[GPEDM_MAX_P], del[GPEDM_MAX_P], grad[GPEDM_MAX_P], panew[GPEDM_MAX_P];
    double nll = 0, s = 0, df = 10;
    int iter = 0, nevals = 0, status = 0;
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
      }
    }
    first = false;
    active.swap(next_active);
    if (active.empty()) break;
    rc = launch(0);
  }
  // ---- final full evaluation at each fit's last parameters (:580)
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
  }
  if (rc) return rc;
  int first_bad = 0;
  std::vector<double> ha, hd;
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
    if ((ds.want_loo && (ds.loo_mean || ds.loo_var)) || ds.insamp_mean || ds.insamp_var) {
      ha.resize(ds.N);
      hd.resize(ds.N);
      e = cudaMemcpy(ha.data(), base + lay[q].a, (size_t)ds.N * 8, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(hd.data(), base + lay[q].diagS, (size_t)ds.N * 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "mid-size batch download failed: %s", cudaGetErrorString(e));
      const double ve = hp[q].ve;
      for (int64_t i = 0; i < ds.N; ++i) {
        const double a = ha[i], sii = hd[i], y = ds.Y[i];
        if (ds.want_loo && ds.loo_mean) ds.loo_mean[i] = y - a / sii;
        if (ds.want_loo && ds.loo_var) ds.loo_var[i] = 1.0 / sii - ve;
        if (ds.insamp_mean) ds.insamp_mean[i] = y - ve * a;
        if (ds.insamp_var) ds.insamp_var[i] = ve - ve * ve * sii;
      }
    }
  }
  if (first_bad) return set_err(first_bad, "at least one mid-size fit failed: covariance matrix not positive definite (pivot %d)", first_bad);
  return 0;
}

// Concurrent fits per GPU for the fits that go through one handle each (everything above the lockstep sizes, and
// single fits).  The persistent factorisation kernel occupies every SM, so two resident fits no longer fill each
// other's chain-bound stretches - their factorisations simply alternate (N ~ 8190: 56.4 evaluations/s one at a time,
// 56.8 with three resident) - and every extra resident fit stretches each fit's own duration, which is what bounds
// a grid spread over several GPUs (two resident fits run at half speed each: twice the imbalance at the end of the
// queue).  Two workers hide handle creation / teardown and the host side of the Rprop loop when a GPU has a long
// queue of fits (16 or more); one otherwise.
static int default_workers_per_gpu(int64_t nmax, int nfits, int ngpus) {
  if (g_batch_workers > 0) return g_batch_workers;
  int w;
  if (nmax > 6144) {
    const double per_gpu = (double)nfits / std::max(1, ngpus);
    w = per_gpu >= 16 ? 2 : 1;
  } else {
    // Smaller fits that still come through one handle each (gpedm_fit_grid cells, a single mid-size fit next to larger
    // ones): the evaluation is a dozen launches around a chain-bound factorisation and several of them interleave
    // (measured aggregate evaluations/s on one B200, tools/mid_batch_bench.py with GPEDM_NO_MID_BATCH=1: N ~ 325 2.3 k
    // with one worker, 16.5 k with 12; N ~ 1000 4.5 k with 8; N ~ 2050 0.9 k with 6; N ~ 4100 265 with 5).
    w = nmax <= 640 ? 12 : nmax <= 1536 ? 8 : nmax <= 3072 ? 6 : 5;
  }
  return std::max(1, std::min(w, nfits));
}

// Fixed-width record exchanged by the final all-gather.
struct BatchRecord {
  int32_t fit_index;
  int32_t valid;
  gpedm_result res;
  gpedm_fit_stats st;  // filled by gpedm_fit_grid only
};

// Process-wide NCCL state of the in-process multi-GPU batch: the library is dlopen'ed once and the
// communicators of a device list are created once and reused by every later batch / grid call
// (ncclCommInitAll costs ~0.1-1 s; round 1 paid it per call).
struct NcclState {
  std::mutex mu;
  NcclApi api;
  int loaded = 0;  // 0 not tried, 1 ok, -1 unavailable
  std::vector<int> devs;
  std::vector<void*> comms;
};
static NcclState g_nccl;
// 0: all-gather done and verified against the host records; 1: not needed (one GPU);
// <0: NCCL unavailable or failed - the records were delivered from host memory (see gpedm_last_gather_status)
static std::atomic<int> g_last_gather_status(1);

// NCCL all-gather of the fixed-width records across the GPUs of the batch (single process, one
// communicator per GPU; every GPU contributes `slab` records padded with valid=0), read back from
// device 0 into `all`.  Returns 0 or a negative status with the error string set.
static int nccl_gather_records(const std::vector<int>& devs, const std::vector<std::vector<BatchRecord>>& per_gpu,
                               size_t slab, std::vector<BatchRecord>& all) {
  const int ngpus = (int)devs.size();
  std::lock_guard<std::mutex> lock(g_nccl.mu);
  if (g_nccl.loaded == 0) g_nccl.loaded = load_nccl(g_nccl.api) ? 1 : -1;
  if (g_nccl.loaded < 0) return set_err(GPEDM_ERR_NCCL, "libnccl.so.2 could not be loaded");
  NcclApi& nc = g_nccl.api;
  if (g_nccl.devs != devs) {
    for (void* c : g_nccl.comms)
      if (c) nc.CommDestroy(c);
    g_nccl.comms.assign(ngpus, nullptr);
    g_nccl.devs.clear();
    int nrc = nc.CommInitAll(g_nccl.comms.data(), ngpus, devs.data());
    if (nrc != 0) {
      g_nccl.comms.clear();
      return set_err(GPEDM_ERR_NCCL, "ncclCommInitAll failed: %s", nc.GetErrorString ? nc.GetErrorString(nrc) : "?");
    }
    g_nccl.devs = devs;
  }
  const size_t slab_bytes = slab * sizeof(BatchRecord);
  std::vector<char*> dsend(ngpus, nullptr), drecv(ngpus, nullptr);
  std::vector<cudaStream_t> sts(ngpus, nullptr);
  cudaError_t e = cudaSuccess;
  int nrc = 0;
  for (int g = 0; g < ngpus && e == cudaSuccess; ++g) {
    e = cudaSetDevice(devs[g]);
    if (e == cudaSuccess) e = cudaStreamCreate(&sts[g]);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dsend[g], slab_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&drecv[g], slab_bytes * ngpus);
    std::vector<BatchRecord> pad(slab);
    memset((void*)pad.data(), 0, slab_bytes);
    for (size_t q = 0; q < per_gpu[g].size(); ++q) pad[q] = per_gpu[g][q];
    if (e == cudaSuccess) e = cudaMemcpy(dsend[g], pad.data(), slab_bytes, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    nc.GroupStart();
    for (int g = 0; g < ngpus; ++g) {
      cudaSetDevice(devs[g]);
      nrc = nc.AllGather(dsend[g], drecv[g], slab_bytes, /*ncclChar*/ 0, g_nccl.comms[g], sts[g]);
      if (nrc != 0) break;
    }
    int nrc2 = nc.GroupEnd();
    if (nrc == 0) nrc = nrc2;
    for (int g = 0; g < ngpus && e == cudaSuccess; ++g) {
      cudaSetDevice(devs[g]);
      e = cudaStreamSynchronize(sts[g]);
    }
  }
  all.resize(slab * ngpus);
  if (e == cudaSuccess && nrc == 0) {
    cudaSetDevice(devs[0]);
    e = cudaMemcpy((void*)all.data(), drecv[0], slab_bytes * ngpus, cudaMemcpyDeviceToHost);
  }
  for (int g = 0; g < ngpus; ++g) {
    cudaSetDevice(devs[g]);
    if (dsend[g]) cudaFree(dsend[g]);
    if (drecv[g]) cudaFree(drecv[g]);
    if (sts[g]) cudaStreamDestroy(sts[g]);
  }
  if (nrc != 0) {  // a failed collective may leave the communicators unusable: rebuild them next time
    for (void* c : g_nccl.comms)
      if (c) nc.CommDestroy(c);
    g_nccl.comms.clear();
    g_nccl.devs.clear();
    return set_err(GPEDM_ERR_NCCL, "ncclAllGather failed: %s", nc.GetErrorString ? nc.GetErrorString(nrc) : "?");
  }
  if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "batch gather failed: %s", cudaGetErrorString(e));
  return 0;
}

// Deliver every fit's fixed-width record to the caller.  The worker threads of all GPUs live in this
// process, so the records are already in host memory: they are copied straight into results/stats
// (a finished fit is never discarded because a collective failed).  With several GPUs the records are
// additionally exchanged by ONE NCCL all-gather (the only collective on the path) and the gathered
// copy is checked bit for bit against the host records; NCCL being unavailable or failing is
// reported through gpedm_last_gather_status(), not as a failure of the batch - unless
// GPEDM_REQUIRE_NCCL=1 is set (tests of the collective path).
static int exchange_records(const std::vector<int>& devs, const std::vector<std::vector<BatchRecord>>& per_gpu,
                            int nfits, gpedm_result* results, gpedm_fit_stats* stats) {
  const int ngpus = (int)devs.size();
  std::vector<char> seen(nfits, 0);
  int got = 0;
  for (int g = 0; g < ngpus; ++g)
    for (auto& r : per_gpu[g]) {
      if (!r.valid || r.fit_index < 0 || r.fit_index >= nfits || seen[r.fit_index])
        return set_err(GPEDM_ERR_STATE, "batch produced an invalid or duplicate record (fit %d)", r.fit_index);
      seen[r.fit_index] = 1;
      results[r.fit_index] = r.res;
      if (stats) stats[r.fit_index] = r.st;
      ++got;
    }
  if (got != nfits) return set_err(GPEDM_ERR_STATE, "batch produced %d of %d records", got, nfits);
  if (ngpus == 1) {
    g_last_gather_status = 1;
    return 0;
  }
  static const bool require = getenv("GPEDM_REQUIRE_NCCL") && atoi(getenv("GPEDM_REQUIRE_NCCL")) != 0;
  static const bool skip = getenv("GPEDM_NO_NCCL") && atoi(getenv("GPEDM_NO_NCCL")) != 0;
  if (skip) {
    g_last_gather_status = GPEDM_ERR_NCCL;
    return 0;
  }
  size_t slab = 1;
  for (int g = 0; g < ngpus; ++g) slab = std::max(slab, per_gpu[g].size());
  std::vector<BatchRecord> all;
  int rc = nccl_gather_records(devs, per_gpu, slab, all);
  if (rc == 0) {
    int ok = 0;
    for (auto& r : all)
      if (r.valid && r.fit_index >= 0 && r.fit_index < nfits &&
          memcmp(&results[r.fit_index], &r.res, sizeof(gpedm_result)) == 0)
        ++ok;
    if (ok != nfits) rc = set_err(GPEDM_ERR_STATE, "NCCL gather returned %d of %d matching records", ok, nfits);
  }
  g_last_gather_status = rc;
  if (rc != 0 && require) return rc;
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
  for (int i = 0; i < nfits; ++i)
    if (descs[i].np >= 1 && check_rhomat(descs[i].np, descs[i].rhomat)) return GPEDM_ERR_ARG;
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
  std::vector<int> small_order, mid_order, big_order;
  for (int q = 0; q < nfits; ++q) {
    const int64_t n = descs[order[q]].N;
    if (n <= TB && !g_no_small_batch)
      small_order.push_back(order[q]);
    else if (n > TB && n <= g_mid_max_n && !g_no_mid_batch)
      mid_order.push_back(order[q]);
    else
      big_order.push_back(order[q]);
  }
  if (small_order.size() < 2) {  // a single small fit gains nothing from the batched kernel
    big_order.insert(big_order.end(), small_order.begin(), small_order.end());
    small_order.clear();
  }
  if (mid_order.size() < 2) {  // likewise for the lockstep mid-size path
    big_order.insert(big_order.end(), mid_order.begin(), mid_order.end());
    mid_order.clear();
  }
  std::stable_sort(big_order.begin(), big_order.end(), [&](int a, int b) { return cost(a) > cost(b); });
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
  const int W = default_workers_per_gpu(nmax, nbig, ngpus);
  // slots W, W+1 of each GPU: the lockstep threads of the small (N <= 128) and mid-size fits
  const int WS = W + 2;
  std::vector<std::vector<BatchRecord>> per_worker(ngpus * WS);
  std::vector<std::string> werrs(ngpus * WS);
  // The queue is ordered by descending cost.  The first pick of every worker is static and interleaved
  // across GPUs (worker w of GPU g takes entry w * ngpus + g) so that the heaviest fits start on
  // different devices; everything after that is pulled dynamically.
  next.store(std::min(nbig, ngpus * W));
  auto worker = [&](int g, int w) {
    const int wi = g * WS + w;
    cudaSetDevice(devs[g]);
    bool first = true;
    for (;;) {
      int q = first ? w * ngpus + g : next.fetch_add(1);
      first = false;
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
  auto lockstep_worker = [&](int g, bool mid) {
    const int wi = g * WS + W + (mid ? 1 : 0);
    const std::vector<int>& pool = mid ? mid_order : small_order;
    std::vector<int> mine;
    for (size_t q = g; q < pool.size(); q += ngpus) mine.push_back(pool[q]);  // strided share of the cost-ordered list
    if (mine.empty()) return;
    std::vector<gpedm_result> tmp(nfits);
    int rc = 0;
    if (!mid) {
      rc = fit_batch_small(devs[g], mine, descs, opts, tmp.data());
    } else {
      // waves of at most ~48 GB of work matrices (4 Np^2 doubles per fit)
      const double cap = 48e9;
      size_t b = 0;
      while (b < mine.size()) {
        double bytes = 0;
        size_t e2 = b;
        while (e2 < mine.size()) {
          const double np_ = (double)ceil_div((int)descs[mine[e2]].N, TB) * TB;
          if (e2 > b && bytes + 32.0 * np_ * np_ > cap) break;
          bytes += 32.0 * np_ * np_;
          ++e2;
        }
        std::vector<int> wave(mine.begin() + b, mine.begin() + e2);
        int rcw = fit_batch_mid(devs[g], wave, descs, opts, tmp.data());
        if (rcw == GPEDM_ERR_CUDA && strstr(gpedm_last_error(), "mid-size batch: allocation of") != nullptr) {
          // no room for the arena of this wave (other work holds the memory): one fit at a time through handles
          cudaGetLastError();
          rcw = 0;
          for (int i : wave) {
            const int r1 = run_one_fit(devs[g], descs[i], opts, &tmp[i]);
            if (r1 != 0 && rcw == 0) rcw = r1;
          }
        }
        if (bytes > 32e9) {  // a very large arena is not left cached in the pool (other allocations may want the memory)
          cudaMemPool_t pool;
          if (cudaSetDevice(devs[g]) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, devs[g]) == cudaSuccess)
            cudaMemPoolTrimTo(pool, (size_t)32 << 30);
          cudaGetLastError();
        }
        if (rcw != 0 && rc == 0) {
          rc = rcw;
          werrs[wi] = gpedm_last_error();
        }
        b = e2;
      }
    }
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
        for (int g = 0; g < ngpus; ++g) th.emplace_back(lockstep_worker, g, false);
      if (!mid_order.empty())
        for (int g = 0; g < ngpus; ++g) th.emplace_back(lockstep_worker, g, true);
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
  for (int wi = 0; wi < ngpus * WS; ++wi) {
    const int g = wi / WS;
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
