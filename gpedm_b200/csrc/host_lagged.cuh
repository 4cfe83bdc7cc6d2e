// libgpedm_b200 host orchestration - on-device data preparation (lagged series), fit statistics and the E x tau grid.
// Part of the single translation unit gpedm_core.cu (included in order; not a standalone header).
#pragma once

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
  if (check_rhomat(S.np, rhomat)) return GPEDM_ERR_ARG;
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

// Grid on a series short enough for the lockstep batch (every cell has at most GPEDM_MID_MAX_N rows): the cells are
// prepared on the device as always (lags, scaling, complete rows - create_from_series on the first GPU), their training
// sets (a few hundred KB each) are handed to the batch scheduler - which advances all cells of a GPU in lockstep through
// ONE persistent factorisation launch per Rprop round (fit_batch_mid / fit_batch_small) - and the fit statistics of
// gpedm_get_fit_stats are reduced from the returned in-sample / leave-one-out means.
static int fit_grid_lockstep(int64_t T, const double* y, const int32_t* pop, int32_t np, int32_t ncells, const int32_t* E,
                             const int32_t* tau, int32_t scaling, double modeprior, int32_t ngpus, const std::vector<int>& devs,
                             const gpedm_rprop_options* opts, gpedm_result* results, gpedm_fit_stats* stats) {
  struct Cell {
    std::vector<double> X, Y, center, scale, loo, ins;
    std::vector<int32_t> pop;
    double ip[GPEDM_MAX_P];
    int groups = 1;
  };
  std::vector<Cell> cells(ncells);
  std::vector<gpedm_fit_desc> descs(ncells);
  {
    LagSeries S;
    int rc = upload_series(devs[0], T, y, pop, np, S);
    if (rc) return rc;
    for (int c = 0; c < ncells; ++c) {
      gpedm_handle h = nullptr;
      rc = create_from_series(S, E[c], tau[c], scaling, nullptr, &h);
      if (rc) return rc;
      Cell& C = cells[c];
      const int64_t N = h->N;
      const size_t nstat = (size_t)(h->d + 1) * h->sc_groups;
      C.groups = h->sc_groups;
      C.Y = h->hY;
      C.pop.assign(h->hpop.begin(), h->hpop.end());
      C.X.resize((size_t)N * h->d);
      C.center.resize(nstat);
      C.scale.resize(nstat);
      C.loo.resize(N);
      C.ins.resize(N);
      cudaError_t e = cudaMemcpy(C.X.data(), h->dX, (size_t)N * h->d * 8, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(C.center.data(), h->dcenter, nstat * 8, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(C.scale.data(), h->dscale, nstat * 8, cudaMemcpyDeviceToHost);
      gpedm_destroy(h);
      if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "grid preparation failed: %s", cudaGetErrorString(e));
      default_initparst(E[c], C.ip);
      gpedm_fit_desc& D = descs[c];
      memset(&D, 0, sizeof(D));
      D.N = N;
      D.d = E[c];
      D.np = np;
      D.X = C.X.data();
      D.Y = C.Y.data();
      D.pop = C.pop.data();
      D.initparst = C.ip;
      D.modeprior = modeprior;
      D.rhofixed = NAN;
      D.want_loo = 1;
      D.loo_mean = C.loo.data();
      D.insamp_mean = C.ins.data();
    }
  }
  const int rc = gpedm_fit_batch_impl(ncells, descs.data(), ngpus, devs.data(), opts, results);
  if (stats) {
    for (int c = 0; c < ncells; ++c) {
      gpedm_fit_stats& st = stats[c];
      memset(&st, 0, sizeof(st));
      if (results[c].status != 0) continue;
      const Cell& C = cells[c];
      const size_t N = C.Y.size();
      const int stride = E[c] + 1;
      auto sc = [&](size_t i, double& ctr, double& s) {  // unscaling of y (first statistic of each group), as fit_stats_kernel
        if (scaling == GPEDM_SCALE_NONE) {
          ctr = 0.0;
          s = 1.0;
        } else {
          const int g = scaling == GPEDM_SCALE_LOCAL ? C.pop[i] : 0;
          ctr = C.center[(size_t)stride * g];
          s = C.scale[(size_t)stride * g];
        }
      };
      double sum = 0.0;
      for (size_t i = 0; i < N; ++i) {
        double ctr, s;
        sc(i, ctr, s);
        sum += C.Y[i] * s + ctr;
      }
      const double mean = sum / (double)N;
      double sst = 0.0, sin = 0.0, sloo = 0.0;
      for (size_t i = 0; i < N; ++i) {
        double ctr, s;
        sc(i, ctr, s);
        const double o = C.Y[i] * s + ctr - mean;
        const double ri = (C.Y[i] - C.ins[i]) * s, rl = (C.Y[i] - C.loo[i]) * s;
        sst += o * o;
        sin += ri * ri;
        sloo += rl * rl;
      }
      st.n = (double)N;
      st.obs_mean = mean;
      st.ss_total = sst;
      st.ss_insample = sin;
      st.ss_loo = sloo;
      st.R2_insample = 1.0 - sin / sst;
      st.R2_loo = 1.0 - sloo / sst;
      st.rmse_insample = sqrt(sin / (double)N);
      st.rmse_loo = sqrt(sloo / (double)N);
    }
  }
  return rc;
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
  if (T <= g_mid_max_n && ncells >= 2 && !g_no_mid_batch)
    return fit_grid_lockstep(T, y, pop, np, ncells, E, tau, scaling, modeprior, ngpus, devs, opts, results, stats);
  // queue by descending estimated cost: N ~ T - E tau rows, d = E predictors
  std::vector<int> order(ncells);
  for (int i = 0; i < ncells; ++i) order[i] = i;
  auto cost = [&](int c) {
    const double n = std::max<double>(1.0, (double)T - (double)E[c] * tau[c]);
    // iterations to convergence grow with the number of length scales (measured 27..58 for E <= 7,
    // up to 105 for E = 8..10 on the Ricker grid): weight the per-evaluation cost accordingly
    return (n * n * n + 40.0 * E[c] * n * n) * (1.0 + 0.15 * E[c]);
  };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
  const int W = default_workers_per_gpu(T, ncells, ngpus);
  std::vector<LagSeries> series(ngpus);
  for (int g = 0; g < ngpus; ++g) {  // the raw series goes to every GPU once
    int rc = upload_series(devs[g], T, y, pop, np, series[g]);
    if (rc) return rc;
  }
  std::atomic<int> next(0);
  std::vector<std::vector<BatchRecord>> per_worker(ngpus * W), per_gpu(ngpus);
  std::vector<std::string> werrs(ngpus * W);
  next.store(std::min((int)ncells, ngpus * W));  // first picks are static, interleaved across GPUs (see gpedm_fit_batch)
  auto worker = [&](int g, int w) {
    const int wi = g * W + w;
    cudaSetDevice(devs[g]);
    bool first = true;
    for (;;) {
      const int q = first ? w * ngpus + g : next.fetch_add(1);
      first = false;
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
