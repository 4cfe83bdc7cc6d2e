// libgpedm_b200 host orchestration - scalar head / tail of getlikegrad, the Rprop loop, and the exports of a resident evaluation.
// Part of the single translation unit gpedm_core.cu (included in order; not a standalone header).
#pragma once

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
  NvtxRange nvtx_entry("gpedm:fit_rprop");
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
  if (d < 1 || d > GPEDM_MAX_D || N < 1 || M < 1 || np < 1) return set_err(GPEDM_ERR_ARG, "bad sizes");
  for (int64_t i = 0; i < N; ++i)
    if (pop[i] < 0 || pop[i] >= np) return set_err(GPEDM_ERR_ARG, "pop[%lld] outside 0..np-1", (long long)i);
  for (int64_t i = 0; i < M; ++i)  // code np = a population that does not occur in `pop` (:796-811)
    if (pop2[i] < 0 || pop2[i] > np) return set_err(GPEDM_ERR_ARG, "pop2[%lld] outside 0..np", (long long)i);
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
