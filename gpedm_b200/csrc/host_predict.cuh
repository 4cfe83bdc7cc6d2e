// libgpedm_b200 host orchestration - prediction entry points: newdata, leave-out (loo / lto), sequential.
// Part of the single translation unit gpedm_core.cu (included in order; not a standalone header).
#pragma once

// ------------------------------------------------------------------------------- prediction
static int gpedm_predict_new_impl(gpedm_handle h, int64_t M, const double* Xnew, const int32_t* popnew, double* mean,
                                 double* var, double* gpgrad) {
  NvtxRange nvtx_entry("gpedm:predict_new");
  if (!h || !Xnew || !popnew || !mean || !var) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (M < 1) return set_err(GPEDM_ERR_ARG, "M must be >= 1");
  if (!h->have_full) return set_err(GPEDM_ERR_STATE, "no full evaluation resident");
  for (int64_t i = 0; i < M; ++i)
    if (popnew[i] < 0 || popnew[i] > h->np)  // code np = a population not present in the training set
      return set_err(GPEDM_ERR_ARG, "popnew[%lld] outside 0..np", (long long)i);
  CU(cudaSetDevice(h->device));
  const int d = h->d;
  const int64_t CH = 4096;  // new points per chunk
  const int64_t chunk_max = std::min<int64_t>(M, CH);
  const int Mp_max = ceil_div((int)chunk_max, TB) * TB;
  const size_t ldv = h->Np;
  // scratch from the stream-ordered pool (predict_iter / getconditionals call this once per step with a handful of
  // points: cudaMalloc + cudaFree of seven buffers synchronise the device on every call)
  PoolBuf bV(h->stream), bU(h->stream), bXn(h->stream), bpn(h->stream), bmean(h->stream), bvar(h->stream), bgrad(h->stream);
  cudaError_t e = bV.alloc(ldv * Mp_max * 8);
  if (e == cudaSuccess) e = bU.alloc(ldv * Mp_max * 8);
  if (e == cudaSuccess) e = bXn.alloc((size_t)chunk_max * d * 8);
  if (e == cudaSuccess) e = bpn.alloc((size_t)chunk_max * 4);
  if (e == cudaSuccess) e = bmean.alloc((size_t)chunk_max * 8);
  if (e == cudaSuccess) e = bvar.alloc((size_t)chunk_max * 8);
  if (e == cudaSuccess && gpgrad) e = bgrad.alloc((size_t)chunk_max * d * 8);
  double *dV = (double*)bV.p, *dU = (double*)bU.p, *dXn = (double*)bXn.p, *dmean = (double*)bmean.p, *dvar = (double*)bvar.p,
         *dgrad = (double*)bgrad.p;
  int* dpn = (int*)bpn.p;
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

// Codes of the time values in order of first appearance (reference: unique(Time), :1142, :1201),
// in O(N log N) instead of a linear search per row.  Returns the number of distinct values.
static int time_codes(const double* time, int N, std::vector<int>& tcode) {
  std::map<double, int> seen;
  tcode.resize(N);
  for (int i = 0; i < N; ++i) {
    auto it = seen.find(time[i]);
    if (it == seen.end()) it = seen.emplace(time[i], (int)seen.size()).first;
    tcode[i] = it->second;
  }
  return (int)seen.size();
}

static int gpedm_predict_loo_impl(gpedm_handle h, int32_t exclradius, const double* time, const int32_t* primary,
                                 double* mean, double* var, double* gpgrad) {
  NvtxRange nvtx_entry("gpedm:predict_loo");
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
  NvtxRange nvtx_entry("gpedm:predict_lto");
  if (!h || !mean || !var || !time) return set_err(GPEDM_ERR_ARG, "NULL argument (time is required)");
  if (!h->have_full) return set_err(GPEDM_ERR_STATE, "no full evaluation resident");
  if (exclradius < 0) return set_err(GPEDM_ERR_ARG, "exclradius must be >= 0");
  const int N = (int)h->N;
  int nd = 0;
  for (int i = 0; i < N; ++i) nd += primary ? (primary[i] != 0) : 1;
  for (int i = 0; i < nd; ++i)
    if (primary && !primary[i]) return set_err(GPEDM_ERR_ARG, "primary rows must precede augmentation rows");
  // timevals = unique(Time) in order of first appearance (:1142)
  std::vector<int> tcode;
  const int nt = time_codes(time, N, tcode);
  std::vector<std::vector<int>> rows_at(nt);  // rows of each time value, ascending
  for (int j = 0; j < N; ++j) rows_at[tcode[j]].push_back(j);
  std::vector<int> members, goff(1, 0), focal, fgroup;
  for (int i = 0; i < nd; ++i) {
    mean[i] = NAN;
    var[i] = NAN;
  }
  std::vector<int> fpos;
  for (int t = 0; t < nt; ++t) {
    int lo = std::max(0, t - exclradius), hi = std::min(nt - 1, t + exclradius);  // :1155-1156
    const size_t start = members.size();
    for (int c = lo; c <= hi; ++c) members.insert(members.end(), rows_at[c].begin(), rows_at[c].end());
    std::sort(members.begin() + start, members.end());  // complement of `train` in row order, :1158
    goff.push_back((int)members.size());
    for (int j : rows_at[t])
      if (j < nd) {  // :1157 focal rows among primary
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
  NvtxRange nvtx_entry("gpedm:predict_sequential");
  if (!h || !time || !ymean || !ysd2) return set_err(GPEDM_ERR_ARG, "NULL argument");
  if (!h->have_full) return set_err(GPEDM_ERR_STATE, "no full evaluation resident");
  CU(cudaSetDevice(h->device));
  const int N = (int)h->N, d = h->d;
  // time-major stable ordering by order of first appearance of each time value (:1201)
  std::vector<int> tcode;
  time_codes(time, N, tcode);
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
