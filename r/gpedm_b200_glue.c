/*
 * R .Call glue for libgpedm_b200.  R is absent from the build image, so the glue is compiled and EXECUTED
 * there against a small stand-in for R's C API (tests/r_stub/, tests/test_r_glue.py: every entry below is
 * called through the registered table and compared with the direct C-ABI path).
 * Same registration pattern as the reference's generated glue (src/RcppExports.cpp:49-59):
 * R_registerRoutines + R_useDynamicSymbols(FALSE).  Every entry maps SEXP <-> plain pointers and
 * forwards to the C ABI of include/gpedm.h; the fit handle is an external pointer with finalizer.
 *
 * Build (where R is installed):
 *   R CMD SHLIB r/gpedm_b200_glue.c -Iinclude -Lgpedm_b200 -lgpedm_b200 -o gpedmb200.so
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>
#include "gpedm.h"

static void check_status(int rc) {
  if (rc != 0) Rf_error("libgpedm_b200 status %d: %s", rc, gpedm_last_error());
}

static void handle_finalizer(SEXP ptr) {
  gpedm_handle h = (gpedm_handle)R_ExternalPtrAddr(ptr);
  if (h) {
    gpedm_destroy(h);
    R_ClearExternalPtr(ptr);
  }
}

static gpedm_handle get_handle(SEXP ptr) {
  gpedm_handle h = (gpedm_handle)R_ExternalPtrAddr(ptr);
  if (!h) Rf_error("gpedm handle has been released");
  return h;
}

/* .Call("gpedmb200_create", X, Y, popcodes, np, rhomat_or_NULL, device) */
SEXP gpedmb200_create(SEXP X, SEXP Y, SEXP pop, SEXP np, SEXP rhomat, SEXP device) {
  gpedm_handle h = NULL;
  int n = Rf_nrows(X), d = Rf_ncols(X);
  check_status(gpedm_create(Rf_asInteger(device), n, d, Rf_asInteger(np), REAL(X), REAL(Y), INTEGER(pop),
                            Rf_isNull(rhomat) ? NULL : REAL(rhomat), &h));
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

static SEXP result_to_list(const gpedm_result* r) {
  const char* names[] = {"nllpost", "grad", "pars", "parst", "lp", "SS", "logdet", "lnL_LOO", "df", "iter", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  int p = r->p;
  SEXP g = PROTECT(Rf_allocVector(REALSXP, p)), pa = PROTECT(Rf_allocVector(REALSXP, p)),
       pt = PROTECT(Rf_allocVector(REALSXP, p));
  memcpy(REAL(g), r->grad, p * sizeof(double));
  memcpy(REAL(pa), r->pars, p * sizeof(double));
  memcpy(REAL(pt), r->parst, p * sizeof(double));
  SET_VECTOR_ELT(out, 0, Rf_ScalarReal(r->nllpost));
  SET_VECTOR_ELT(out, 1, g);
  SET_VECTOR_ELT(out, 2, pa);
  SET_VECTOR_ELT(out, 3, pt);
  SET_VECTOR_ELT(out, 4, Rf_ScalarReal(r->lp));
  SET_VECTOR_ELT(out, 5, Rf_ScalarReal(r->SS));
  SET_VECTOR_ELT(out, 6, Rf_ScalarReal(r->logdet));
  SET_VECTOR_ELT(out, 7, Rf_ScalarReal(r->lnL_LOO));
  SET_VECTOR_ELT(out, 8, Rf_ScalarReal(r->df));
  SET_VECTOR_ELT(out, 9, Rf_ScalarInteger(r->iter));
  UNPROTECT(4);
  return out;
}

/* getlikegrad: .Call("gpedmb200_likegrad", h, parst, modeprior, fixedpars(NA=free), rhofixed(NA=free), full) */
SEXP gpedmb200_likegrad(SEXP h, SEXP parst, SEXP modeprior, SEXP fixedpars, SEXP rhofixed, SEXP full) {
  gpedm_result r;
  check_status(gpedm_likegrad(get_handle(h), REAL(parst), Rf_asReal(modeprior), REAL(fixedpars),
                              Rf_asReal(rhofixed), Rf_asLogical(full), &r));
  return result_to_list(&r);
}

/* fmingrad_Rprop: .Call("gpedmb200_fit", h, initparst, modeprior, fixedpars, rhofixed) */
SEXP gpedmb200_fit(SEXP h, SEXP initparst, SEXP modeprior, SEXP fixedpars, SEXP rhofixed) {
  gpedm_result r;
  check_status(gpedm_fit_rprop(get_handle(h), REAL(initparst), Rf_asReal(modeprior), REAL(fixedpars),
                               Rf_asReal(rhofixed), NULL, &r));
  return result_to_list(&r);
}

SEXP gpedmb200_vector(SEXP h, SEXP which) {
  gpedm_handle hh = get_handle(h);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)gpedm_n(hh)));
  check_status(gpedm_get_vector(hh, Rf_asInteger(which), REAL(out)));
  UNPROTECT(1);
  return out;
}

SEXP gpedmb200_matrix(SEXP h, SEXP which) {
  gpedm_handle hh = get_handle(h);
  int n = (int)gpedm_n(hh);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  check_status(gpedm_get_matrix(hh, Rf_asInteger(which), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* predict.GP newdata core: returns list(mean, var, grad) */
SEXP gpedmb200_predict_new(SEXP h, SEXP Xnew, SEXP popnew, SEXP wantgrad) {
  gpedm_handle hh = get_handle(h);
  int m = Rf_nrows(Xnew), d = Rf_ncols(Xnew), wg = Rf_asLogical(wantgrad);
  const char* names[] = {"mean", "var", "grad", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, m)), var = PROTECT(Rf_allocVector(REALSXP, m));
  SEXP grad = PROTECT(wg ? Rf_allocMatrix(REALSXP, m, d) : R_NilValue);
  check_status(gpedm_predict_new(hh, m, REAL(Xnew), INTEGER(popnew), REAL(mean), REAL(var), wg ? REAL(grad) : NULL));
  SET_VECTOR_ELT(out, 0, mean);
  SET_VECTOR_ELT(out, 1, var);
  SET_VECTOR_ELT(out, 2, grad);
  UNPROTECT(4);
  return out;
}

/* method: 0 loo, 1 lto; time may be NULL for loo without augmentation rows */
SEXP gpedmb200_predict_leaveout(SEXP h, SEXP method, SEXP exclradius, SEXP time, SEXP primary, SEXP wantgrad) {
  gpedm_handle hh = get_handle(h);
  int n = (int)gpedm_n(hh), nd = 0, d = 0, wg = Rf_asLogical(wantgrad);
  for (int i = 0; i < n; ++i) nd += INTEGER(primary)[i] != 0;
  const char* names[] = {"mean", "var", "grad", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, nd)), var = PROTECT(Rf_allocVector(REALSXP, nd));
  if (wg) d = gpedm_d(hh);
  SEXP grad = PROTECT(wg ? Rf_allocMatrix(REALSXP, nd, d) : R_NilValue);
  const double* t = Rf_isNull(time) ? NULL : REAL(time);
  if (Rf_asInteger(method) == 0)
    check_status(gpedm_predict_loo(hh, Rf_asInteger(exclradius), t, INTEGER(primary), REAL(mean), REAL(var),
                                   wg ? REAL(grad) : NULL));
  else
    check_status(gpedm_predict_lto(hh, Rf_asInteger(exclradius), t, INTEGER(primary), REAL(mean), REAL(var),
                                   wg ? REAL(grad) : NULL));
  SET_VECTOR_ELT(out, 0, mean);
  SET_VECTOR_ELT(out, 1, var);
  SET_VECTOR_ELT(out, 2, grad);
  UNPROTECT(4);
  return out;
}

SEXP gpedmb200_predict_sequential(SEXP h, SEXP time) {
  gpedm_handle hh = get_handle(h);
  int n = (int)gpedm_n(hh);
  const char* names[] = {"ymean", "ysd2", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SEXP m = PROTECT(Rf_allocVector(REALSXP, n)), s = PROTECT(Rf_allocVector(REALSXP, n));
  check_status(gpedm_predict_sequential(hh, REAL(time), REAL(m), REAL(s)));
  SET_VECTOR_ELT(out, 0, m);
  SET_VECTOR_ELT(out, 1, s);
  UNPROTECT(3);
  return out;
}

/* getcov: Cd (M x N) */
SEXP gpedmb200_getcov(SEXP phi, SEXP sigma2, SEXP rho, SEXP X, SEXP pop, SEXP X2, SEXP pop2, SEXP np, SEXP rhomat,
                      SEXP device) {
  int n = Rf_nrows(X), m = Rf_nrows(X2), d = Rf_ncols(X);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, m, n));
  check_status(gpedm_getcov(Rf_asInteger(device), d, REAL(phi), Rf_asReal(sigma2), Rf_asReal(rho), n, REAL(X),
                            INTEGER(pop), m, REAL(X2), INTEGER(pop2), Rf_asInteger(np),
                            Rf_isNull(rhomat) ? NULL : REAL(rhomat), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* getcovinv: list(L, iKVs) of a symmetric positive definite matrix (reference R/GPfunctions.R:818-830;
 * L is returned lower triangular, t(L) is the reference's upper factor) */
SEXP gpedmb200_covinv(SEXP Sigma, SEXP device) {
  int n = Rf_nrows(Sigma);
  const char* names[] = {"L", "iKVs", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SEXP L = PROTECT(Rf_allocMatrix(REALSXP, n, n)), iK = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  check_status(gpedm_covinv(Rf_asInteger(device), n, REAL(Sigma), REAL(L), REAL(iK)));
  SET_VECTOR_ELT(out, 0, L);
  SET_VECTOR_ELT(out, 1, iK);
  UNPROTECT(3);
  return out;
}

/* fitGP fast path when x is the lag set of y: .Call("gpedmb200_create_lagged", y, popcodes_or_NULL, np, E, tau,
 * scaling (0 none, 1 global, 2 local), rhomat_or_NULL, device) -> handle (lags / scaling / complete rows on the device) */
SEXP gpedmb200_create_lagged(SEXP y, SEXP pop, SEXP np, SEXP E, SEXP tau, SEXP scaling, SEXP rhomat, SEXP device) {
  gpedm_handle h = NULL;
  check_status(gpedm_create_lagged(Rf_asInteger(device), (int64_t)Rf_length(y), REAL(y),
                                   Rf_isNull(pop) ? NULL : INTEGER(pop), Rf_asInteger(np), Rf_asInteger(E),
                                   Rf_asInteger(tau), Rf_asInteger(scaling), Rf_isNull(rhomat) ? NULL : REAL(rhomat),
                                   &h));
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

static SEXP stats_to_vector(const gpedm_fit_stats* st) {
  SEXP v = PROTECT(Rf_allocVector(REALSXP, 4));
  REAL(v)[0] = st->R2_insample;
  REAL(v)[1] = st->R2_loo;
  REAL(v)[2] = st->rmse_insample;
  REAL(v)[3] = st->rmse_loo;
  UNPROTECT(1);
  return v;
}

/* c(R2_insample, R2_loo, rmse_insample, rmse_loo) in the original units of y */
SEXP gpedmb200_fit_stats(SEXP h) {
  gpedm_fit_stats st;
  check_status(gpedm_get_fit_stats(get_handle(h), &st));
  return stats_to_vector(&st);
}

/* E x tau grid (vignettes/choosingE.Rmd:63-86): list of per-cell lists (fit result + stats) */
SEXP gpedmb200_fit_grid(SEXP y, SEXP pop, SEXP np, SEXP E, SEXP tau, SEXP scaling, SEXP modeprior, SEXP ngpus) {
  int nc = Rf_length(E);
  gpedm_result* res = (gpedm_result*)R_alloc(nc, sizeof(gpedm_result));
  gpedm_fit_stats* st = (gpedm_fit_stats*)R_alloc(nc, sizeof(gpedm_fit_stats));
  check_status(gpedm_fit_grid((int64_t)Rf_length(y), REAL(y), Rf_isNull(pop) ? NULL : INTEGER(pop), Rf_asInteger(np), nc,
                              INTEGER(E), INTEGER(tau), Rf_asInteger(scaling), Rf_asReal(modeprior),
                              Rf_asInteger(ngpus), NULL, NULL, res, st));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, nc));
  for (int c = 0; c < nc; ++c) {
    const char* names[] = {"fit", "stats", ""};
    SEXP cell = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(cell, 0, result_to_list(&res[c]));
    SET_VECTOR_ELT(cell, 1, stats_to_vector(&st[c]));
    SET_VECTOR_ELT(out, c, cell);
    UNPROTECT(1);
  }
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef CallEntries[] = {
    {"gpedmb200_create", (DL_FUNC)&gpedmb200_create, 6},
    {"gpedmb200_likegrad", (DL_FUNC)&gpedmb200_likegrad, 6},
    {"gpedmb200_fit", (DL_FUNC)&gpedmb200_fit, 5},
    {"gpedmb200_vector", (DL_FUNC)&gpedmb200_vector, 2},
    {"gpedmb200_matrix", (DL_FUNC)&gpedmb200_matrix, 2},
    {"gpedmb200_predict_new", (DL_FUNC)&gpedmb200_predict_new, 4},
    {"gpedmb200_predict_leaveout", (DL_FUNC)&gpedmb200_predict_leaveout, 6},
    {"gpedmb200_predict_sequential", (DL_FUNC)&gpedmb200_predict_sequential, 2},
    {"gpedmb200_getcov", (DL_FUNC)&gpedmb200_getcov, 10},
    {"gpedmb200_covinv", (DL_FUNC)&gpedmb200_covinv, 2},
    {"gpedmb200_create_lagged", (DL_FUNC)&gpedmb200_create_lagged, 8},
    {"gpedmb200_fit_stats", (DL_FUNC)&gpedmb200_fit_stats, 1},
    {"gpedmb200_fit_grid", (DL_FUNC)&gpedmb200_fit_grid, 8},
    {NULL, NULL, 0}};

void R_init_gpedmb200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
