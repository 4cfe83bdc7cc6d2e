/*
 * gpedm.h - C ABI of libgpedm_b200: the B200-native (sm_100a, FP64) replacement for the
 * numerical core of the GPEDM R package (GP log-posterior + gradient, Rprop fit, prediction).
 *
 * Boundary.  The reference's only first-party native code is three fine-grained Rcpp
 * utilities (multMat / matrix2vector / innerProd, reference src/rcpp_la_utils.cpp:6-19,
 * bound through .Call in src/RcppExports.cpp:15-59 and R/RcppExports.R:4-14).  Calling a GPU
 * per Hadamard product would move N^2 doubles across PCIe per call, so the boundary sits one
 * level up: each entry point below replaces one R-internal function of
 * R/GPfunctions.R that calls those utilities / laGP::distance / Matrix::chol / chol2inv.
 * An R `.Call` shim (r/gpedm_b200_glue.c) or any FFI (ctypes, cgo, JNI)
 * binds these directly: plain pointers and sizes, column-major FP64 like R matrices,
 * int32 population codes 0..np-1, NaN for R's NA, no C++ / CUDA / torch types.
 *
 * Conventions
 *   - every function returns int status: 0 ok; >0 = 1-based index of the first non-positive
 *     Cholesky pivot (matrix not positive definite; the reference raises an R error from
 *     Matrix::chol, R/GPfunctions.R:827); <0 = argument / CUDA error (GPEDM_ERR_*).
 *     gpedm_last_error() returns a thread-local message.  Nothing throws across the boundary.
 *   - the caller owns all host buffers; inputs are copied to the device in gpedm_create;
 *     outputs are written to caller buffers; device state lives in the handle.
 *   - entry points are blocking; one handle must not be used from two threads at once.
 *   - there is no CPU fallback: without a CUDA device every call fails with GPEDM_ERR_CUDA.
 */
#ifndef GPEDM_H
#define GPEDM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPEDM_ABI_VERSION 3
#define GPEDM_MAX_D 32              /* max number of predictors (columns of X) */
#define GPEDM_MAX_P (GPEDM_MAX_D + 3) /* phi_1..d, ve, sigma2, rho */

#define GPEDM_ERR_ARG (-1)
#define GPEDM_ERR_CUDA (-2)
#define GPEDM_ERR_STATE (-3)
#define GPEDM_ERR_NCCL (-4)

typedef struct gpedm_fit_s* gpedm_handle;

/* Result record of one evaluation / one fit.  Mirrors the list returned by getlikegrad
 * (reference R/GPfunctions.R:723-749) plus `iter` from fmingrad_Rprop (:581). */
typedef struct gpedm_result {
  double nllpost;  /* -(like + lp), :721                       */
  double lp;       /* log prior, :773                          */
  double like;     /* -0.5 Y'a - sum log L_ii, :648            */
  double SS;       /* Y' a, :728            (full eval only)   */
  double logdet;   /* sum log L_ii, :729                       */
  double lnL_LOO;  /* :742                  (full eval only)   */
  double df;       /* tr(Cd Sigma^-1), :743 (full eval only)   */
  double gradnorm; /* ||grad||_2 of the last evaluation        */
  int32_t status;  /* same as the function return value        */
  int32_t iter;    /* Rprop iterations, :575                   */
  int32_t nevals;  /* likelihood evaluations performed         */
  int32_t p;       /* d + 3                                    */
  double pars[GPEDM_MAX_P];  /* phi.., ve, sigma2, rho  (:732)          */
  double parst[GPEDM_MAX_P]; /* transformed, after fixing (:607-619)    */
  double grad[GPEDM_MAX_P];  /* gradient of nllpost wrt parst (:720)    */
} gpedm_result;

/* Rprop constants of fmingrad_Rprop (reference R/GPfunctions.R:531-536, :556). */
typedef struct gpedm_rprop_options {
  double delta0;    /* 0.1   */
  double deltamin;  /* 1e-6  */
  double deltamax;  /* 50    */
  double eta_minus; /* 0.5   */
  double eta_plus;  /* 1.2   */
  int32_t maxiter;  /* 200   */
  double tol_grad;  /* 1e-4 : loop while ||grad|| > tol_grad           */
  double tol_df;    /* 1e-7 : loop while |nll_new/nll - 1| > tol_df    */
} gpedm_rprop_options;

/* ---- library ------------------------------------------------------------------------- */
int gpedm_abi_version(void);
const char* gpedm_last_error(void);
int gpedm_device_count(void);
void gpedm_default_rprop_options(gpedm_rprop_options* opts);

/* ---- model state --------------------------------------------------------------------- */
/* Upload the training set of one fit: what fitGP hands to fmingrad_Rprop
 * (reference R/GPfunctions.R:447-454): Y (N), X (N x d, column-major), Pop as codes
 * 0..np-1 in order of first appearance, optional rhomatrix (np x np column-major with a UNIT diagonal - a
 * between-population correlation matrix as getrhomatrix returns it; anything else is GPEDM_ERR_ARG -, ordered
 * like the codes; NULL = single rho).  N >= 1, 1 <= d <= GPEDM_MAX_D. */
int gpedm_create(int device, int64_t N, int32_t d, int32_t np, const double* X, const double* Y,
                 const int32_t* pop, const double* rhomat, gpedm_handle* out);
/* Replace the training set of an existing handle (same N, d, np) without reallocating: the
 * host->device copy a caller pays when it re-enters with fresh host buffers, as every R-level
 * call of getlikegrad does (reference R/GPfunctions.R:545, :560, :580). */
int gpedm_set_data(gpedm_handle h, const double* X, const double* Y, const int32_t* pop);
int gpedm_destroy(gpedm_handle h);
int64_t gpedm_n(gpedm_handle h);
int32_t gpedm_d(gpedm_handle h); /* number of predictors (columns of X); -1 for a NULL handle */

/* ---- getlikegrad (reference R/GPfunctions.R:586-750) --------------------------------- */
/* parst: d+3 transformed parameters.  fixedpars: d+2 values, NaN = free (may be NULL).
 * rhofixed: NaN = free.  full = 0 is `returngradonly=TRUE`; full = 1 also fills SS, lnL_LOO,
 * df and leaves mean / variance / Sigma^-1 / L resident in the handle for prediction. */
int gpedm_likegrad(gpedm_handle h, const double* parst, double modeprior, const double* fixedpars,
                   double rhofixed, int full, gpedm_result* out);

/* ---- fmingrad_Rprop (reference R/GPfunctions.R:522-584) ------------------------------ */
/* 1 initial + <=maxiter loop + 1 final (full) evaluation.  opts may be NULL (defaults). */
int gpedm_fit_rprop(gpedm_handle h, const double* initparst, double modeprior, const double* fixedpars,
                    double rhofixed, const gpedm_rprop_options* opts, gpedm_result* out);

/* ---- products of the last full evaluation -------------------------------------------- */
enum {
  GPEDM_VEC_MEAN = 0,      /* in-sample mean Cd a (:734), length N                          */
  GPEDM_VEC_VAR = 1,       /* diag(Cd - Cd Sigma^-1 Cd) (:735, :460), length N              */
  GPEDM_VEC_ALPHA = 2,     /* a = Sigma^-1 Y (:647)                                          */
  GPEDM_VEC_DIAG_IKVS = 3, /* diag(Sigma^-1)                                                 */
  GPEDM_VEC_LOO_MEAN = 4,  /* closed-form leave-one-out mean  Y_i - a_i / Sigma^-1_ii       */
  GPEDM_VEC_LOO_VAR = 5    /* closed-form leave-one-out latent variance 1/Sigma^-1_ii - ve  */
};
int gpedm_get_vector(gpedm_handle h, int which, double* out);

enum {
  GPEDM_MAT_CD = 0,      /* covm$Cd      (:635)  N x N                                       */
  GPEDM_MAT_SIGMA = 1,   /* covm$Sigma   (:638)                                              */
  GPEDM_MAT_IKVS = 2,    /* covm$iKVs    (:645)                                              */
  GPEDM_MAT_POPSAME = 3, /* covm$popsame (:796)                                              */
  GPEDM_MAT_L = 4,       /* lower Cholesky factor (transpose of R's upper `L`, :827)         */
  GPEDM_MAT_LINV = 5     /* L^-1 (lower)                                                      */
};
/* out: N x N column-major, fully populated (symmetric matrices mirrored, triangular zero-filled). */
int gpedm_get_matrix(gpedm_handle h, int which, double* out);

/* ---- getcov (reference R/GPfunctions.R:779-816), stateless ---------------------------- */
/* Cd (M x N, column-major; rows = X2 / "new" points, columns = X / training points). */
int gpedm_getcov(int device, int32_t d, const double* phi, double sigma2, double rho, int64_t N, const double* X,
                 const int32_t* pop, int64_t M, const double* X2, const int32_t* pop2, int32_t np,
                 const double* rhomat, double* Cd_out);

/* ---- getcovinv (reference R/GPfunctions.R:818-830), stateless --------------------------- */
/* Sigma: n x n column-major (symmetrised internally as the reference does at :641 / :1076).
 * L_out: lower Cholesky factor (= transpose of R's upper `L`), iKVs_out: full inverse; either may
 * be NULL.  Returns >0 (failing pivot) when Sigma is not positive definite. */
int gpedm_covinv(int device, int64_t n, const double* Sigma, double* L_out, double* iKVs_out);

/* ---- predict.GP numerical cores -------------------------------------------------------- */
/* newdata (reference R/GPfunctions.R:1012-1036 + getGPgrad :1863-1865).  Xnew: M x d
 * column-major (already scaled), popnew codes 0..np (np = a population absent from the training set:
 * rho cross-covariance, or 1 under a rhomatrix, as getcov :796-811 gives).  mean, var: length M (var is the latent
 * variance `yvar`; predfsd = sqrt(var), predsd = sqrt(var + ve), :1041-1042).
 * gpgrad: M x d column-major or NULL. */
int gpedm_predict_new(gpedm_handle h, int64_t M, const double* Xnew, const int32_t* popnew, double* mean,
                      double* var, double* gpgrad);

/* leave-one-out with Theiler window (reference :1056-1086).  time: length N (used only to
 * match augmentation duplicates, may be NULL), primary: length N 0/1 (NULL = all primary;
 * primary rows must come first, as in the reference).  mean/var: length nd = #primary.
 * gpgrad: nd x d column-major or NULL. */
int gpedm_predict_loo(gpedm_handle h, int32_t exclradius, const double* time, const int32_t* primary,
                      double* mean, double* var, double* gpgrad);

/* leave-timepoint-out (reference :1140-1176).  time: length N (required). */
int gpedm_predict_lto(gpedm_handle h, int32_t exclradius, const double* time, const int32_t* primary,
                      double* mean, double* var, double* gpgrad);

/* sequential / leave-future-out (reference :1196-1245).  time: length N.  Outputs over all N
 * rows: ymean (NaN where the reference leaves NA) and ysd2 (the reference's `ysd2`, which for
 * never-updated rows is sigma2+ve, a variance - quirk kept). */
int gpedm_predict_sequential(gpedm_handle h, const double* time, double* ymean, double* ysd2);

/* ---- batched independent fits (E/tau grids, multistart, per-population models) -------- */
/* One descriptor per fit: its own training set plus optimiser inputs.  Fits are independent
 * (reference vignettes/choosingE.Rmd:63-86; R/rhomatrix.R:169-182). */
typedef struct gpedm_fit_desc {
  int64_t N;
  int32_t d;
  int32_t np;
  const double* X;        /* N x d column-major */
  const double* Y;        /* N */
  const int32_t* pop;     /* N */
  const double* rhomat;   /* np x np or NULL */
  const double* initparst;/* d+3 */
  const double* fixedpars;/* d+2 or NULL */
  double modeprior;
  double rhofixed;        /* NaN = free */
  int32_t want_loo;       /* 1: also compute closed-form LOO (exclradius 0) mean/var */
  double* loo_mean;       /* N or NULL */
  double* loo_var;        /* N or NULL */
  double* insamp_mean;    /* N or NULL */
  double* insamp_var;     /* N or NULL */
} gpedm_fit_desc;

/* Shards `nfits` fits over `ngpus` devices (devices[] may be NULL = 0..ngpus-1): one worker
 * thread per GPU pulling from a shared queue ordered by estimated cost; results[] (nfits) is
 * filled on return.  With ngpus > 1 the fixed-width result records are exchanged with one
 * NCCL all-gather over NVLink (single process, ncclCommInitAll once per process and device list);
 * no other collective. */
int gpedm_fit_batch(int32_t nfits, const gpedm_fit_desc* descs, int32_t ngpus, const int32_t* devices,
                    const gpedm_rprop_options* opts, gpedm_result* results);
/* The worker threads of every GPU live in the calling process, so results[] is always delivered from
 * host memory; the NCCL all-gather additionally leaves every record on every GPU and is verified
 * against the host copy.  Status of the collective of the most recent gpedm_fit_batch / gpedm_fit_grid:
 * 0 = done and verified, 1 = not needed (one GPU), < 0 = NCCL unavailable or failed (records were still
 * delivered; set GPEDM_REQUIRE_NCCL=1 to turn this into a failure of the call). */
int gpedm_last_gather_status(void);

/* ---- on-device data preparation + model-selection grids -------------------------------- */
/* fitGP(data, y, E, tau, scaling) builds its training set from the raw series: lags inside each
 * population (makelags, reference R/GPfunctions.R:1824-1857), centring / scaling (:359-405) and
 * removal of incomplete rows (:445-451).  gpedm_create_lagged does this on the device, straight
 * into the handle's buffers.  y: T raw values (NaN = missing); pop: T codes 0..np-1 in row order
 * (NULL = one population); scaling: GPEDM_SCALE_*; predictors are the E lags tau, 2 tau, .. E tau
 * of y.  The handle then behaves like one made by gpedm_create. */
enum { GPEDM_SCALE_NONE = 0, GPEDM_SCALE_GLOBAL = 1, GPEDM_SCALE_LOCAL = 2 };
int gpedm_create_lagged(int device, int64_t T, const double* y, const int32_t* pop, int32_t np, int32_t E,
                        int32_t tau, int32_t scaling, const double* rhomat, gpedm_handle* out);
/* What gpedm_create_lagged built: rows (gpedm_n entries: original row of each training row),
 * center / scale ((E+1) x G column-major, column 0 = y, k = lag k; G = np for local scaling,
 * else 1), the scaled training set Y (N), X (N x E), pop (N).  Any pointer may be NULL. */
int gpedm_lagged_info(gpedm_handle h, int64_t* rows, double* center, double* scale, double* Y, double* X,
                      int32_t* pop);

/* In-sample and leave-one-out fit statistics in the ORIGINAL units of y (getR2, reference
 * :1366-1370, on the back-transformed predictions :469-482), reduced on the device from the
 * resident full evaluation.  For handles not made by gpedm_create_lagged the units are those of Y. */
typedef struct gpedm_fit_stats {
  double n, obs_mean, ss_total, ss_insample, ss_loo;
  double R2_insample, R2_loo, rmse_insample, rmse_loo;
} gpedm_fit_stats;
int gpedm_get_fit_stats(gpedm_handle h, gpedm_fit_stats* out);

/* The E x tau model-selection grid of vignettes/choosingE.Rmd:63-86 as one call: the raw series
 * is uploaded once per GPU, every cell (E[c], tau[c]) is prepared on the device, fitted from the
 * default start (:409) with predictmethod "loo" statistics, and the fixed-width records are
 * exchanged as in gpedm_fit_batch.  results / stats: ncells entries. */
int gpedm_fit_grid(int64_t T, const double* y, const int32_t* pop, int32_t np, int32_t ncells, const int32_t* E,
                   const int32_t* tau, int32_t scaling, double modeprior, int32_t ngpus, const int32_t* devices,
                   const gpedm_rprop_options* opts, gpedm_result* results, gpedm_fit_stats* stats);

/* ---- measurement helpers ---------------------------------------------------------------- */
enum {
  GPEDM_PHASE_ASSEMBLE = 0,
  GPEDM_PHASE_POTRF = 1,
  GPEDM_PHASE_TRTRI = 2,
  GPEDM_PHASE_LAUUM = 3,
  GPEDM_PHASE_VEC = 4,
  GPEDM_PHASE_TRACE = 5,
  GPEDM_PHASE_TOTAL = 6,
  GPEDM_NPHASE = 7
};
/* Device time (ms, CUDA events on the handle's stream) of each phase of the last evaluation. */
int gpedm_last_phase_ms(gpedm_handle h, float* ms /* GPEDM_NPHASE */);
/* Number of kernel launches issued by the handle so far. */
int64_t gpedm_launch_count(gpedm_handle h);
/* FP64 issue-rate probes used as roofline denominators: which = 0 DMMA m8n8k4, 1 DFMA. */
int gpedm_fp64_peak(int device, int which, double* tflops);
/* Device-resident evaluation loop for benchmarking: `steps` gradient-only evaluations at
 * parst with no host work in between; returns the average device ms per evaluation. */
int gpedm_bench_evals(gpedm_handle h, const double* parst, double modeprior, int steps, float* ms_per_eval);

#ifdef __cplusplus
}
#endif
#endif /* GPEDM_H */
