# R overlay: the reference's internal numerical functions rewired to libgpedm_b200 through the .Call
# glue in r/gpedm_b200_glue.c.  fitGP(), predict.GP()'s front and back end, update(), summary(), plot(),
# scaling, lag generation and back-transforms stay exactly as in the reference (R/GPfunctions.R:221-520,
# :880-1341); only the numerical core below the call at R/GPfunctions.R:454 changes.
#
# Two pieces:
#   1. this file - new bodies for fmingrad_Rprop / getcov / getcovinv (source it after the package code, or
#      drop it into R/ as zzz_overlay.R);
#   2. r/predict_GP.ed - an ed-style edit script for R/GPfunctions.R that swaps the numeric blocks of
#      fitGP (:460-461) and predict.GP (:884, :1012-1030, :1057-1086, :1144-1176, :1203-1245) for the
#      .predict_* helpers below (apply with r/apply_overlay.py; line anchors are checked).
# Without (2) everything still works - `object$covm` is an environment whose four matrices materialise on
# first use, so the unmodified predict.GP finds `covm$iKVs`, `covm$Cd`, `covm$Sigma` - but then the
# leave-out loops refactorise in R (through the GPU getcovinv) instead of using the closed forms.
# The glue is executed in the build image against a stand-in for R's C API (tests/test_r_glue.py); this R
# code itself has not run (no R there) - see INTEGRATION.md.

.pop_codes <- function(pop) match(pop, unique(pop)) - 1L

.rhomat_in_code_order <- function(rhomatrix, pop) {
  if (is.null(rhomatrix)) return(NULL)
  up <- as.character(unique(pop))                 # reference addresses rhomatrix by dimnames (:799-808)
  rhomatrix[up, up, drop = FALSE]
}

# The four N x N matrices of `covm` (:638-645, :815) stay on the device; each is copied to the host the
# first time R code touches it and cached (N = 32768 would otherwise mean 34 GB of host memory per fit).
.lazy_covm <- function(h) {
  covm <- new.env(parent = emptyenv())
  ids <- c(Cd = 0L, Sigma = 1L, iKVs = 2L, popsame = 3L)       # GPEDM_MAT_* in include/gpedm.h
  for (nm in names(ids)) local({
    name <- nm; id <- ids[[nm]]; cache <- NULL
    makeActiveBinding(name, function() {
      if (is.null(cache)) cache <<- .Call("gpedmb200_matrix", h, id)
      cache
    }, covm)
  })
  covm
}

# replaces fmingrad_Rprop (R/GPfunctions.R:522-584); returns what fitGP reads at :455-509
fmingrad_Rprop <- function(initparst, Y, X, pop, modeprior, fixedpars, rhofixed, rhomatrix) {
  h <- .Call("gpedmb200_create", X, Y, .pop_codes(pop), length(unique(pop)),
             .rhomat_in_code_order(rhomatrix, pop), 0L)
  res <- .Call("gpedmb200_fit", h, initparst, modeprior, as.numeric(fixedpars),
               if (is.null(rhofixed)) NA_real_ else rhofixed)
  d <- ncol(X)
  nm <- c(paste0("phi", 1:d), "ve", "sigma2", "rho")
  names(res$pars) <- nm; names(res$parst) <- nm; names(res$grad) <- nm
  res$mean <- .Call("gpedmb200_vector", h, 0L)            # Cd %*% a              (:734)
  res$covdiag <- .Call("gpedmb200_vector", h, 1L)         # diag(Cd - Cd iKVs Cd) (:735): all fitGP uses of `cov` (:460-461)
  if (!isTRUE(getOption("gpedmb200.patched")))            # unpatched fitGP still evaluates sqrt(diag(output$cov))
    res$cov <- diag(res$covdiag, nrow = length(Y))
  res$handle <- h                                         # external pointer; finalizer frees the device state
  res$covm <- .lazy_covm(h)
  res
}

# replaces getcov (R/GPfunctions.R:779-816); a new-data population absent from `pop` gets code np
getcov <- function(phi, sigma2, rho, X, X2, pop, pop2, rhomat = NULL) {
  up <- unique(pop)
  c2 <- match(pop2, up) - 1L
  c2[is.na(c2)] <- length(up)
  Cd <- .Call("gpedmb200_getcov", phi, sigma2, rho, X, match(pop, up) - 1L, X2, c2,
              length(up), .rhomat_in_code_order(rhomat, pop), 0L)
  list(popsame = outer(pop2, pop, "==") * 1, Cd = Cd)
}

# replaces getcovinv (R/GPfunctions.R:818-830); the reference returns the UPPER factor
getcovinv <- function(Sigma) {
  r <- .Call("gpedmb200_covinv", Sigma, 0L)
  list(L = t(r$L), iKVs = r$iKVs)
}

# numeric blocks of predict.GP, called from the lines r/predict_GP.ed puts in their place
.predict_new <- function(object, Xnew, Popnew, returnGPgrad) {          # :1012-1030
  up <- unique(object$inputs$Pop)
  codes <- match(Popnew, up) - 1L
  codes[is.na(codes)] <- length(up)
  .Call("gpedmb200_predict_new", object$handle, Xnew, codes, returnGPgrad)
}
.predict_loo <- function(object, exclradius, returnGPgrad)              # :1067-1086
  .Call("gpedmb200_predict_leaveout", object$handle, 0L, as.integer(exclradius), as.numeric(object$inputs$Time),
        as.integer(object$inputs$Primary), returnGPgrad)
.predict_lto <- function(object, exclradius, returnGPgrad)              # :1153-1176
  .Call("gpedmb200_predict_leaveout", object$handle, 1L, as.integer(exclradius), as.numeric(object$inputs$Time),
        as.integer(object$inputs$Primary), returnGPgrad)
.predict_sequential <- function(object)                                 # :1223-1245
  .Call("gpedmb200_predict_sequential", object$handle, as.numeric(object$inputs$Time))

# E x tau model-selection grid (vignettes/choosingE.Rmd:63-86) as one call: lags, scaling and complete-row
# selection happen on the device, fits are sharded over `ngpus`
fitGP_grid <- function(y, Es, taus, pop = NULL, scaling = c("global", "local", "none"), modeprior = 1, ngpus = 1L) {
  scaling <- match.arg(scaling)
  cells <- expand.grid(tau = taus, E = Es)
  res <- .Call("gpedmb200_fit_grid", as.numeric(y), if (is.null(pop)) NULL else .pop_codes(pop),
               if (is.null(pop)) 1L else length(unique(pop)), as.integer(cells$E), as.integer(cells$tau),
               match(scaling, c("none", "global", "local")) - 1L, modeprior, as.integer(ngpus))
  cells$R2_insample <- vapply(res, function(r) r$stats[1], 0)
  cells$R2_loo <- vapply(res, function(r) r$stats[2], 0)
  cells$ln_post <- vapply(res, function(r) -r$fit$nllpost, 0)
  cells$iter <- vapply(res, function(r) r$fit$iter, 0L)
  cells
}
