# R overlay: bodies of the reference's internal functions rewired to libgpedm_b200 through the
# .Call glue in r/gpedm_b200_glue.c.  fitGP(), predict.GP(), update(), summary(), plot(), scaling,
# lag generation and back-transforms stay exactly as in the reference (R/GPfunctions.R:221-520,
# :880-1341); only the numerical core below the call at R/GPfunctions.R:454 changes.
# Untested here (no R in the build image) - see INTEGRATION.md.

.pop_codes <- function(pop) match(pop, unique(pop)) - 1L

.rhomat_in_code_order <- function(rhomatrix, pop) {
  if (is.null(rhomatrix)) return(NULL)
  up <- as.character(unique(pop))                 # reference addresses rhomatrix by dimnames (:799-808)
  rhomatrix[up, up, drop = FALSE]
}

# replaces fmingrad_Rprop (R/GPfunctions.R:522-584)
fmingrad_Rprop <- function(initparst, Y, X, pop, modeprior, fixedpars, rhofixed, rhomatrix) {
  h <- .Call("gpedmb200_create", X, Y, .pop_codes(pop), length(unique(pop)),
             .rhomat_in_code_order(rhomatrix, pop), 0L)
  res <- .Call("gpedmb200_fit", h, initparst, modeprior, as.numeric(fixedpars),
               if (is.null(rhofixed)) NA_real_ else rhofixed)
  d <- ncol(X)
  nm <- c(paste0("phi", 1:d), "ve", "sigma2", "rho")
  names(res$pars) <- nm; names(res$parst) <- nm; names(res$grad) <- nm
  res$mean <- .Call("gpedmb200_vector", h, 0L)            # Cd %*% a            (:734)
  res$covdiag <- .Call("gpedmb200_vector", h, 1L)         # diag(Cd - Cd iKVs Cd) (:735); fitGP uses only the diagonal (:460)
  res$cov <- diag(res$covdiag, nrow = length(Y))          # keeps `sqrt(diag(output$cov))` at :460 working unchanged
  res$handle <- h
  # covm is materialised lazily: each of the four matrices is N x N (:638-645, :815)
  res$covm <- list(handle = h)
  makeActiveBinding("popsame", function() .Call("gpedmb200_matrix", h, 3L), as.environment(res$covm))
  res
}

# replaces getcov (R/GPfunctions.R:779-816)
getcov <- function(phi, sigma2, rho, X, X2, pop, pop2, rhomat = NULL) {
  up <- unique(pop)
  Cd <- .Call("gpedmb200_getcov", phi, sigma2, rho, X, match(pop, up) - 1L, X2, match(pop2, up) - 1L,
              length(up), .rhomat_in_code_order(rhomat, pop), 0L)
  list(popsame = outer(pop2, pop, "==") * 1, Cd = Cd)
}

# numeric blocks of predict.GP (R/GPfunctions.R:1012-1036, :1056-1104, :1140-1194, :1196-1256)
.predict_new <- function(object, Xnew, Popnew, returnGPgrad)
  .Call("gpedmb200_predict_new", object$handle, Xnew, match(Popnew, unique(object$inputs$Pop)) - 1L, returnGPgrad)
.predict_loo <- function(object, exclradius, returnGPgrad)
  .Call("gpedmb200_predict_leaveout", object$handle, 0L, as.integer(exclradius), as.numeric(object$inputs$Time),
        as.integer(object$inputs$Primary), returnGPgrad)
.predict_lto <- function(object, exclradius, returnGPgrad)
  .Call("gpedmb200_predict_leaveout", object$handle, 1L, as.integer(exclradius), as.numeric(object$inputs$Time),
        as.integer(object$inputs$Primary), returnGPgrad)
.predict_sequential <- function(object)
  .Call("gpedmb200_predict_sequential", object$handle, as.numeric(object$inputs$Time))

# replaces getcovinv (R/GPfunctions.R:818-830); the reference returns the UPPER factor
getcovinv <- function(Sigma) {
  r <- .Call("gpedmb200_covinv", Sigma, 0L)
  list(L = t(r$L), iKVs = r$iKVs)
}

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
