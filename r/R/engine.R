# Translation of R-level objects (family, prior, sampler) into the enums and doubles of cggibbs.h.
# Anything outside the supported set stops here: the engine has no CPU fallback.

cgg_family_code <- function(family) {
  fam <- if (is.character(family)) get(family, mode = "function", envir = parent.frame())() else
         if (is.function(family)) family() else family
  if (is.null(fam$family)) stop("'family' not recognized")
  key <- paste(fam$family, fam$link)
  code <- switch(key,
    "gaussian identity" = c(family = 0, link = 0),
    "binomial logit"    = c(family = 1, link = 1),
    "poisson log"       = c(family = 2, link = 2),
    stop("family/link '", key, "' is not supported by the GPU engine ",
         "(supported: gaussian/identity, binomial/logit, poisson/log)"))
  list(object = fam, code = code)
}

cgg_prior_code <- function(beta_prior) {
  if (length(beta_prior) != 1L)
    stop("only a single iid prior is supported by the GPU engine (got a list of ", length(beta_prior), ")")
  kind <- class(vctrs::vec_data(beta_prior)[[1]])[1]
  pars <- distributional::parameters(beta_prior)
  switch(kind,
    dist_normal    = c(prior = 0, prior_mu = pars$mu, prior_sigma = pars$sigma, prior_df = 1),
    dist_laplace   = c(prior = 1, prior_mu = pars$mu, prior_sigma = pars$sigma, prior_df = 1),
    dist_student_t = c(prior = 2, prior_mu = pars$mu, prior_sigma = pars$sigma, prior_df = pars$df),
    stop("prior '", kind, "' is not supported by the GPU engine (supported: dist_normal, dist_laplace, dist_student_t)"))
}

cgg_sampler_args <- function(qslice_fun, dots) {
  ok <- requireNamespace("qslice", quietly = TRUE) && identical(qslice_fun, qslice::slice_stepping_out)
  if (!ok) stop("only qslice::slice_stepping_out is implemented on the GPU")
  if (is.null(dots$w)) stop("slice_stepping_out needs the slice width `w`")
  extra <- setdiff(names(dots), c("w", "max"))
  if (length(extra)) stop("unknown tuning argument(s) for slice_stepping_out: ", paste(extra, collapse = ", "))
  max <- if (is.null(dots$max)) Inf else dots$max
  c(w = dots$w, max_steps = if (is.finite(max)) max else -1)
}

cgg_config <- function(n, p, fam_code, prior_code, sampler, sd, n_chains, K, device, seed) {
  as.list(c(n = n, p = p, fam_code, prior_code, sampler, sd = sd, n_chains = n_chains, K = K,
            device = device, driver = 0, chain_offset = 0, seed = seed, spec_tau = 0.12))
}
