# Translation of R-level objects (family, prior, sampler) into the enums and doubles of cggibbs.h.
# Anything outside the supported set stops here: the engine has no CPU fallback.
# EXPERIMENTAL: written against include/cggibbs.h, never executed (the build image has no R).

cgg_family_code <- function(family) {
  fam <- if (is.character(family)) get(family, mode = "function", envir = parent.frame())() else
         if (is.function(family)) family() else family
  if (is.null(fam$family)) stop("'family' not recognized")
  # MASS::negative.binomial(theta) reports "Negative Binomial(theta)"; the reference dispatches on the letters only
  name <- gsub("[^[:alpha:]]", "", sub("\\(.*$", "", fam$family))
  key <- paste(name, fam$link)
  code <- switch(key,
    "gaussian identity"      = c(family = 0, link = 0),
    "binomial logit"         = c(family = 1, link = 1),
    "binomial probit"        = c(family = 1, link = 3),
    "poisson log"            = c(family = 2, link = 2),
    "NegativeBinomial log"   = c(family = 3, link = 2),
    stop("family/link '", key, "' is not supported by the GPU engine (supported: gaussian/identity, ",
         "binomial/logit, binomial/probit, poisson/log, negative binomial/log)"))
  list(object = fam, code = code)
}

# one distributional distribution -> c(kind, a, b, c) as cgg_add_prior takes them
cgg_prior_one <- function(d) {
  kind <- stats::family(d)
  pars <- distributional::parameters(d)
  switch(kind,
    normal      = c(0, pars$mu, pars$sigma, 1),
    laplace     = c(1, pars$mu, pars$sigma, 1),
    student_t   = c(2, pars$mu, pars$sigma, pars$df),
    gamma       = c(3, pars$shape, pars$rate, 1),
    exponential = c(4, 0, pars$rate, 1),
    stop("prior '", kind, "' is not supported by the GPU engine (supported: dist_normal, dist_laplace, ",
         "dist_student_t, dist_gamma, dist_exponential and lists of them)"))
}

# beta_prior: one distribution, or a list / vector of them (every prior of a list is applied to every
# coordinate, as log_prior_density.list of the original package does); at most 8
cgg_prior_codes <- function(beta_prior) {
  n <- length(beta_prior)
  if (n > 8L) stop("a list of more than 8 priors is not supported by the GPU engine")
  lapply(seq_len(n), function(k) cgg_prior_one(beta_prior[[k]]))
}

cgg_sampler_args <- function(qslice_fun, dots) {
  ok <- requireNamespace("qslice", quietly = TRUE) && identical(qslice_fun, qslice::slice_stepping_out)
  if (!ok) stop("only qslice::slice_stepping_out is implemented on the GPU")
  if (is.null(dots$w)) stop("slice_stepping_out needs the slice width `w`")
  extra <- setdiff(names(dots), c("w", "max"))
  if (length(extra)) stop("unknown tuning argument(s) for slice_stepping_out: ", paste(extra, collapse = ", "))
  max <- if (is.null(dots$max)) Inf else dots$max
  if (is.finite(max) && max != floor(max)) stop("a finite `max` must be a whole number")
  # qslice: max = Inf steps out without limit; a finite max <= 0 does not step out at all
  c(w = dots$w[[1]], max_steps = if (is.finite(max)) max(max, 0) else -1)
}

cgg_config <- function(n, p, fam_code, prior1, sampler, sd, n_chains, K, device, seed, flags = 0) {
  as.list(c(n = n, p = p, fam_code, prior = prior1[1], prior_mu = prior1[2], prior_sigma = prior1[3], prior_df = prior1[4],
            sampler, sd = sd, n_chains = n_chains, K = K, device = device, driver = 0, chain_offset = 0, seed = seed,
            spec_tau = 0.12, flags = flags))
}
