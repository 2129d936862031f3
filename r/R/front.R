# mcmcglm() with the signature of the original package; the (k, j) sampling loop runs on the GPU.
# Engine-only arguments come after `burnin`.  rng = "R" draws the uniforms with R's own runif() and has
# the engine replay them, so a seeded call reproduces the original package's chain (and, like it,
# row 1 of the samples is the prior draw made with distributional::generate()).
mcmcglm <- function(formula, family = gaussian, data,
                    beta_prior = distributional::dist_normal(0, 1),
                    log_likelihood_extra_args = list(sd = 1),
                    linear_predictor_calc = c("update", "naive"),
                    sample_method = c("slice_sampling", "normal-normal"),
                    qslice_fun = qslice::slice_stepping_out, ...,
                    n_samples = 500, burnin = 100,
                    n_chains = 1L, device = 0L, K = 8L, rng = c("philox", "R"), seed = NULL) {
  cl <- match.call()
  linear_predictor_calc <- match.arg(linear_predictor_calc)
  sample_method <- match.arg(sample_method)
  rng <- match.arg(rng)
  if (burnin >= n_samples) stop("Need more iterations than burnin")
  dots <- list(...)
  if (length(dots) == 0 && sample_method == "slice_sampling")
    stop("A tuning parameter for the `qslice_fun` is missing. For default choice of `qslice::slice_stepping_out` a slice width w needs to be provided")
  if (sample_method != "slice_sampling") stop("sample_method = 'normal-normal' is not part of the GPU path")
  if (linear_predictor_calc != "update") stop("linear_predictor_calc = 'naive' is not part of the GPU path")

  fam <- cgg_family_code(family)
  if (missing(data)) data <- environment(formula)
  mf <- stats::model.frame(formula, data = data)
  Y <- as.double(stats::model.response(mf, "any"))
  X <- stats::model.matrix(attr(mf, "terms"), mf)
  storage.mode(X) <- "double"
  p <- ncol(X)

  cfg <- cgg_config(nrow(X), p, fam$code, cgg_prior_code(beta_prior), cgg_sampler_args(qslice_fun, dots),
                    sd = if (is.null(log_likelihood_extra_args$sd)) 1 else log_likelihood_extra_args$sd,
                    n_chains = n_chains, K = K, device = device,
                    seed = if (is.null(seed)) sample.int(.Machine$integer.max, 1) else seed)
  h <- .Call(C_cgg_create, cfg)
  on.exit(.Call(C_cgg_destroy, h), add = TRUE)
  .Call(C_cgg_set_data, h, X, Y)
  beta0 <- matrix(NA_real_, n_chains, p)
  for (ch in seq_len(n_chains)) {
    beta0[ch, ] <- distributional::generate(beta_prior, p)[[1]]
    .Call(C_cgg_init_chain, h, ch, beta0[ch, ])
  }
  replay <- NULL
  if (rng == "R")  # generous upper bound; the engine reports how many were consumed
    replay <- matrix(stats::runif(n_chains * n_samples * p * 40), ncol = n_chains)
  res <- .Call(C_cgg_run, h, n_samples, n_chains, p, replay)

  arr <- aperm(array(res$samples, c(p, n_samples, n_chains)), c(2, 1, 3))   # iteration x coef x chain
  first <- rbind(beta0[1, ], arr[, , 1])
  beta_samples <- stats::setNames(as.data.frame(first), colnames(X))
  beta_samples$iteration <- seq_len(nrow(first)) - 1L
  beta_samples$burnin <- beta_samples$iteration <= burnin + 1
  keep <- !beta_samples$burnin
  beta_mean <- as.data.frame(lapply(beta_samples[keep, seq_len(p), drop = FALSE], mean), check.names = FALSE)

  out <- c(list(beta_samples = beta_samples, beta_mean = beta_mean, data = data, model_matrix = X,
                param_list = NULL, family = fam$object, formula = formula, call = cl, burnin = burnin,
                sample_method = sample_method, qslice_fun = qslice_fun),
           dots,
           list(chains = arr, engine_stats = res[-1]))
  structure(out, class = c("mcmcglm", class(out)))
}

# The exported operators, evaluated by the same kernels (j is 1-based as in the original package).
log_potential_from_betaj <- function(new_beta_j, j, current_beta, current_eta, Y, X, family, beta_prior,
                                     linear_predictor_calc = "update", ...) {
  if (linear_predictor_calc != "update") stop("linear_predictor_calc = 'naive' is not part of the GPU path")
  extra <- list(...)
  storage.mode(X) <- "double"
  cfg <- cgg_config(nrow(X), ncol(X), cgg_family_code(family)$code, cgg_prior_code(beta_prior),
                    c(w = 1, max_steps = -1), sd = if (is.null(extra$sd)) 1 else extra$sd,
                    n_chains = 1L, K = 8L, device = 0L, seed = 0)
  cfg$driver <- 1
  h <- .Call(C_cgg_create, cfg)
  on.exit(.Call(C_cgg_destroy, h), add = TRUE)
  .Call(C_cgg_set_data, h, X, as.double(Y))
  .Call(C_cgg_set_state, h, 1L, as.double(current_beta), as.double(current_eta))
  .Call(C_cgg_log_potential, h, 1L, as.integer(j), as.double(new_beta_j))
}

update_linear_predictor <- function(new_beta_j, current_beta_j, current_eta, X_j) {
  X <- matrix(as.double(X_j), ncol = 1)
  cfg <- cgg_config(nrow(X), 1L, c(family = 0, link = 0), c(prior = 0, prior_mu = 0, prior_sigma = 1, prior_df = 1),
                    c(w = 1, max_steps = -1), sd = 1, n_chains = 1L, K = 8L, device = 0L, seed = 0)
  cfg$driver <- 1
  h <- .Call(C_cgg_create, cfg)
  on.exit(.Call(C_cgg_destroy, h), add = TRUE)
  .Call(C_cgg_set_data, h, X, numeric(nrow(X)))
  .Call(C_cgg_set_state, h, 1L, as.double(current_beta_j), as.double(current_eta))
  .Call(C_cgg_update_eta, h, 1L, 1L, as.double(new_beta_j))
  .Call(C_cgg_get_state, h, 1L, nrow(X), 1L)[[2]]
}

mcmcglm_across_tuningparams <- function(..., tuning_parameter_name = "w") {
  args <- list(...)
  values <- args[[1]]
  rest <- args[-1]
  out <- lapply(values, function(v) do.call(mcmcglm, c(stats::setNames(list(v), tuning_parameter_name), rest)))
  attr(out, "tuning_parameter_name") <- tuning_parameter_name
  out
}
