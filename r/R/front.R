# mcmcglm() with the signature of the original package; the (k, j) sampling loop runs on the GPU.
# Engine-only arguments come after `burnin`.  rng = "R" draws the uniforms with R's own runif() and has the engine
# replay them, so a seeded call reproduces the original package's chain (and, like it, row 1 of the samples is the prior
# draw made with distributional::generate()).
# EXPERIMENTAL: written against include/cggibbs.h, never executed (the build image has no R).
mcmcglm <- function(formula, family = gaussian, data,
                    beta_prior = distributional::dist_normal(0, 1),
                    log_likelihood_extra_args = list(sd = 1),
                    linear_predictor_calc = c("update", "naive"),
                    sample_method = c("slice_sampling", "normal-normal"),
                    qslice_fun = qslice::slice_stepping_out, ...,
                    n_samples = 500, burnin = 100,
                    n_chains = 1L, device = 0L, K = 8L, rng = c("philox", "R"), seed = NULL,
                    chunk = 10L, w_per_chain = NULL) {
  cl <- match.call()
  linear_predictor_calc <- match.arg(linear_predictor_calc)
  sample_method <- match.arg(sample_method)
  rng <- match.arg(rng)
  if (burnin >= n_samples) stop("Need more iterations than burnin")
  dots <- list(...)
  if (length(dots) == 0 && sample_method == "slice_sampling")
    stop("A tuning parameter for the `qslice_fun` is missing. For default choice of `qslice::slice_stepping_out` a slice width w needs to be provided")
  if (sample_method != "slice_sampling") stop("sample_method = 'normal-normal' is not part of the GPU path")

  fam <- cgg_family_code(family)
  if (missing(data)) data <- environment(formula)
  mf <- stats::model.frame(formula, data = data)
  Y <- as.double(stats::model.response(mf, "any"))
  X <- stats::model.matrix(attr(mf, "terms"), mf)
  storage.mode(X) <- "double"
  p <- ncol(X)

  priors <- cgg_prior_codes(beta_prior)
  list_of_marginal_priors <- length(beta_prior) > 1
  if (list_of_marginal_priors && length(beta_prior) != p)
    stop("The list length of the `beta_prior` specification needs to match the number of parameters in the model (potentially including intercept)")
  flags <- if (linear_predictor_calc == "naive") 16 else 0      # CGG_FLAG_NAIVE
  cfg <- cgg_config(nrow(X), p, fam$code, priors[[1]], cgg_sampler_args(qslice_fun, dots),
                    sd = if (is.null(log_likelihood_extra_args$sd)) 1 else log_likelihood_extra_args$sd,
                    n_chains = n_chains, K = K, device = device,
                    seed = if (is.null(seed)) sample.int(.Machine$integer.max, 1) else seed, flags = flags)
  h <- .Call(C_cgg_create, cfg)
  on.exit(.Call(C_cgg_destroy, h), add = TRUE)
  for (pr in priors[-1]) .Call(C_cgg_add_prior, h, as.integer(pr[1]), pr[2], pr[3], pr[4])
  .Call(C_cgg_set_data, h, X, Y)
  if (!is.null(w_per_chain)) .Call(C_cgg_set_chain_w, h, as.double(w_per_chain))
  beta0 <- matrix(NA_real_, n_chains, p)
  for (ch in seq_len(n_chains)) {
    beta0[ch, ] <- if (list_of_marginal_priors)
      vapply(seq_len(p), function(j) distributional::generate(beta_prior[[j]], 1)[[1]], numeric(1)) else
      distributional::generate(beta_prior, p)[[1]]
    .Call(C_cgg_init_chain, h, ch, beta0[ch, ])
  }

  # the engine runs in chunks of a few iterations: the progress bar moves and an interrupt is honoured between chunks
  has_cli <- requireNamespace("cli", quietly = TRUE)
  if (has_cli) cli::cli_progress_bar("Sampling from posterior", total = n_samples)
  smp <- array(NA_real_, c(p, n_samples, n_chains))
  stats_sum <- NULL
  done <- 0L
  while (done < n_samples) {
    it <- min(as.integer(chunk), n_samples - done)
    replay <- NULL
    if (rng == "R") {
      # R's own stream, drawn chunk by chunk; a generous bound per update, and the draws the chunk did not consume are
      # put back by restoring the RNG state and re-drawing exactly the consumed count (single chain only)
      if (n_chains != 1L) stop("rng = 'R' reproduces the original package's single chain: n_chains must be 1")
      state <- get(".Random.seed", envir = globalenv())
      replay <- matrix(stats::runif(it * p * 64), ncol = 1)
    }
    res <- .Call(C_cgg_run, h, it, n_chains, p, replay)
    if (rng == "R") {
      assign(".Random.seed", state, envir = globalenv())
      used <- res$uniforms_used[1] - (if (is.null(stats_sum)) 0 else stats_sum$uniforms_used)
      if (used > 0) invisible(stats::runif(used))
    }
    smp[, done + seq_len(it), ] <- array(res$samples, c(p, it, n_chains))
    st <- res[-1]
    stats_sum <- if (is.null(stats_sum)) st else {
      keep <- st$uniforms_used
      out <- Map(`+`, stats_sum, st)
      out$uniforms_used <- keep                                  # cumulative already
      out
    }
    done <- done + it
    if (has_cli) cli::cli_progress_update(set = done)
  }

  arr <- aperm(smp, c(2, 1, 3))                                  # iteration x coef x chain
  first <- rbind(beta0[1, ], matrix(arr[, , 1], n_samples, p))
  beta_samples <- stats::setNames(as.data.frame(first), colnames(X))
  beta_samples$iteration <- seq_len(nrow(first)) - 1L
  beta_samples$burnin <- beta_samples$iteration <= burnin + 1
  keep <- !beta_samples$burnin
  beta_mean <- as.data.frame(lapply(beta_samples[keep, seq_len(p), drop = FALSE], mean), check.names = FALSE)

  out <- c(list(beta_samples = beta_samples, beta_mean = beta_mean, data = data, model_matrix = X,
                param_list = NULL, family = fam$object, formula = formula, call = cl, burnin = burnin,
                sample_method = sample_method, qslice_fun = qslice_fun),
           dots,
           list(chains = arr, beta0 = beta0, engine_stats = stats_sum))
  structure(out, class = c("mcmcglm", class(out)))
}

# The exported operators, evaluated by the same kernels (j is 1-based as in the original package).
log_potential_from_betaj <- function(new_beta_j, j, current_beta, current_eta, Y, X, family, beta_prior,
                                     linear_predictor_calc = "update", ...) {
  extra <- list(...)
  storage.mode(X) <- "double"
  priors <- cgg_prior_codes(beta_prior)
  cfg <- cgg_config(nrow(X), ncol(X), cgg_family_code(family)$code, priors[[1]],
                    c(w = 1, max_steps = -1), sd = if (is.null(extra$sd)) 1 else extra$sd,
                    n_chains = 1L, K = 8L, device = 0L, seed = 0)
  cfg$driver <- 1
  h <- .Call(C_cgg_create, cfg)
  on.exit(.Call(C_cgg_destroy, h), add = TRUE)
  for (pr in priors[-1]) .Call(C_cgg_add_prior, h, as.integer(pr[1]), pr[2], pr[3], pr[4])
  .Call(C_cgg_set_data, h, X, as.double(Y))
  if (linear_predictor_calc == "naive") .Call(C_cgg_init_chain, h, 1L, as.double(current_beta))   # eta = X %*% beta on the device
  else .Call(C_cgg_set_state, h, 1L, as.double(current_beta), as.double(current_eta))
  .Call(C_cgg_log_potential, h, 1L, as.integer(j), as.double(new_beta_j))
}

update_linear_predictor <- function(new_beta_j, current_beta_j, current_eta, X_j) {
  X <- matrix(as.double(X_j), ncol = 1)
  cfg <- cgg_config(nrow(X), 1L, c(family = 0, link = 0), c(0, 0, 1, 1),
                    c(w = 1, max_steps = -1), sd = 1, n_chains = 1L, K = 8L, device = 0L, seed = 0)
  cfg$driver <- 1
  h <- .Call(C_cgg_create, cfg)
  on.exit(.Call(C_cgg_destroy, h), add = TRUE)
  .Call(C_cgg_set_data, h, X, numeric(nrow(X)))
  .Call(C_cgg_set_state, h, 1L, as.double(current_beta_j), as.double(current_eta))
  .Call(C_cgg_update_eta, h, 1L, 1L, as.double(new_beta_j))
  .Call(C_cgg_get_state, h, 1L, nrow(X), 1L)[[2]]
}

# The tuning sweep of the original package (lapply / future_lapply over mcmcglm()) as ONE engine run when the tuning
# parameter is `w`: the values become the chains of a single upload, every chain with its own slice width.
mcmcglm_across_tuningparams <- function(..., tuning_parameter_name = "w", parallelise = FALSE, n_cores = NULL) {
  args <- list(...)
  values <- args[[1]]
  rest <- args[-1]
  if (tuning_parameter_name != "w" || length(values) > 32L) {
    out <- lapply(values, function(v) do.call(mcmcglm, c(stats::setNames(list(v), tuning_parameter_name), rest)))
  } else {
    fit <- do.call(mcmcglm, c(list(w = values[[1]], n_chains = length(values), w_per_chain = as.double(values)), rest))
    p <- ncol(fit$model_matrix)
    out <- lapply(seq_along(values), function(k) {
      f <- fit
      f$beta_samples[, seq_len(p)] <- rbind(fit$beta0[k, ], matrix(fit$chains[, , k], ncol = p))
      keep <- !f$beta_samples$burnin
      f$beta_mean <- as.data.frame(lapply(f$beta_samples[keep, seq_len(p), drop = FALSE], mean), check.names = FALSE)
      f$w <- values[[k]]
      f
    })
  }
  attr(out, "tuning_parameter_name") <- tuning_parameter_name
  out
}
