#!/usr/bin/env Rscript
# record_reference.R -- records TRUE-REFERENCE fixtures of the CGGibbs hot path for tests/test_oracle_fixtures.py.
#
# Needs an R installation with mcmcglm (mathiaslj/mcmcglm), qslice and distributional; none of them exists in the build
# image of this repository, so this script has NOT been executed there (see DESIGN.md, "Oracle and parity status").
# Run it on any R-equipped host from the repository root:
#
#     Rscript tools/record_reference.R [outdir = tests/fixtures]
#
# For every case below it runs the reference's own mcmcglm() (R/mcmcglm.R:147-299) and writes
#     <outdir>/<case>.json   manifest: sizes, family, prior, w, max, package versions, array table
#     <outdir>/<case>.bin    little-endian float64 arrays, concatenated in the order of the manifest:
#         X        n x p column-major (model.matrix, R/mcmcglm.R:176-178)
#         y        n
#         beta0    p          row 0 of beta_samples = the prior draw (R/mcmcglm.R:200-222)
#         uniforms every runif(1) qslice::slice_stepping_out drew, in order
#         evals    3 x n_evals: (j [1-based], b, f) of every log_target evaluation (log_potential_from_betaj, R/glm_utils.R:187-218)
#         samples  (n_samples + 1) x p row-major: beta_samples without the iteration / burnin columns
# The slice sampler is wrapped, not replaced: mcmcglm() accepts any `qslice_fun(x, log_target, ...)` (R/mcmcglm.R:258-261),
# so the wrapper logs the evaluations and calls qslice::slice_stepping_out unchanged; runif is traced in the namespaces
# qslice resolves it from.
suppressPackageStartupMessages({ library(mcmcglm); library(qslice); library(distributional) })
args <- commandArgs(trailingOnly = TRUE)
outdir <- if (length(args) >= 1) args[[1]] else file.path("tests", "fixtures")
dir.create(outdir, recursive = TRUE, showWarnings = FALSE)

rec <- new.env()
start_recording <- function() { rec$u <- numeric(0); rec$ev <- list(); rec$j <- 0L }
tracer <- quote({ v <- returnValue(); if (is.numeric(v) && length(v) == 1L) assign("u", c(get("u", envir = rec), v), envir = rec) })
for (ns in c("stats", "qslice")) {
  try(suppressMessages(trace("runif", where = asNamespace(ns), exit = tracer, print = FALSE)), silent = TRUE)
}
recording_slice <- function(x, log_target, ...) {
  rec$j <- rec$j + 1L
  j <- rec$j
  lt <- function(b) { f <- log_target(b); rec$ev[[length(rec$ev) + 1L]] <- c(j, b, f); f }
  qslice::slice_stepping_out(x = x, log_target = lt, ...)
}

make_data <- function(family, n, p, seed) {
  set.seed(seed)
  X <- matrix(rnorm(n * (p - 1)), n, p - 1)
  colnames(X) <- paste0("X", seq_len(p - 1))
  beta <- rnorm(p) / sqrt(p)
  eta <- drop(cbind(1, X) %*% beta)
  y <- switch(family,
              gaussian = eta + rnorm(n),
              binomial = rbinom(n, 1, 1 / (1 + exp(-eta))),
              poisson = rpois(n, exp(eta)))
  data.frame(Y = y, X)
}

record_case <- function(name, family, prior, prior_desc, n, p, w, max = Inf, n_samples = 30, seed = 42) {
  dat <- make_data(family, n, p, seed)
  start_recording()
  extra <- if (is.finite(max)) list(max = max) else list()
  fit <- do.call(mcmcglm, c(list(formula = Y ~ ., family = family, data = dat, beta_prior = prior, qslice_fun = recording_slice,
                                 w = w, n_samples = n_samples, burnin = 5), extra))
  p_ <- ncol(fit$model_matrix)
  S <- as.matrix(fit$beta_samples[, seq_len(p_), drop = FALSE])
  ev <- do.call(cbind, rec$ev)
  ev[1, ] <- ((ev[1, ] - 1) %% p_) + 1                      # call counter -> coordinate
  arrays <- list(X = as.numeric(fit$model_matrix), y = as.numeric(dat$Y), beta0 = as.numeric(S[1, ]),
                 uniforms = rec$u, evals = as.numeric(ev), samples = as.numeric(t(S)))
  con <- file(file.path(outdir, paste0(name, ".bin")), "wb")
  for (a in arrays) writeBin(as.double(a), con, size = 8, endian = "little")
  close(con)
  q <- function(s) paste0('"', s, '"')
  man <- c(sprintf('"name": %s', q(name)), sprintf('"family": %s', q(family)), sprintf('"prior": %s', prior_desc),
           sprintf('"n": %d', n), sprintf('"p": %d', p_), sprintf('"w": %.17g', w),
           sprintf('"max_steps": %s', if (is.finite(max)) as.character(as.integer(max)) else "-1"),
           sprintf('"n_samples": %d', n_samples), sprintf('"seed": %d', seed),
           sprintf('"versions": {"R": %s, "mcmcglm": %s, "qslice": %s, "distributional": %s}', q(R.version.string),
                   q(as.character(packageVersion("mcmcglm"))), q(as.character(packageVersion("qslice"))),
                   q(as.character(packageVersion("distributional")))),
           sprintf('"arrays": [%s]', paste(sprintf('{"name": %s, "count": %d}', q(names(arrays)), lengths(arrays)), collapse = ", ")))
  writeLines(paste0("{", paste(man, collapse = ", "), "}"), file.path(outdir, paste0(name, ".json")))
  message(sprintf("%s: %d uniforms, %d evaluations, %d x %d samples", name, length(rec$u), ncol(ev), nrow(S), p_))
}

record_case("binomial_laplace", "binomial", dist_laplace(0, 1), '{"kind": "laplace", "mu": 0, "sigma": 1}', n = 2000, p = 5, w = 0.5)
record_case("poisson_student_t", "poisson", dist_student_t(4, 0, 1), '{"kind": "student_t", "mu": 0, "sigma": 1, "df": 4}', n = 2000, p = 4, w = 0.5)
record_case("gaussian_normal_max5", "gaussian", dist_normal(0, 1), '{"kind": "normal", "mu": 0, "sigma": 1}', n = 1000, p = 3, w = 0.05, max = 5)
record_case("binomial_normal_max3", "binomial", dist_normal(0, 1), '{"kind": "normal", "mu": 0, "sigma": 1}', n = 1500, p = 4, w = 0.02, max = 3)
record_case("binomial_laplace_prior_start", "binomial", dist_laplace(0, 1), '{"kind": "laplace", "mu": 0, "sigma": 1}', n = 3000, p = 200, w = 0.5, n_samples = 3)
