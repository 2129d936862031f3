#!/usr/bin/env Rscript
# run_reference.R -- times the TRUE reference (mathiaslj/mcmcglm, R) on the host cores for the metric of BASELINE.json:
# coordinate updates / second, chains spread with parallel::mclapply (north star).  Needs R with mcmcglm, qslice and
# distributional -- absent from this repository's build image, so the numbers this prints come from other hosts; the
# in-image stand-in is `bench.py --impl reference` (C restatement of the same algorithm, oracle/oracle.c).
#
#     Rscript baseline/run_reference.R [workload = cfg2] [n_samples = 2] [cores = detectCores()]
suppressPackageStartupMessages({ library(mcmcglm); library(qslice); library(distributional); library(parallel) })
args <- commandArgs(trailingOnly = TRUE)
workload <- if (length(args) >= 1) args[[1]] else "cfg2"
n_samples <- if (length(args) >= 2) as.integer(args[[2]]) else 2L
cores <- if (length(args) >= 3) as.integer(args[[3]]) else detectCores()
cfg <- switch(workload,
  cfg1 = list(family = "gaussian", n = 1e3, p = 3, prior = dist_normal(0, 1), chains = 1),
  cfg2 = list(family = "binomial", n = 1e5, p = 100, prior = dist_normal(0, 1), chains = 4),
  cfg3 = list(family = "binomial", n = 1e6, p = 1000, prior = dist_laplace(0, 1), chains = 8),
  cfg4 = list(family = "poisson", n = 1e6, p = 500, prior = dist_student_t(4, 0, 1), chains = 8),
  stop("unknown workload"))
set.seed(42)
n <- cfg$n; p <- cfg$p
X <- matrix(rnorm(n * (p - 1)), n, p - 1)
beta <- rnorm(p) / sqrt(p)
eta <- drop(cbind(1, X) %*% beta)
y <- switch(cfg$family, gaussian = eta + rnorm(n), binomial = rbinom(n, 1, 1 / (1 + exp(-eta))), poisson = rpois(n, exp(eta)))
dat <- data.frame(Y = y, X)
chains <- max(cfg$chains, cores)
t0 <- Sys.time()
fits <- mclapply(seq_len(chains), function(c) {
  set.seed(1000 + c)
  mcmcglm(Y ~ ., family = cfg$family, data = dat, beta_prior = cfg$prior, qslice_fun = qslice::slice_stepping_out, w = 0.5,
          n_samples = n_samples, burnin = 0)
}, mc.cores = cores)
wall <- as.numeric(difftime(Sys.time(), t0, units = "secs"))
updates <- chains * n_samples * p
cat(sprintf('{"impl": "reference-R", "workload": "%s", "metric": "coordinate updates/sec", "value": %.6g, "unit": "updates/s", "cores": %d, "chains": %d, "n_samples": %d, "wall_s": %.3f, "R": "%s", "mcmcglm": "%s", "qslice": "%s"}\n',
            workload, updates / wall, cores, chains, n_samples, wall, R.version.string,
            as.character(packageVersion("mcmcglm")), as.character(packageVersion("qslice"))))
