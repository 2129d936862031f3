# Accessors over the returned object; same behaviour as the original package's methods, including that
# quantile() summarises the rows flagged burnin == TRUE.
print.mcmcglm <- function(x, ...) {
  cat("Object of class 'mcmcglm'\n\nCall:  ", paste(deparse(x$call), collapse = "\n"),
      "\n\nAverage of parameter samples:\n", sep = "")
  print(x$beta_mean)
  cat("\n")
  invisible(x)
}

samples <- function(x) UseMethod("samples")
samples.mcmcglm <- function(x) x$beta_samples
coef.mcmcglm <- function(object, ...) object$beta_mean

quantile.mcmcglm <- function(x, probs = c(0.025, 0.5, 0.975), ...) {
  p <- ncol(x$model_matrix)
  s <- samples(x)
  s <- s[s$burnin, seq_len(p), drop = FALSE]
  q <- t(vapply(s, function(v) c(mean(v), stats::quantile(v, probs = probs, names = FALSE, ...)),
                numeric(1 + length(probs))))
  out <- data.frame(var = names(s), q, row.names = NULL, check.names = FALSE)
  names(out) <- c("var", "mean", paste0("q_", gsub("\\.", "", probs)))
  out
}
