## tools/make_golden.R -- NOT runnable in this repository's CI (needs R + linxihui/NNLM).
## Emits true NNLM outputs for the committed fixtures so the oracle (and the CUDA path) can be pinned against
## the real binary:   Rscript tools/make_golden.R tests/golden
## Reads each tests/golden/m*.npz through reticulate/numpy, runs NNLM::nnmf from the identical W0/H0 and writes
## <name>.nnlm.rds / .csv next to it.
library(NNLM); library(reticulate)
np <- import("numpy")
dir <- commandArgs(trailingOnly = TRUE)[1]
for (f in list.files(dir, pattern = "^m.*\\.npz$", full.names = TRUE)) {
  g <- np$load(f, allow_pickle = TRUE)
  A <- g$f[["A"]]; k <- as.integer(g$f[["k"]]); loss <- as.character(g$f[["loss"]])
  z <- nnmf(A, k, alpha = as.numeric(g$f[["alpha"]]), beta = as.numeric(g$f[["beta"]]), loss = loss, method = "scd",
            init = list(W = g$f[["W0"]], H = g$f[["H0"]]), max.iter = as.integer(g$f[["max_iter"]]), rel.tol = -1,
            trace = as.integer(g$f[["trace"]]), verbose = 0, n.threads = 1)
  saveRDS(z, sub("\\.npz$", ".nnlm.rds", f))
  write.csv(data.frame(target = z$target.loss, mse = z$mse, mkl = z$mkl), sub("\\.npz$", ".nnlm.trace.csv", f), row.names = FALSE)
}
