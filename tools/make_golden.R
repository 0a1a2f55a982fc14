## tools/make_golden.R -- NOT runnable in this repository's build image (needs R + linxihui/NNLM + reticulate/numpy).
## Emits TRUE NNLM outputs for the committed fixtures, so that the oracle (and through it the CUDA path) can be pinned
## against the real package:
##     Rscript tools/make_golden.R tests/golden
## For each tests/golden/m*.npz (inputs A, W0, H0, k, alpha, beta, loss, max_iter, trace written by tests/golden/make_golden.py)
## it runs NNLM::nnmf from the identical start with the identical knobs and writes <name>.nnlm.npz next to it with the same
## field names the oracle fixtures use.  tests/test_oracle.py::test_true_nnlm_goldens_pin_the_oracle and
## tests/test_gpu_parity.py::test_true_nnlm_goldens pick these files up when they exist (and say "parity unpinned" when not).
## Also records NNLM's version and its default `trace` (the two details SURVEY Appendix A recalls without certainty).
library(NNLM); library(reticulate)
np <- import("numpy")
dir <- commandArgs(trailingOnly = TRUE)[1]
for (f in list.files(dir, pattern = "^m[a-z0-9_]*\\.npz$", full.names = TRUE)) {
  g <- np$load(f, allow_pickle = TRUE)
  A <- g$f[["A"]]; k <- as.integer(g$f[["k"]]); loss <- as.character(g$f[["loss"]])
  z <- nnmf(A, k, alpha = as.numeric(g$f[["alpha"]]), beta = as.numeric(g$f[["beta"]]), loss = loss, method = "scd",
            init = list(W = g$f[["W0"]], H = g$f[["H0"]]), max.iter = as.integer(g$f[["max_iter"]]), rel.tol = -1,
            trace = as.integer(g$f[["trace"]]), verbose = 0, n.threads = 1)
  np$savez_compressed(sub("\\.npz$", ".nnlm.npz", f), W = z$W, H = z$H, mse = z$mse, mkl = z$mkl,
                      target_loss = z$target.loss, average_epochs = z$average.epochs, n_iteration = as.integer(z$n.iteration),
                      nnlm_version = as.character(packageVersion("NNLM")), default_trace = as.integer(formals(nnmf)$trace))
}
