## r/R/nmf_cuda.R -- drop-in overlay for R/nmf.R:96-134 (RunNMF).  COMPILE-UNTESTED here (no R in the image).
## Same formals and return layout; only the NNLM::nnmf call is replaced by the CUDA path.  n.cores is accepted
## and ignored.  options(swne.cuda.device = 0) selects the GPU.

.nnmf_cuda <- function(A, k, alpha = rep(0, 3), beta = rep(0, 3), init = NULL, loss = "mse", max.iter = 500,
                       rel.tol = 1e-4, inner.max.iter = ifelse(loss == "mse", 50, 1), inner.rel.tol = 1e-9,
                       trace = 100 / inner.max.iter, verbose = 0) {
  alpha <- c(as.double(alpha), 0, 0, 0)[1:3]; beta <- c(as.double(beta), 0, 0, 0)[1:3]   # NNLM's padding
  if (is.null(init)) init <- list(W = matrix(runif(nrow(A) * k) * 0.01, nrow(A), k),
                                  H = matrix(runif(k * ncol(A)) * 0.01, k, ncol(A)))       # NNLM's random start
  t0 <- proc.time()
  out <- swne_nmf_cuda(A, k, init$W, init$H, alpha, beta, as.integer(max.iter), rel.tol, as.integer(inner.max.iter),
                       inner.rel.tol, ifelse(loss == "mse", 1L, 3L), as.integer(trace),
                       as.integer(getOption("swne.cuda.device", 0L)))
  dimnames(out$W) <- list(rownames(A), NULL); dimnames(out$H) <- list(NULL, colnames(A))
  structure(list(W = out$W, H = out$H, mse = out$mse_error, mkl = out$mkl_error, target.loss = out$target_error,
                 average.epochs = out$average_epoch, n.iteration = out$n_iteration, run.time = proc.time() - t0,
                 options = list(method = "scd", loss = loss, alpha = alpha, beta = beta, max.iter = max.iter,
                                rel.tol = rel.tol, inner.max.iter = inner.max.iter, trace = trace),
                 call = match.call()), class = "nnmf")
}

RunNMF <- function(A, k, alpha = 0, init = "ica", n.cores = 1, loss = "mse", max.iter = 500, ica.fast = F) {
  if (any(A < 0)) stop('The input matrix contains negative elements !')
  if (k < 3) stop("k must be greater than or equal to 3 to create a viable SWNE plot")
  if (!init %in% c("ica", "nnsvd", "random")) stop("Invalid initialization method")
  sparse <- methods::is(A, "dgCMatrix")        # new: no as.matrix() densify on the host (R/nmf.R:105)
  if (!sparse) A <- as.matrix(A)
  if (init == "ica") {
    nmf.init <- ica_init(as.matrix(A), k, ica.fast = ica.fast)
  } else if (init == "nnsvd") {
    nmf.init <- nnsvd_init(as.matrix(A), k, LINPACK = T)   # or the GPU rSVD: swne_nnsvd_init through a second shim
  } else {
    nmf.init <- NULL
  }
  if (!is.null(nmf.init)) {
    A.mean <- mean(A); zero.eps <- 1e-6
    nmf.init$W[nmf.init$W < zero.eps] <- 0; nmf.init$H[nmf.init$H < zero.eps] <- 0
    zero.idx.w <- which(nmf.init$W == 0); zero.idx.h <- which(nmf.init$H == 0)
    nmf.init$W[zero.idx.w] <- runif(length(zero.idx.w), 0, A.mean/100)
    nmf.init$H[zero.idx.h] <- runif(length(zero.idx.h), 0, A.mean/100)
  }
  nmf.res <- .nnmf_cuda(A, k = k, alpha = alpha, init = nmf.init, loss = loss, max.iter = max.iter)
  colnames(nmf.res$W) <- rownames(nmf.res$H) <- sapply(1:ncol(nmf.res$W), function(i) paste("factor", i, sep = "_"))
  return(nmf.res)
}
