## r/R/nmf_cuda.R -- drop-in overlay for R/nmf.R (RunNMF :96-134, FindNumFactors :157-205, ProjectFeatures :222-225,
## ProjectSamples :241-249, nnsvd_init :12-47, ica_init :56-66) and for the two hot helpers of R/swne_plotting.R
## (get_sample_coords :70-84, get_factor_coords :32-64 with its distance block on the device).  COMPILE-UNTESTED here (no R in the image).
## Same formals and return layouts; only the NNLM / ica / irlba calls are replaced by the CUDA path.  n.cores is accepted
## and ignored.  options(swne.cuda.device = 0) selects the GPU.  R's own RNG only provides one seed per call (the device
## generator is keyed by it), so set.seed() still makes runs reproducible.

.swne_dev <- function() as.integer(getOption("swne.cuda.device", 0L))
.swne_seed <- function() floor(runif(1, 1, 2^52))
.swne_mat <- function(A) if (methods::is(A, "dgCMatrix")) A else as.matrix(A)   # no as.matrix() densify of a dgCMatrix (R/nmf.R:105)
.swne_init_mode <- c(given = 0L, random = 1L, nnsvd = 2L, ica = 3L, ica_fast = 4L)

## NNLM::nnmf(A, k, alpha, beta, init, method = "scd", loss, max.iter, rel.tol, ...) on the GPU
.nnmf_cuda <- function(A, k, alpha = rep(0, 3), beta = rep(0, 3), init = NULL, loss = "mse", max.iter = 500,
                       rel.tol = 1e-4, inner.max.iter = ifelse(loss == "mse", 50, 1), inner.rel.tol = 1e-9,
                       trace = 100 / inner.max.iter, verbose = 0, init.mode = NULL) {
  alpha <- c(as.double(alpha), 0, 0, 0)[1:3]; beta <- c(as.double(beta), 0, 0, 0)[1:3]   # NNLM's padding
  if (is.null(init.mode)) init.mode <- if (is.null(init)) "random" else "given"          # NNLM's random start: U(0,1) * 0.01
  Wi <- if (is.null(init)) matrix(0, 0, 0) else init$W
  Hi <- if (is.null(init)) matrix(0, 0, 0) else init$H
  t0 <- proc.time()
  out <- swne_nmf_cuda(.swne_mat(A), k, Wi, Hi, alpha, beta, as.integer(max.iter), rel.tol, as.integer(inner.max.iter),
                       inner.rel.tol, ifelse(loss == "mse", 1L, 3L), as.integer(trace), .swne_init_mode[[init.mode]],
                       .swne_seed(), .swne_dev())
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
  ## the seed (ica_init / nnsvd_init / NNLM's random start) and the zero replacement of R/nmf.R:121-127 are computed on the
  ## device inside the same call as the fit
  mode <- if (init == "ica" && ica.fast) "ica_fast" else init
  nmf.res <- .nnmf_cuda(A, k = k, alpha = alpha, init = NULL, loss = loss, max.iter = max.iter, init.mode = mode)
  colnames(nmf.res$W) <- rownames(nmf.res$H) <- sapply(1:ncol(nmf.res$W), function(i) paste("factor", i, sep = "_"))
  return(nmf.res)
}

nnsvd_init <- function(A, k, LINPACK) {     # LINPACK kept for the signature (R/nmf.R:12), unused
  swne_seed_cuda(.swne_mat(A), as.integer(k), .swne_init_mode[["nnsvd"]], .swne_seed(), .swne_dev())
}

ica_init <- function(A, k, ica.fast = F) {
  swne_seed_cuda(.swne_mat(A), as.integer(k), .swne_init_mode[[if (ica.fast) "ica_fast" else "ica"]], .swne_seed(), .swne_dev())
}

FindNumFactors <- function(A, k.range = seq(2, 12, 2), n.cores = 1, do.plot = T, seed = NULL, loss = "mse", max.iter = 250) {
  if (!loss %in% c("mse", "mkl")) stop("Invalid loss function")
  if (ncol(A) > 20000) print("Warning: This function can be slow for very large datasets")
  if (!is.null(seed)) set.seed(seed)
  if (length(k.range) < 3) stop("k.range sequence must have at least 3 values")
  ## R/nmf.R:163-181 (the permuted matrix, the 2 x length(k.range) mse fits, their errors) in one device call
  k.err.mat <- swne_find_num_factors_cuda(.swne_mat(A), as.integer(k.range), ifelse(loss == "mse", 0L, 1L), as.integer(max.iter),
                                          .swne_seed(), .swne_dev())
  rownames(k.err.mat) <- c("Decomp", "RandomDecomp"); colnames(k.err.mat) <- k.range
  ## R/nmf.R:183-204, unchanged
  err.del <- sapply(1:(ncol(k.err.mat) - 1), function(i) k.err.mat[, i] - k.err.mat[, i + 1])
  colnames(err.del) <- colnames(k.err.mat)[-1]
  err.del.diff <- err.del[1, ] - err.del[2, ]
  if (sum(err.del.diff < 0) > 0) {
    min.idx <- which.max(err.del.diff < 0) - 1
  } else {
    min.idx <- which.min(err.del.diff)
  }
  res <- list()
  res$err <- err.del.diff
  res$k <- as.numeric(names(err.del.diff)[[min.idx]])
  if (do.plot) print(PlotFactorSelection(res, font.size = 12))
  return(res)
}

## NNLM::nnlm(x, y, alpha, method = "scd", loss) on the GPU: coefficients ncol(x) x ncol(y)
.nnlm_cuda <- function(A, X, side, alpha, loss) {
  alpha <- c(as.double(alpha), 0, 0, 0)[1:3]
  k <- if (side == 0L) ncol(X) else nrow(X)
  init <- if (side == 0L) matrix(runif(k * ncol(A)) * 0.01, k, ncol(A)) else matrix(runif(nrow(A) * k) * 0.01, nrow(A), k)
  swne_nnlm_cuda(.swne_mat(A), X, init, side, alpha, ifelse(loss == "mse", 0L, 1L), 10000L, 1e-12, .swne_dev())
}

ProjectFeatures <- function(newdata, H, alpha = rep(0, 3), loss = "mse", n.cores = 1) {   # R/nmf.R:222-225
  W <- .nnlm_cuda(newdata, H, 1L, alpha, loss)           # nnlm(t(H), t(newdata)) transposed: features x factors
  rownames(W) <- rownames(newdata); colnames(W) <- rownames(H)
  return(W)
}

ProjectSamples <- function(newdata, W, features.use = NULL, alpha = rep(0, 3), loss = "mse", n.cores = 1) {   # R/nmf.R:241-249
  if (is.null(features.use)) features.use <- intersect(rownames(newdata), rownames(W))
  H <- .nnlm_cuda(newdata[features.use, ], W[features.use, ], 0L, alpha, loss)
  rownames(H) <- colnames(W); colnames(H) <- colnames(newdata)
  return(H)
}

## get_sample_coords (R/swne_plotting.R:70-84): one device call instead of an sapply over all cells
get_sample_coords <- function(H, H.coords, alpha, n_pull) {
  if (n_pull < 3) { n_pull <- 3 }
  if (is.null(n_pull) || n_pull > nrow(H.coords)) { n_pull <- nrow(H.coords) }
  sample.coords <- swne_sample_coords_cuda(as.matrix(H), as.matrix(H.coords), alpha, as.integer(n_pull), .swne_dev())
  colnames(sample.coords) <- c("x", "y"); rownames(sample.coords) <- colnames(H)
  return(sample.coords)
}

## get_factor_coords (R/swne_plotting.R:32-64): the k x k factor distance over all cells on the device (cosine, pearson,
## euclidean); the mutual-information distance ("IC": randomised KDE, R/swne_plotting.R:44) keeps the original code; the
## Sammon mapping of the k x k matrix stays MASS::sammon
get_factor_coords <- function(H, method, distance, n.neighbors, min.dist) {
  distance <- tolower(distance)
  if (distance == "cor" || distance == "correlation") distance <- "pearson"
  if (distance == "mutual" || distance == "information" || distance == "mutual information" || distance == "ic") distance <- "IC"
  if (!distance %in% c("pearson", "IC", "cosine", "euclidean")) {
    stop(paste(c("Distance must be one of:", paste(c("pearson", "IC", "cosine", "euclidean"), collapse = ", ")), collapse = " "))
  }
  if (method == "sammon") {
    if (distance == "IC") {
      Ht <- t(H)
      H.dist <- sqrt(2 * (1 - as.matrix(usedist::dist_make(t(Ht), distance_fcn = MutualInf, method = "IC"))))
    } else {
      H.dist <- swne_factor_distance_cuda(as.matrix(H), c(cosine = 0L, pearson = 1L, euclidean = 2L)[[distance]], .swne_dev())
      dimnames(H.dist) <- list(rownames(H), rownames(H))
      if (distance != "pearson") H.dist <- stats::as.dist(H.dist)       # proxy::simil / proxy::dist return dist objects
    }
    H.coords <- MASS::sammon(H.dist, k = 2, niter = 250)$points
    rownames(H.coords) <- rownames(H)
  } else {
    stop("Invalid factor projection method")
  }
  H.coords <- apply(H.coords, 2, normalize_vector, method = "bounded")
  colnames(H.coords) <- c("x", "y")
  return(H.coords)
}
