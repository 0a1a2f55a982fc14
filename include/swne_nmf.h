/*
 * swne_nmf.h -- C-ABI of the B200-native NMF hot path of SWNE (libswne_nmf.so).
 *
 * This is the boundary a maintainer of yanwu2014/swne binds from the package's Rcpp layer
 * (src/swne_cpp_funcs.cpp + src/RcppExports.cpp) and that this repo binds with
 * ctypes (swne_b200/_capi.py).  Plain pointers and sizes only; no C++ or torch types; nothing throws
 * across the boundary.  Every entry point returns an int status (SWNE_OK == 0) and leaves a message for
 * swne_nmf_last_error().  All host matrices use R's memory layout (column-major):
 *     A : n x m (genes x cells)   A[i + n*j]
 *     W : n x k                   W[i + n*f]
 *     H : k x m                   H[f + k*j]
 * The library owns all device memory inside the opaque handle and keeps no host pointer after a call
 * returns.  One handle drives one GPU; multi-GPU runs use one process (or thread) per GPU, cells
 * (columns of A and H) sharded across ranks, see swne_nmf_comm_init.
 *
 * Reference interfaces replaced (file:line in /root/reference):
 *   swne_nmf_fit*            <- NNLM::nnmf(...) as called at R/nmf.R:117, R/nmf.R:129, R/nmf.R:167-168,
 *                               i.e. .Call("NNLM_c_nnmf", A, k, Wi, Hi, Wm, Hm, alpha, beta, max_iter,
 *                               rel_tol, n_threads, verbose, show_warning, inner_max_iter,
 *                               inner_rel_tol, method, trace)  
 *   swne_nmf_set_matrix_csc* <- the dgCMatrix slot access pattern of src/swne_cpp_funcs.cpp:27-31
 *   swne_nnlm                <- NNLM::nnlm(...) as called at R/nmf.R:223 and R/nmf.R:247
 *   swne_nnsvd_init          <- nnsvd_init, R/nmf.R:12-47 (svd call at :20)
 *   swne_zero_replace        <- R/nmf.R:121-127 (thresholding only; the runif draws stay with the caller)
 *   swne_sample_coords       <- get_sample_coords, R/swne_plotting.R:70-84
 *   swne_factor_distance     <- the k x k distance of get_factor_coords, R/swne_plotting.R:43-52
 *   swne_recon_error         <- the error block of FindNumFactors, R/nmf.R:170-179 (+ kl_div :71-75)
 */
#ifndef SWNE_NMF_H
#define SWNE_NMF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWNE_NMF_ABI_VERSION 2

/* status codes */
#define SWNE_OK 0
#define SWNE_WARN_NOT_CONVERGED 1 /* NNLM: "Target tolerance not reached. Try a larger max.iter." */
#define SWNE_ERR_BAD_ARG -1       /* negative / non-finite input, k out of range, masks, bad enum */
#define SWNE_ERR_CUDA -2
#define SWNE_ERR_NCCL -3
#define SWNE_ERR_NO_MATRIX -4     /* fit called before set_matrix */
#define SWNE_ERR_UNSUPPORTED -5
#define SWNE_ERR_INTERRUPTED -6   /* progress callback asked to stop (Rcpp::checkUserInterrupt) */

#define SWNE_LOSS_MSE 0 /* NNLM method code 1: scd + mse */
#define SWNE_LOSS_MKL 1 /* NNLM method code 3: scd + mkl */

/* which implementation of the two skinny contractions W'A and A H' the mse path uses */
#define SWNE_ENGINE_AUTO 0
#define SWNE_ENGINE_SIMT 1   /* fp32 CUDA-core kernels, three launches per half step */
#define SWNE_ENGINE_TC 2     /* fused one-pass tcgen05 kernel on split-fp16 (fp32-emulating) operands */

#define SWNE_MAX_K 64

typedef struct swne_nmf_handle_s *swne_nmf_handle_t;

typedef struct swne_nmf_params {
    int k;                 /* number of factors, 1..SWNE_MAX_K (RunNMF itself enforces k >= 3) */
    double alpha[3];       /* penalties on W: [L2, angle, L1]  (RunNMF's scalar alpha -> alpha[0]) */
    double beta[3];        /* penalties on H: [L2, angle, L1] */
    int loss;              /* SWNE_LOSS_MSE | SWNE_LOSS_MKL */
    int max_iter;          /* 500 RunNMF, 250 FindNumFactors */
    double rel_tol;        /* 1e-4; a negative value disables the convergence test (fixed #iterations) */
    int inner_max_iter;    /* <=0 -> NNLM default: 50 (mse) / 1 (mkl) */
    double inner_rel_tol;  /* <=0 -> 1e-9 */
    int trace;             /* <=0 -> NNLM default 100/inner_max_iter; error block every `trace` iterations */
    int verbose;           /* 0 silent, 2 prints the per-trace table to stderr */
    int engine;            /* SWNE_ENGINE_* */
    int full_traces;       /* the loss that is not being optimised (mkl in an mse fit) is evaluated
                              1: at every trace point, as NNLM does; 0: only in the final error block;
                              -1: never (NaN in the mkl trace) */
    int profile;           /* 1: bracket every kernel family with CUDA events (per-family times in the result) */
    int nnlm_variant;      /* SWNE_VARIANT_* bits: readings of NNLM details that no golden vector pins yet (0 = as
                              recalled in SURVEY.md Appendix A) */
} swne_nmf_params_t;

/* scd_kl_update with an L2 penalty: Appendix A recalls "a += beta0; b += a*h - ..." (the default); this bit selects
   "b += a*h - ...; a += beta0" (the exact Newton step).  Only matters for loss = mkl with alpha[0] / beta[0] != 0. */
#define SWNE_VARIANT_KL_L2_LATE 1

#define SWNE_PROF_SLOTS 8
typedef struct swne_nmf_result {
    int n_iteration;       /* outer iterations performed */
    int n_trace;           /* number of valid entries in the trace arrays */
    int not_converged;     /* 1 when |rel_err| > rel_tol at exit */
    int engine_used;       /* SWNE_ENGINE_SIMT | SWNE_ENGINE_TC */
    double loop_ms;        /* CUDA-event time of the iteration loop on the library's stream */
    double upload_ms;      /* host->device time of W0/H0 (and A for the one-shot calls) */
    double download_ms;    /* device->host time of W/H */
    long long gpu_launches;/* kernels launched inside the loop */
    /* profile==1: per kernel family summed event time and launch count.
       slot 0 pass/contract-H, 1 solve-H (mkl: the cell side's KL update), 2 contract-W, 3 solve-W, 4 gram/loss,
       5 mkl: the gene side's KL update, 6 allreduce */
    double prof_ms[SWNE_PROF_SLOTS];
    long long prof_launches[SWNE_PROF_SLOTS];
} swne_nmf_result_t;

/* called on the calling thread after every error block; return non-zero to stop (user interrupt) */
typedef int (*swne_nmf_progress_fn)(void *user, int iteration, double mse, double mkl, double target,
                                    double rel_err);

const char *swne_nmf_last_error(void);
int swne_nmf_abi_version(void);
int swne_nmf_device_count(int *count);
void swne_nmf_default_params(swne_nmf_params_t *p, int k);

int swne_nmf_create(swne_nmf_handle_t *out, int device);
int swne_nmf_destroy(swne_nmf_handle_t h);
int swne_nmf_set_progress(swne_nmf_handle_t h, swne_nmf_progress_fn fn, void *user);

/* ---- multi-GPU: cells sharded over ranks.  loss = mse: one exchange of {A H' (n x k), H H' (k x k), loss partials} per
 * outer iteration (NVLink peer memory or ncclAllReduce, see swne_nmf_comm_info).  loss = mkl: the cell side is local, the
 * gene side takes one ncclAllReduce of 2 n floats per coordinate (k per sweep) plus H H' / row sums per iteration.  W is
 * replicated bit for bit on all ranks.  id is an ncclUniqueId (128 bytes) made by rank 0 and sent to the other ranks by
 * the caller (torch.distributed / MPI / R parallel). */
int swne_nmf_comm_unique_id(void *id128);
int swne_nmf_comm_init(swne_nmf_handle_t h, const void *id128, int rank, int nranks);
/* rank / number of ranks of the handle, and whether the per-iteration sums of the fused engine travel over NVLink peer
 * memory (1: CUDA IPC mappings of every rank's exchange region, set up at the first sharded fit; 0: NCCL all-reduces --
 * no peer access, SWNE_PEER_XCHG=0, or no sharded fit yet) */
int swne_nmf_comm_info(swne_nmf_handle_t h, int *rank, int *nranks, int *peer_exchange);

/* ---- the data matrix (this rank's cells).  m is the LOCAL number of cells. Values must be finite and
 * >= 0 (RunNMF R/nmf.R:98,106). The matrix stays resident until replaced, so several fits
 * (FindNumFactors) reuse one upload. */
int swne_nmf_set_matrix_dense_f64(swne_nmf_handle_t h, const double *A, int n, int m);
int swne_nmf_set_matrix_dense_f32(swne_nmf_handle_t h, const float *A, int n, int m);
/* dgCMatrix slots: p (m+1), i (nnz, 0-based row = gene), x (nnz) */
int swne_nmf_set_matrix_csc_f64(swne_nmf_handle_t h, const int *p, const int *i, const double *x, int n, int m);
int swne_nmf_set_matrix_csc_f32(swne_nmf_handle_t h, const int *p, const int *i, const float *x, int n, int m);
/* Raw counts as dgCMatrix slots with ScaleCounts (R/data_processing.R:275-307) fused into the upload, so the normalised
 * matrix never exists on the host (R/nmf.R:105 would make it dense there):
 *   a = f(x * cell_scale[cell] * gene_scale[gene])   on the stored entries, zeros stay zero (the reference transforms @x)
 *   cell_scale[j] = depthScale / depth[j]  (NormalizeCounts, R/data_processing.R:167,184; depth = column sums of the FULL
 *                   count matrix, depthScale = their median by default, :278-280), NULL = 1
 *   gene_scale[g] = the variance scale factor gsf of AdjustVariance (:284-287; its GAM fit stays on the host), NULL = 1
 *   transform     = SWNE_TRANSFORM_NONE | _LOG (log(x + 1), :301) | _FT (sqrt(x) + sqrt(x + 1), :256,:299)
 * The batch correction of NormalizeCounts and method = "logrank" are not covered.  (i, j) pairs must be unique. */
#define SWNE_TRANSFORM_NONE 0
#define SWNE_TRANSFORM_LOG 1
#define SWNE_TRANSFORM_FT 2
int swne_nmf_set_matrix_csc_counts(swne_nmf_handle_t h, const int *p, const int *i, const double *x, int n, int m,
                                   const double *cell_scale, const double *gene_scale, int transform);
/* A already in HBM (fp32, cell-major = R column-major, leading dimension n): no copy is made on the
 * host; used by bench.py for the HBM-resident measurement. */
int swne_nmf_set_matrix_dense_f32_device(swne_nmf_handle_t h, const float *dA, int n, int m);
/* shape, nonzero count, sum and sum of squares of the resident matrix */
int swne_nmf_matrix_info(swne_nmf_handle_t h, int *n, int *m, long long *nnz, double *sum, double *sumsq);

/* ---- the fit.  W (n x k) and H (k x m_local) hold the init on entry (NNLM's init=list(W,H)) and the
 * result on return.  Trace arrays hold trace_len >= ceil(max_iter/trace)+1 doubles (may be NULL). */
int swne_nmf_fit(swne_nmf_handle_t h, const swne_nmf_params_t *p, double *W, double *H,
                 double *mse, double *mkl, double *target_loss, double *average_epochs, int trace_len,
                 swne_nmf_result_t *res);
/* one-shot forms with the NNLM_c_nnmf argument list (A uploaded inside the call) */
int swne_nmf_fit_dense_f64(swne_nmf_handle_t h, const swne_nmf_params_t *p, const double *A, int n, int m,
                           double *W, double *H, double *mse, double *mkl, double *target_loss,
                           double *average_epochs, int trace_len, swne_nmf_result_t *res);
int swne_nmf_fit_csc_f64(swne_nmf_handle_t h, const swne_nmf_params_t *p, const int *cp, const int *ci,
                         const double *cx, int n, int m, double *W, double *H, double *mse, double *mkl,
                         double *target_loss, double *average_epochs, int trace_len, swne_nmf_result_t *res);

/* ---- RunNMF in one call on the resident matrix (R/nmf.R:96-134): seed -> zero replacement (R/nmf.R:121-127) -> nnmf.
 * The seed is computed on the device and handed to the fit without a host round trip; W (n x k) and H (k x m) are
 * outputs, except for SWNE_INIT_GIVEN where they hold the init (then this is swne_nmf_fit).  Random draws come from a
 * counter-based generator keyed by `seed` (R's RNG streams cannot be reproduced by any other program).
 *   RANDOM    NNLM's own start, U(0,1) * 0.01            (RunNMF(init = "random"), FindNumFactors)
 *   NNSVD     nnsvd_init (R/nmf.R:12-47) from a randomized truncated SVD
 *   ICA       ica_init(ica.fast = F) (R/nmf.R:63-64): FastICA on the top-k principal subspace of the centred data
 *   ICA_FAST  ica_init(ica.fast = T) (R/nmf.R:58-61): 50 principal components, W = (A - rowMeans) S */
#define SWNE_INIT_GIVEN 0
#define SWNE_INIT_RANDOM 1
#define SWNE_INIT_NNSVD 2
#define SWNE_INIT_ICA 3
#define SWNE_INIT_ICA_FAST 4
int swne_run_nmf(swne_nmf_handle_t h, const swne_nmf_params_t *p, int init, uint64_t seed, double *W, double *H,
                 double *mse, double *mkl, double *target_loss, double *average_epochs, int trace_len,
                 swne_nmf_result_t *res);

/* ---- FindNumFactors (R/nmf.R:157-205) on the resident matrix in one call: for every k in k_range two mse fits from
 * NNLM's random start (max_iter, rel_tol 1e-4) -- on A and on a random permutation of all its entries (R/nmf.R:163) --
 * and their reconstruction errors (loss_metric SWNE_LOSS_MSE: mean squared error, SWNE_LOSS_MKL: mean kl_div,
 * R/nmf.R:71-75,170-181).  k_err: 2 x nk row-major (row 0: A, row 1: permuted); err_diff (nk - 1, may be NULL) and
 * k_pick (may be NULL) follow R/nmf.R:186-194.  The permuted matrix, seeds, W and H never leave the device.  The
 * permutation is a function of `seed`, every fit's start of (seed, k): a replica (one GPU of several) may sweep any subset
 * of k.range -- with k_pick and err_diff NULL fewer than 3 values are accepted -- and obtains the errors of the full sweep. */
int swne_find_num_factors(swne_nmf_handle_t h, const int *k_range, int nk, int loss_metric, int max_iter, uint64_t seed,
                          int engine, double *k_err, int *k_pick, double *err_diff);

/* ---- NNLM::nnlm on the resident matrix: side 0 solves H (k x m) given W (ProjectSamples),
 * side 1 solves W (n x k) given H (ProjectFeatures on the fitted genes).  coef holds the init. */
int swne_nnlm(swne_nmf_handle_t h, int side, int k, const double *X, double *coef, const double *alpha3,
              int loss, int max_iter, double rel_tol, long long *sweeps);

/* ---- seeding */
/* NNDSVD of the resident matrix from a randomized truncated SVD (oversample p, q power iterations). */
int swne_nnsvd_init(swne_nmf_handle_t h, int k, int oversample, int power_iters, uint64_t seed,
                    double *W, double *H, double *singular_values);
/* ica_init (R/nmf.R:56-66) of the resident matrix: signed W (n x k) and H (k x m), components up to sign and order */
int swne_ica_init(swne_nmf_handle_t h, int k, int ica_fast, uint64_t seed, double *W, double *H);
/* x < eps -> 0 in place; returns the number of zeros so the caller can draw its runif()s */
int swne_zero_replace(double *x, size_t len, double eps, const double *fill, size_t n_fill, size_t *n_zero);

/* ---- consumers of H.  H = NULL: the H the most recent swne_nmf_fit / swne_run_nmf of this handle left on the device
 * ("while H is still resident"; k and m must be that fit's; any other call that works on the handle's
 * factors -- nnlm, recon_error, a seed, set_matrix -- ends the residency).  Identical results to passing the returned H. */
int swne_sample_coords(swne_nmf_handle_t h, const double *H, int k, int m, const double *H_coords /* k x 2 col-major */,
                       double alpha_exp, int n_pull, double *out /* m x 2 col-major */);
/* distance: 0 cosine (1 - cos), 1 pearson sqrt(2(1-r)), 2 euclidean; out k x k */
int swne_factor_distance(swne_nmf_handle_t h, const double *H, int k, int m, int distance, double *out);
/* mean((W H - A)^2) and mean(kl_div(W H, A)) on the resident matrix (FindNumFactors error block) */
int swne_recon_error(swne_nmf_handle_t h, const double *W, const double *H, int k, double *mse, double *kl);

#ifdef __cplusplus
}
#endif
#endif /* SWNE_NMF_H */
