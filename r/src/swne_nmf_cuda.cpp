// r/src/swne_nmf_cuda.cpp -- Rcpp shim a maintainer adds next to src/swne_cpp_funcs.cpp.
// COMPILE-UNTESTED in this repository's build container (no R toolchain there).
// It only marshals R objects to the C-ABI of libswne_nmf.so (include/swne_nmf.h); all numerics are on the GPU.
// Conventions follow src/RcppExports.cpp:12-21 ([[Rcpp::export]] -> .Call('_swne_<fn>', PACKAGE='swne', ...)),
// the dgCMatrix slot access follows src/swne_cpp_funcs.cpp:27-31, errors become Rcpp::stop / Rcpp::warning.
#include <Rcpp.h>
#include "swne_nmf.h"

using namespace Rcpp;

static void check(int status)
{
    if (status == SWNE_OK) return;
    if (status == SWNE_WARN_NOT_CONVERGED) { Rcpp::warning("Target tolerance not reached. Try a larger max.iter."); return; }
    Rcpp::stop(std::string("swne_nmf: ") + swne_nmf_last_error());
}

static int progress_cb(void *, int, double, double, double, double)
{
    try { Rcpp::checkUserInterrupt(); } catch (...) { return 1; }   // NNLM checks once per outer iteration
    return 0;
}

// one GPU handle with the matrix resident; destroyed on every exit path
struct Handle {
    swne_nmf_handle_t h = nullptr;
    int n = 0, m = 0;
    Handle(SEXP A, int device)
    {
        check(swne_nmf_create(&h, device));
        swne_nmf_set_progress(h, progress_cb, nullptr);
        try {
            if (Rf_isS4(A)) {                      // dgCMatrix: slots i, p, x, Dim (src/swne_cpp_funcs.cpp:27-31)
                S4 mat(A);
                IntegerVector dims = mat.slot("Dim"), ip = mat.slot("p"), ii = mat.slot("i");
                NumericVector xx = mat.slot("x");
                n = dims[0]; m = dims[1];
                check(swne_nmf_set_matrix_csc_f64(h, ip.begin(), ii.begin(), xx.begin(), n, m));
            } else {                               // dense double matrix, column-major genes x cells
                NumericMatrix Ad(A);
                n = Ad.nrow(); m = Ad.ncol();
                check(swne_nmf_set_matrix_dense_f64(h, Ad.begin(), n, m));
            }
        } catch (...) { swne_nmf_destroy(h); throw; }
    }
    ~Handle() { swne_nmf_destroy(h); }
};

static swne_nmf_params_t make_params(int k, NumericVector alpha, NumericVector beta, int max_iter, double rel_tol, int inner_max_iter,
                                     double inner_rel_tol, int method, int trace)
{
    swne_nmf_params_t p;
    swne_nmf_default_params(&p, k);
    for (int q = 0; q < 3; ++q) { p.alpha[q] = q < alpha.size() ? alpha[q] : 0.0; p.beta[q] = q < beta.size() ? beta[q] : 0.0; }
    p.loss = (method == 3) ? SWNE_LOSS_MKL : SWNE_LOSS_MSE;      // NNLM method codes 1 / 3
    p.max_iter = max_iter; p.rel_tol = rel_tol; p.inner_max_iter = inner_max_iter; p.inner_rel_tol = inner_rel_tol;
    p.trace = trace; p.full_traces = 1;                          // NNLM evaluates both losses at every trace point
    return p;
}

// NNLM::nnmf (R/nmf.R:117,129,167-168) and RunNMF's seeding in one call.
//   init_mode: 0 given (Wi / Hi), 1 random (NNLM's U(0,1) * 0.01), 2 nnsvd, 3 ica, 4 ica (ica.fast = T): 1-4 are computed on
//   the device together with the zero replacement of R/nmf.R:121-127; `seed` keys the device generator.
// [[Rcpp::export]]
List swne_nmf_cuda(SEXP A, int k, NumericMatrix Wi, NumericMatrix Hi, NumericVector alpha, NumericVector beta,
                   int max_iter, double rel_tol, int inner_max_iter, double inner_rel_tol, int method, int trace,
                   int init_mode, double seed, int device)
{
    Handle hd(A, device);
    swne_nmf_params_t p = make_params(k, alpha, beta, max_iter, rel_tol, inner_max_iter, inner_rel_tol, method, trace);
    const int tr = trace > 0 ? trace : std::max(1, 100 / (inner_max_iter > 0 ? inner_max_iter : (method == 3 ? 1 : 50)));
    const int len = (max_iter + tr - 1) / tr + 1;
    NumericMatrix W = (init_mode == SWNE_INIT_GIVEN) ? clone(Wi) : NumericMatrix(hd.n, k);
    NumericMatrix H = (init_mode == SWNE_INIT_GIVEN) ? clone(Hi) : NumericMatrix(k, hd.m);
    NumericVector mse(len), mkl(len), terr(len), epochs(len);
    swne_nmf_result_t res;
    check(swne_run_nmf(hd.h, &p, init_mode, (uint64_t)seed, W.begin(), H.begin(), mse.begin(), mkl.begin(), terr.begin(), epochs.begin(), len, &res));
    Range keep(0, res.n_trace - 1);
    return List::create(_["W"] = W, _["H"] = H, _["mse_error"] = mse[keep], _["mkl_error"] = mkl[keep],
                        _["target_error"] = terr[keep], _["average_epoch"] = epochs[keep], _["n_iteration"] = res.n_iteration);
}

// FindNumFactors' 2 x |k.range| fits and reconstruction errors (R/nmf.R:163-181) in one call; the permuted matrix is
// built on the device.  Returns the 2 x nk error matrix (row 1: A, row 2: permuted A).
// [[Rcpp::export]]
NumericMatrix swne_find_num_factors_cuda(SEXP A, IntegerVector k_range, int loss_metric, int max_iter, double seed, int device)
{
    Handle hd(A, device);
    const int nk = k_range.size();
    std::vector<double> kerr(2 * (size_t)nk);
    check(swne_find_num_factors(hd.h, k_range.begin(), nk, loss_metric, max_iter, (uint64_t)seed, SWNE_ENGINE_AUTO, kerr.data(), nullptr, nullptr));
    NumericMatrix out(2, nk);
    for (int c = 0; c < nk; ++c) { out(0, c) = kerr[c]; out(1, c) = kerr[nk + c]; }
    return out;
}

// NNLM::nnlm on the resident matrix (R/nmf.R:223, :247): side 0 solves H (k x m) given W = X (n x k) -- ProjectSamples;
// side 1 solves W (n x k) given H = X (k x m) -- ProjectFeatures.  Start: NNLM's U(0,1) * 0.01 drawn by R.
// [[Rcpp::export]]
NumericMatrix swne_nnlm_cuda(SEXP A, NumericMatrix X, NumericMatrix init, int side, NumericVector alpha, int loss, int max_iter,
                             double rel_tol, int device)
{
    Handle hd(A, device);
    const int k = (side == 0) ? X.ncol() : X.nrow();
    NumericMatrix coef = clone(init);
    double a3[3] = {0, 0, 0};
    for (int q = 0; q < 3 && q < alpha.size(); ++q) a3[q] = alpha[q];
    long long sweeps = 0;
    check(swne_nnlm(hd.h, side, k, X.begin(), coef.begin(), a3, loss, max_iter, rel_tol, &sweeps));
    return coef;
}

// nnsvd_init (R/nmf.R:12-47) / ica_init (R/nmf.R:56-66) as stand-alone calls (RunNMF itself uses swne_nmf_cuda's init_mode)
// [[Rcpp::export]]
List swne_seed_cuda(SEXP A, int k, int init_mode, double seed, int device)
{
    Handle hd(A, device);
    NumericMatrix W(hd.n, k), H(k, hd.m);
    if (init_mode == SWNE_INIT_NNSVD) {
        NumericVector sv(k);
        check(swne_nnsvd_init(hd.h, k, -1, -1, (uint64_t)seed, W.begin(), H.begin(), sv.begin()));
    } else {
        check(swne_ica_init(hd.h, k, init_mode == SWNE_INIT_ICA_FAST, (uint64_t)seed, W.begin(), H.begin()));
    }
    return List::create(_["W"] = W, _["H"] = H);
}

// get_sample_coords (R/swne_plotting.R:70-84)
// [[Rcpp::export]]
NumericMatrix swne_sample_coords_cuda(NumericMatrix H, NumericMatrix H_coords, double alpha, int n_pull, int device)
{
    swne_nmf_handle_t h = nullptr;
    check(swne_nmf_create(&h, device));
    NumericMatrix out(H.ncol(), 2);
    const int st = swne_sample_coords(h, H.begin(), H.nrow(), H.ncol(), H_coords.begin(), alpha, n_pull, out.begin());
    swne_nmf_destroy(h);
    check(st);
    return out;
}

// the factor-factor distance block of get_factor_coords (R/swne_plotting.R:43-52): 0 cosine, 1 pearson, 2 euclidean
// [[Rcpp::export]]
NumericMatrix swne_factor_distance_cuda(NumericMatrix H, int distance, int device)
{
    swne_nmf_handle_t h = nullptr;
    check(swne_nmf_create(&h, device));
    NumericMatrix out(H.nrow(), H.nrow());
    const int st = swne_factor_distance(h, H.begin(), H.nrow(), H.ncol(), distance, out.begin());
    swne_nmf_destroy(h);
    check(st);
    return out;
}
