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

// [[Rcpp::export]]
List swne_nmf_cuda(SEXP A, int k, NumericMatrix Wi, NumericMatrix Hi, NumericVector alpha, NumericVector beta,
                   int max_iter, double rel_tol, int inner_max_iter, double inner_rel_tol, int method, int trace,
                   int device)
{
    swne_nmf_handle_t h = nullptr;
    check(swne_nmf_create(&h, device));
    swne_nmf_set_progress(h, progress_cb, nullptr);
    int n = 0, m = 0;
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
        swne_nmf_params_t p;
        swne_nmf_default_params(&p, k);
        for (int q = 0; q < 3; ++q) { p.alpha[q] = q < alpha.size() ? alpha[q] : 0.0; p.beta[q] = q < beta.size() ? beta[q] : 0.0; }
        p.loss = (method == 3) ? SWNE_LOSS_MKL : SWNE_LOSS_MSE;      // NNLM method codes 1 / 3
        p.max_iter = max_iter; p.rel_tol = rel_tol; p.inner_max_iter = inner_max_iter; p.inner_rel_tol = inner_rel_tol;
        p.trace = trace; p.full_traces = 1;
        const int tr = trace > 0 ? trace : std::max(1, 100 / (inner_max_iter > 0 ? inner_max_iter : (method == 3 ? 1 : 50)));
        const int len = (max_iter + tr - 1) / tr + 1;
        NumericMatrix W = clone(Wi), H = clone(Hi);
        NumericVector mse(len), mkl(len), terr(len), epochs(len);
        swne_nmf_result_t res;
        check(swne_nmf_fit(h, &p, W.begin(), H.begin(), mse.begin(), mkl.begin(), terr.begin(), epochs.begin(), len, &res));
        swne_nmf_destroy(h);
        Range keep(0, res.n_trace - 1);
        return List::create(_["W"] = W, _["H"] = H, _["mse_error"] = mse[keep], _["mkl_error"] = mkl[keep],
                            _["target_error"] = terr[keep], _["average_epoch"] = epochs[keep], _["n_iteration"] = res.n_iteration);
    } catch (...) {
        swne_nmf_destroy(h);
        throw;
    }
}

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
