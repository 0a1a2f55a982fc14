/*
 * oracle/nnmf_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (fp64, plain C + OpenMP) of the algorithm behind
 *   RunNMF  -> NNLM::nnmf   (/root/reference/R/nmf.R:96-134, call sites :117, :129, :167-168)
 *   ProjectFeatures/ProjectSamples -> NNLM::nnlm (/root/reference/R/nmf.R:222-249)
 *
 * The arithmetic lives in the un-vendored, unpinned dependency linxihui/NNLM (0.4.x line;
 * /root/reference/DESCRIPTION:42-43, NAMESPACE:61).  Its source is not in the reference tree, so this
 * file restates NNLM's published sequential-coordinate-descent algorithm exactly as written down in
 * SURVEY.md Appendix A (c_nnmf / update / scd_ls_update / scd_kl_update / add_penalty).
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors for this path and neither
 * R nor NNLM can run in the build container.  The restatement is pinned only by independent
 * cross-checks (scipy.optimize.nnls KKT agreement, monotone target loss, direct-vs-Gram mse) in
 * tests/test_oracle.py; tools/make_golden.R lets a maintainer with R + NNLM produce true goldens.
 *
 * Two details of Appendix A are recalled, not verified, and both are switchable so that a true NNLM golden
 * (tools/make_golden.R) can settle them without a code change -- tests assert the GPU path matches this oracle under
 * BOTH settings of each:
 *   1. the default of `trace` (error block every `trace` iterations, which also sets when the stopping rule is
 *      tested): 100 / inner_max_iter in newer NNLM, 10 in 0.4.1.  `trace` is an explicit argument here and in the
 *      C-ABI; the callers' default is the newer rule.
 *   2. scd_kl_update with an L2 penalty: "a += beta0; b += a*h - ..." (default, as recalled) against
 *      "b += a*h - ...; a += beta0" (oracle_set_variant(1); C-ABI: SWNE_VARIANT_KL_L2_LATE).  Only loss = mkl with
 *      alpha[0] / beta[0] != 0 can tell them apart.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product path (swne_b200/csrc) never links or calls it.
 *
 * Memory layouts are R's (column-major doubles):
 *   A : n x m  -> A[i + n*j]   (gene i, cell j)     W : n x k -> W[i + n*f]     H : k x m -> H[f + k*j]
 * Internally NNLM works on Wt = t(W) (k x n), restated here the same way.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TINY_NUM 1e-16

/* scd_ls_update: Gauss-Seidel NNLS on one column.
 * h[k], mu[k] in/out, G k x k (symmetric, column kk contiguous at G + kk*k). returns #sweeps */
static int scd_ls_update(double *h, const double *G, double *mu, int k, int max_iter, double rel_tol)
{
    double rel_err = 1.0 + rel_tol;
    int t = 0;
    for (; t < max_iter && rel_err > rel_tol; ++t) {
        rel_err = 0.0;
        for (int kk = 0; kk < k; ++kk) {
            double tmp = h[kk] - mu[kk] / G[kk + (size_t)kk * k];
            if (tmp < 0.0) tmp = 0.0;
            if (tmp != h[kk]) {
                const double d = tmp - h[kk];
                const double *g = G + (size_t)kk * k;
                for (int r = 0; r < k; ++r) mu[r] += d * g[r];
                const double e = 2.0 * fabs(h[kk] - tmp) / (tmp + h[kk] + TINY_NUM);
                if (e > rel_err) rel_err = e;
                h[kk] = tmp;
            }
        }
    }
    return t;
}

/* scd_kl_update: one column, KL loss.
 * Wt k x r (column i = Wt + i*k), a r-vector (the data column), sumW k-vector, scratch Ajt r-vector */
/* switch for a recalled-but-unverified detail of NNLM (see the header comment); set between fits, read-only inside them */
static int g_kl_l2_late = 0;
void oracle_set_variant(int kl_l2_late) { g_kl_l2_late = kl_l2_late; }

static int scd_kl_update(double *h, const double *Wt, const double *a, const double *sumW,
                         const double *beta, int k, int r, int max_iter, double rel_tol, double *Ajt)
{
    double sumH = 0.0;
    for (int f = 0; f < k; ++f) sumH += h[f];
    for (int i = 0; i < r; ++i) {
        const double *w = Wt + (size_t)i * k;
        double s = 0.0;
        for (int f = 0; f < k; ++f) s += w[f] * h[f];
        Ajt[i] = s;
    }
    double rel_err = 1.0 + rel_tol;
    int t = 0;
    for (; t < max_iter && rel_err > rel_tol; ++t) {
        rel_err = 0.0;
        for (int kk = 0; kk < k; ++kk) {
            double a2 = 0.0, b = 0.0;
            for (int i = 0; i < r; ++i) {
                const double mu = Wt[kk + (size_t)i * k] / (Ajt[i] + TINY_NUM);
                a2 += a[i] * mu * mu;
                b += a[i] * mu;
            }
            b -= sumW[kk];
            /* [NNLM-mem, uncertain] as recalled: "a += beta0; b += a*h - ...".  g_kl_l2_late selects the other reading
               (b += a*h before a += beta0, the exact Newton step); they differ only when beta[0] != 0. */
            double a2h;
            if (g_kl_l2_late) { a2h = a2 * h[kk]; a2 += beta[0]; }
            else { a2 += beta[0]; a2h = a2 * h[kk]; }
            b += a2h - beta[2] - beta[1] * (sumH - h[kk]);
            double tmp = b / (a2 + TINY_NUM);
            if (tmp < 0.0) tmp = 0.0;
            if (tmp != h[kk]) {
                const double d = tmp - h[kk];
                for (int i = 0; i < r; ++i) Ajt[i] += d * Wt[kk + (size_t)i * k];
                const double e = 2.0 * fabs(h[kk] - tmp) / (tmp + h[kk] + TINY_NUM);
                if (e > rel_err) rel_err = e;
                sumH += d;
                h[kk] = tmp;
            }
        }
    }
    return t;
}

/* update: solves A ~= Wt' * H for H (k x c), Wt k x r, A r x c column-major.
 * method 1 = scd+mse, 3 = scd+mkl.  returns total #sweeps over all columns. */
long nnlm_update_ref(double *H, const double *Wt, const double *A, int k, int r, int c,
                     const double *beta, int max_iter, double rel_tol, int n_threads, int method)
{
    double *G = NULL, *sumW = NULL;
    if (method == 1) {
        G = (double *)calloc((size_t)k * k, sizeof(double));
        for (int i = 0; i < r; ++i) {
            const double *w = Wt + (size_t)i * k;
            for (int q = 0; q < k; ++q)
                for (int p = 0; p < k; ++p) G[p + (size_t)q * k] += w[p] * w[q];
        }
        if (beta[0] != beta[1])
            for (int p = 0; p < k; ++p) G[p + (size_t)p * k] += beta[0] - beta[1];
        if (beta[1] != 0.0)
            for (int p = 0; p < k * k; ++p) G[p] += beta[1];
        for (int p = 0; p < k; ++p) G[p + (size_t)p * k] += TINY_NUM;
    } else {
        sumW = (double *)calloc((size_t)k, sizeof(double));
        for (int i = 0; i < r; ++i)
            for (int f = 0; f < k; ++f) sumW[f] += Wt[f + (size_t)i * k];
    }
    long total = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(+ : total)
    {
        double *mu = (double *)malloc(sizeof(double) * (size_t)k);
        double *Ajt = (method == 3) ? (double *)malloc(sizeof(double) * (size_t)r) : NULL;
#pragma omp for schedule(dynamic, 16)
        for (int j = 0; j < c; ++j) {
            double *h = H + (size_t)j * k;
            const double *a = A + (size_t)j * r;
            if (method == 1) {
                /* mu = G h - Wt a (+ L1) */
                for (int p = 0; p < k; ++p) {
                    double s = 0.0;
                    for (int q = 0; q < k; ++q) s += G[p + (size_t)q * k] * h[q];
                    mu[p] = s;
                }
                for (int i = 0; i < r; ++i) {
                    const double ai = a[i];
                    if (ai == 0.0) continue;
                    const double *w = Wt + (size_t)i * k;
                    for (int p = 0; p < k; ++p) mu[p] -= w[p] * ai;
                }
                if (beta[2] != 0.0)
                    for (int p = 0; p < k; ++p) mu[p] += beta[2];
                total += scd_ls_update(h, G, mu, k, max_iter, rel_tol);
            } else {
                total += scd_kl_update(h, Wt, a, sumW, beta, k, r, max_iter, rel_tol, Ajt);
            }
        }
        free(mu);
        free(Ajt);
    }
    free(G);
    free(sumW);
    return total;
}

/* error block of c_nnmf + add_penalty.  Wt k x n, H k x m. */
static void error_block(const double *A, const double *Wt, const double *H, int n, int m, int k,
                        const double *alpha, const double *beta, int method, double const_kl,
                        double *mse_out, double *mkl_out, double *terr_out)
{
    const double N = (double)n * (double)m;
    double sse = 0.0, skl = 0.0;
#pragma omp parallel for reduction(+ : sse, skl) schedule(static)
    for (int j = 0; j < m; ++j) {
        const double *h = H + (size_t)j * k;
        const double *a = A + (size_t)j * n;
        for (int i = 0; i < n; ++i) {
            const double *w = Wt + (size_t)i * k;
            double ah = 0.0;
            for (int f = 0; f < k; ++f) ah += w[f] * h[f];
            const double d = a[i] - ah;
            sse += d * d;
            skl += -(a[i] + TINY_NUM) * log(ah + TINY_NUM) + ah;
        }
    }
    *mse_out = sse / N;
    *mkl_out = const_kl + skl / N;
    double terr = (method < 3) ? 0.5 * (*mse_out) : *mkl_out;
    /* add_penalty */
    double w2 = 0, wg = 0, w1 = 0, h2 = 0, hg = 0, h1 = 0;
    {
        double *s = (double *)calloc((size_t)k, sizeof(double));
        for (int i = 0; i < n; ++i)
            for (int f = 0; f < k; ++f) {
                const double v = Wt[f + (size_t)i * k];
                w2 += v * v; w1 += v; s[f] += v;
            }
        /* sum of all entries of the k x k Gram Wt*Wt' = sum_i (sum_f w_if)^2 */
        for (int i = 0; i < n; ++i) {
            double r = 0; for (int f = 0; f < k; ++f) r += Wt[f + (size_t)i * k];
            wg += r * r;
        }
        memset(s, 0, sizeof(double) * (size_t)k);
        for (int j = 0; j < m; ++j) {
            double r = 0;
            for (int f = 0; f < k; ++f) {
                const double v = H[f + (size_t)j * k];
                h2 += v * v; h1 += v; r += v;
            }
            hg += r * r;
        }
        free(s);
    }
    if (alpha[0] != alpha[1]) terr += 0.5 * (alpha[0] - alpha[1]) * w2 / N;
    if (beta[0] != beta[1]) terr += 0.5 * (beta[0] - beta[1]) * h2 / N;
    if (alpha[1] != 0.0) terr += 0.5 * alpha[1] * wg / N;
    if (beta[1] != 0.0) terr += 0.5 * beta[1] * hg / N;
    if (alpha[2] != 0.0) terr += alpha[2] * w1 / N;
    if (beta[2] != 0.0) terr += beta[2] * h1 / N;
    *terr_out = terr;
}

/* c_nnmf.  W (n x k) and H (k x m) are in/out, R layout.  Trace arrays must hold
 * ceil(max_iter/trace)+1 doubles.  method: 1 = scd+mse, 3 = scd+mkl.
 * returns 0, or 1 when the tolerance was not reached ("Target tolerance not reached" warning). */
int nnmf_ref(const double *A, int n, int m, int k, double *W, double *H,
             const double *alpha, const double *beta, int max_iter, double rel_tol,
             int inner_max_iter, double inner_rel_tol, int method, int trace, int n_threads,
             double *mse_err, double *mkl_err, double *terr, double *ave_epoch,
             int *n_iteration, int *n_trace)
{
    const double N = (double)n * (double)m;
    if (trace < 1) trace = 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    double const_kl = 0.0;
    {
        double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
        for (long q = 0; q < (long)n * m; ++q) s += (A[q] + TINY_NUM) * log(A[q] + TINY_NUM) - A[q];
        const_kl = s / N;
    }
    double *Wt = (double *)malloc(sizeof(double) * (size_t)k * n);
    double *At = (double *)malloc(sizeof(double) * (size_t)n * m);
    for (int i = 0; i < n; ++i)
        for (int f = 0; f < k; ++f) Wt[f + (size_t)i * k] = W[i + (size_t)f * n];

    double rel_err = rel_tol + 1.0, terr_last = 1e99;
    int i_e = 0, i = 0;
    long total_raw_iter = 0;
    for (; i < max_iter && fabs(rel_err) > rel_tol; ++i) {
        /* NNLM materialises A.t() every outer iteration; kept so the cost profile matches */
#pragma omp parallel for schedule(static)
        for (int g = 0; g < n; ++g)
            for (int j = 0; j < m; ++j) At[j + (size_t)g * m] = A[g + (size_t)j * n];
        total_raw_iter += nnlm_update_ref(Wt, H, At, k, m, n, alpha, inner_max_iter, inner_rel_tol, n_threads, method);
        total_raw_iter += nnlm_update_ref(H, Wt, A, k, n, m, beta, inner_max_iter, inner_rel_tol, n_threads, method);
        if (i % trace == 0) {
            error_block(A, Wt, H, n, m, k, alpha, beta, method, const_kl, &mse_err[i_e], &mkl_err[i_e], &terr[i_e]);
            ave_epoch[i_e] = (double)total_raw_iter / (double)(n + m);
            rel_err = 2.0 * (terr_last - terr[i_e]) / (terr_last + terr[i_e] + TINY_NUM);
            terr_last = terr[i_e];
            total_raw_iter = 0;
            ++i_e;
        }
    }
    if ((i - 1) % trace != 0 && i > 0) {
        error_block(A, Wt, H, n, m, k, alpha, beta, method, const_kl, &mse_err[i_e], &mkl_err[i_e], &terr[i_e]);
        ave_epoch[i_e] = (double)total_raw_iter / (double)(n + m);
        rel_err = 2.0 * (terr_last - terr[i_e]) / (terr_last + terr[i_e] + TINY_NUM);
        ++i_e;
    }
    for (int g = 0; g < n; ++g)
        for (int f = 0; f < k; ++f) W[g + (size_t)f * n] = Wt[f + (size_t)g * k];
    free(Wt);
    free(At);
    *n_iteration = i;
    *n_trace = i_e;
    return (rel_err > rel_tol) ? 1 : 0;
}

/* nnlm: one update() call: coefficients (k x c) from x (r x k), y (r x c).
 * xt is t(x) = k x r.  coef in/out (init).  returns total sweeps. */
long nnlm_ref(double *coef, const double *xt, const double *y, int k, int r, int c,
              const double *alpha, int max_iter, double rel_tol, int n_threads, int method)
{
    return nnlm_update_ref(coef, xt, y, k, r, c, alpha, max_iter, rel_tol, n_threads, method);
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
