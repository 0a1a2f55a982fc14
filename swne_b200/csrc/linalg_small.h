// linalg_small.h -- host-side l x l (l <= 64) dense linear algebra used between GPU contractions of the
// randomized SVD (Cholesky QR, eigen-decomposition of the small Gram).  Row-major doubles.
#pragma once
#include <math.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace swne {

// G = R'R, R upper triangular.  false when G is not numerically positive definite.
static inline bool small_cholesky_upper(const double *G, int l, double *R)
{
    std::fill(R, R + l * l, 0.0);
    for (int j = 0; j < l; ++j) {
        double d = G[j * l + j];
        for (int q = 0; q < j; ++q) d -= R[q * l + j] * R[q * l + j];
        if (!(d > 0.0) || !(d > 1e-14 * fabs(G[j * l + j]))) return false;
        const double rjj = sqrt(d);
        R[j * l + j] = rjj;
        for (int c = j + 1; c < l; ++c) {
            double s = G[j * l + c];
            for (int q = 0; q < j; ++q) s -= R[q * l + j] * R[q * l + c];
            R[j * l + c] = s / rjj;
        }
    }
    return true;
}

// inverse of an upper triangular matrix
static inline void small_upper_inverse(const double *R, int l, double *Ri)
{
    std::fill(Ri, Ri + l * l, 0.0);
    for (int j = 0; j < l; ++j) {
        Ri[j * l + j] = 1.0 / R[j * l + j];
        for (int i = j - 1; i >= 0; --i) {
            double s = 0.0;
            for (int q = i + 1; q <= j; ++q) s += R[i * l + q] * Ri[q * l + j];
            Ri[i * l + j] = -s / R[i * l + i];
        }
    }
}

// cyclic Jacobi eigen-decomposition of a symmetric matrix; eigenvalues descending, eigenvectors in
// the COLUMNS of V (V[a*l + b] = component a of eigenvector b).
static inline void small_jacobi_eigh(const double *S, int l, double *eval, double *V)
{
    std::vector<double> A(S, S + l * l), Q(l * l, 0.0);
    for (int i = 0; i < l; ++i) Q[i * l + i] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < l; ++i)
            for (int j = 0; j < l; ++j) (i == j ? diag : off) += A[i * l + j] * A[i * l + j];
        if (off <= 1e-30 * diag) break;
        for (int p = 0; p < l - 1; ++p)
            for (int q = p + 1; q < l; ++q) {
                const double apq = A[p * l + q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (A[q * l + q] - A[p * l + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int r = 0; r < l; ++r) {
                    const double arp = A[r * l + p], arq = A[r * l + q];
                    A[r * l + p] = c * arp - s * arq;
                    A[r * l + q] = s * arp + c * arq;
                }
                for (int r = 0; r < l; ++r) {
                    const double apr = A[p * l + r], aqr = A[q * l + r];
                    A[p * l + r] = c * apr - s * aqr;
                    A[q * l + r] = s * apr + c * aqr;
                }
                for (int r = 0; r < l; ++r) {
                    const double qrp = Q[r * l + p], qrq = Q[r * l + q];
                    Q[r * l + p] = c * qrp - s * qrq;
                    Q[r * l + q] = s * qrp + c * qrq;
                }
            }
    }
    std::vector<int> order(l);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return A[a * l + a] > A[b * l + b]; });
    for (int b = 0; b < l; ++b) {
        eval[b] = A[order[b] * l + order[b]];
        for (int a = 0; a < l; ++a) V[a * l + b] = Q[a * l + order[b]];
    }
}

}  // namespace swne
