// seed.cuh -- device side of RunNMF's seeding (R/nmf.R:12-66, 121-127): random start, NNDSVD split of a truncated SVD,
// FastICA fixed point, and the zero replacement, all writing the factors directly in the fit's fp32 layouts
//   Wt [n][KP] (gene-major)   H [m][KP] (cell-major)
// so that the fit starts without a host round trip.  Random numbers come from a counter-based generator (splitmix64 of
// seed and element index): R's runif / sample streams cannot be reproduced by any other program, so only the
// distribution is part of the contract (U(0, mean(A)/100) for the replacement, U(0,1) * 0.01 for NNLM's random start).
#pragma once
#include "common.cuh"

namespace swne {

__device__ __forceinline__ float seed_uniform(uint64_t seed, uint64_t idx)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return ((float)(uint32_t)(z >> 40) + 0.5f) * (1.f / 16777216.f);     // (0, 1)
}

// X [rows][KP]: columns < k <- U(0,1) * scale, padding columns <- 0
__global__ void seed_uniform_kernel(float *__restrict__ X, size_t rows, int k, int KP, float scale, uint64_t seed)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * (size_t)KP) return;
    const int f = (int)(t % KP);
    X[t] = (f < k) ? scale * seed_uniform(seed, t) : 0.f;
}

// R'R = G (upper Cholesky) and T = R^-1 as a KP x KP fp32 right-multiplier, in one block of 64 threads; *flag |= 1 when
// the leading l x l block of G is not numerically positive definite.  G64: KP x KP doubles (gram_kernel's output).
__global__ void __launch_bounds__(64) seed_chol_inv_kernel(const double *__restrict__ G64, int KP, int l, float *__restrict__ T,
                                                           unsigned *flag)
{
    __shared__ double R[64][65];
    __shared__ double sv[64];
    __shared__ int bad;
    const int c = threadIdx.x;
    if (c == 0) bad = 0;
    for (int a = 0; a < 64; ++a) R[a][c] = 0.0;
    __syncthreads();
    for (int j = 0; j < l; ++j) {
        if (c >= j && c < l) {
            double s = G64[j * KP + c];
            for (int q = 0; q < j; ++q) s -= R[q][j] * R[q][c];
            sv[c] = s;
        }
        __syncthreads();
        const double d = sv[j];
        if (!(d > 0.0) || !(d > 1e-14 * fabs(G64[j * KP + j]))) { if (c == 0) bad = 1; }
        const double rjj = sqrt(fmax(d, 1e-300));
        if (c >= j && c < l) R[j][c] = (c == j) ? rjj : sv[c] / rjj;
        __syncthreads();
    }
    // column c of the inverse by back substitution (kept by its thread), written straight into T
    double col[64];
    if (c < l) {
        col[c] = 1.0 / R[c][c];
        for (int i = c - 1; i >= 0; --i) {
            double s = 0.0;
            for (int q = i + 1; q <= c; ++q) s += R[i][q] * col[q];
            col[i] = -s / R[i][i];
        }
    }
    const int isbad = bad;
    if (c < KP)
        for (int a = 0; a < KP; ++a) T[a * KP + c] = (a <= c && c < l && !isbad) ? (float)col[a] : ((a == c) ? 1.f : 0.f);
    if (c == 0 && isbad) atomicOr(flag, 1u);
}

// squared norms of the positive and negative parts of every column: out[f] += sum x+^2, out[KP + f] += sum x-^2
__global__ void __launch_bounds__(256) seed_posneg_kernel(const float *__restrict__ X, size_t rows, int KP, double *__restrict__ out)
{
    // thread (r, f): rows strided by the grid; KP <= 64 columns
    const int f = threadIdx.x % KP, rsub = threadIdx.x / KP, rper = 256 / KP;
    double sp = 0.0, sn = 0.0;
    if (rsub < rper) {
        for (size_t r = (size_t)blockIdx.x * rper + rsub; r < rows; r += (size_t)gridDim.x * rper) {
            const float x = X[r * KP + f];
            if (x >= 0.f) sp += (double)x * x; else sn += (double)x * x;
        }
    }
    __shared__ double sp_s[256], sn_s[256];
    sp_s[threadIdx.x] = sp; sn_s[threadIdx.x] = sn;
    __syncthreads();
    if (threadIdx.x < KP) {
        double a = 0.0, b = 0.0;
        for (int q = 0; q < rper; ++q) { a += sp_s[q * KP + threadIdx.x]; b += sn_s[q * KP + threadIdx.x]; }
        atomicAdd(out + threadIdx.x, a);
        atomicAdd(out + KP + threadIdx.x, b);
    }
}

// dst[r][f] = coef-weighted part of src[r][f] followed by RunNMF's zero replacement (R/nmf.R:121-127):
//   mode[f] = 0: |x| * coef[f] (first NNDSVD triplet), 1: positive part * coef[f], 2: negative part * coef[f] (coef < 0),
//             3: x * coef[f] as is (signed: the ICA seed; its negatives are "< 1e-6" and get replaced like zeros)
//   then v < eps -> 0 and zeros -> U(0, fill_max)
// src [rows][KPs], dst [rows][KPd]; columns >= k of dst are zeroed.
__global__ void seed_apply_kernel(const float *__restrict__ src, int KPs, float *__restrict__ dst, int KPd, size_t rows, int k,
                                  const float *__restrict__ coef, const int *__restrict__ mode, float eps, float fill_max, uint64_t seed)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * (size_t)KPd) return;
    const size_t r = t / KPd;
    const int f = (int)(t % KPd);
    float v = 0.f;
    if (f < k) {
        const float x = src[r * KPs + f], c = coef[f];
        const int md = mode[f];
        if (md == 0) v = fabsf(x) * c;
        else if (md == 1) v = (x >= 0.f) ? x * c : 0.f;
        else if (md == 2) v = (x < 0.f) ? x * c : 0.f;
        else v = x * c;
        if (!(v >= eps)) v = 0.f;
        if (v == 0.f) v = fill_max * seed_uniform(seed, t);
    }
    dst[t] = v;
}

// out[f] += sum_r w[r] X[r][f]   (w == nullptr: plain column sums).  X [rows][KP], KP <= 64
__global__ void __launch_bounds__(256) seed_wcolsum_kernel(const float *__restrict__ X, size_t rows, int KP, const float *__restrict__ w,
                                                           double *__restrict__ out)
{
    const int f = threadIdx.x % KP, rsub = threadIdx.x / KP, rper = 256 / KP;
    double sacc = 0.0;
    if (rsub < rper)
        for (size_t r = (size_t)blockIdx.x * rper + rsub; r < rows; r += (size_t)gridDim.x * rper)
            sacc += (double)X[r * KP + f] * (w ? (double)w[r] : 1.0);
    __shared__ double sh[256];
    sh[threadIdx.x] = sacc;
    __syncthreads();
    if (threadIdx.x < KP) {
        double a = 0.0;
        for (int q = 0; q < rper; ++q) a += sh[q * KP + threadIdx.x];
        atomicAdd(out + threadIdx.x, a);
    }
}

// X[r][f] -= rowfac[r] * colvec[f]   (rowfac == nullptr: 1): the rank-one centring correction of a sketch product
__global__ void seed_rank1_sub_kernel(float *__restrict__ X, size_t rows, int KP, const float *__restrict__ rowfac,
                                      const double *__restrict__ colvec)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * (size_t)KP) return;
    X[t] -= (rowfac ? rowfac[t / KP] : 1.f) * (float)colvec[t % KP];
}

// per-gene sums over the cells of the resident matrix: out[g] += sum_j A[j][g]   (A cell-major [m][ldA])
__global__ void __launch_bounds__(256) seed_gene_sums_kernel(const float *__restrict__ A, size_t ldA, int n, int m, int cells_per_block,
                                                             double *__restrict__ out)
{
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g >= n) return;
    const int c0 = blockIdx.y * cells_per_block, c1 = min(m, c0 + cells_per_block);
    double sacc = 0.0;
    for (int j = c0; j < c1; ++j) sacc += (double)A[(size_t)j * ldA + g];
    atomicAdd(out + g, sacc);
}
__global__ void seed_scale_to_float_kernel(const double *__restrict__ in, float *__restrict__ out, int n, double scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)(in[i] * scale);
}

}  // namespace swne
