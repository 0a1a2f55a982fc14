// kernels_misc.cuh -- ingest, layout conversion, sparse-view construction and bookkeeping kernels that
// do not depend on the padded factor count KP.  Included by swne_nmf.cu only.
#pragma once
#include "common.cuh"

namespace swne {

// device-side zero fill (a kernel, so it stays on the compute queue between the kernels it separates)
__global__ void zero_fill_kernel(uint4 *p16, size_t n16, uint32_t *tail, int ntail)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        p16[i] = make_uint4(0, 0, 0, 0);
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0;
}

// ----------------------------------------------------------------------------------------------
// ingest
// ----------------------------------------------------------------------------------------------

template <typename T>
__global__ void ingest_dense_kernel(const T *__restrict__ src, size_t src_ld, float *__restrict__ dst, size_t dst_ld,
                                    int n, int rows, MatStats *st)
{
    // one block per few rows (cells); threads stride over genes
    __shared__ double scratch[32];
    double s = 0, s2 = 0, kc = 0;
    unsigned long long nz = 0;
    unsigned int bad = 0;
    float mx = 0.f;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const T *sp = src + (size_t)r * src_ld;
        float *dp = dst ? dst + (size_t)r * dst_ld : nullptr;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double v = (double)sp[i];
            if (!(v >= 0.0)) bad |= (v < 0.0) ? 1u : 2u;
            if (isinf(v)) bad |= 2u;
            if (dp) dp[i] = (float)v;
            mx = fmaxf(mx, (float)v);
            s += v; s2 += v * v;
            if (v != 0.0) { ++nz; }
            kc += (v + TINY_NUM_D) * log(v + TINY_NUM_D) - v;
        }
        if (dp) for (int i = n + threadIdx.x; i < (int)dst_ld; i += blockDim.x) dp[i] = 0.f;
    }
    double r0 = block_sum(s, scratch), r1 = block_sum(s2, scratch), r2 = block_sum(kc, scratch);
    double r3 = block_sum((double)nz, scratch);
    bad = __syncthreads_or(bad);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(&st->max_bits, __float_as_uint(mx));
    if (threadIdx.x == 0) {
        atomicAdd(&st->sum, r0); atomicAdd(&st->sumsq, r1); atomicAdd(&st->klc, r2);
        atomicAdd(&st->nnz, (unsigned long long)(r3 + 0.5));
        if (bad) atomicOr(&st->bad, bad);
    }
}

// CSC (one column = one cell) -> dense rows.  dst must be zeroed.  one warp per cell.
template <typename T>
__global__ void ingest_csc_kernel(const int *__restrict__ p, const int *__restrict__ idx, const T *__restrict__ x,
                                  float *__restrict__ dst, size_t dst_ld, int n, int col0, int cols, int base,
                                  MatStats *st, const double *__restrict__ cell_scale = nullptr,
                                  const double *__restrict__ gene_scale = nullptr, int transform = 0)
{
    __shared__ double scratch[32];
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double s = 0, s2 = 0, kc = 0;
    unsigned long long nz = 0;
    unsigned int bad = 0;
    float mx = 0.f;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < cols; j += gridDim.x * wpb) {
        const int b = p[col0 + j] - base, e = p[col0 + j + 1] - base;
        for (int q = b + lane; q < e; q += 32) {
            const int g = idx[q];
            double v = (double)x[q];
            if (g < 0 || g >= n) { bad |= 4u; continue; }
            // ScaleCounts fused into the upload (R/data_processing.R:275-307): depth scaling per cell (NormalizeCounts,
            // :184), the variance scale factor per gene (:284-287), then log(x + 1) (:301) or Freeman-Tukey (:256,:299) on
            // the stored entries only, as the reference transforms the @x slot
            if (cell_scale) v *= cell_scale[col0 + j];
            if (gene_scale) v *= gene_scale[g];
            if (transform == 1) v = log1p(v);
            else if (transform == 2) v = sqrt(v) + sqrt(v + 1.0);
            if (!(v >= 0.0)) bad |= (v < 0.0) ? 1u : 2u;
            if (isinf(v) || isinf((float)v)) bad |= 2u;      // also a finite double beyond the fp32 range
            // duplicate (i, j) entries add up, as in a dgCMatrix built by Matrix::sparseMatrix (the statistics below count
            // them separately either way, which only matters for nnz)
            atomicAdd(&dst[(size_t)j * dst_ld + g], (float)v);
            mx = fmaxf(mx, (float)v);
            s += v; s2 += v * v;
            if (v != 0.0) { ++nz; kc += (v + TINY_NUM_D) * log(v + TINY_NUM_D) - v; }
        }
    }
    double r0 = block_sum(s, scratch), r1 = block_sum(s2, scratch), r2 = block_sum(kc, scratch);
    double r3 = block_sum((double)nz, scratch);
    bad = __syncthreads_or(bad);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(&st->max_bits, __float_as_uint(mx));
    if (threadIdx.x == 0) {
        atomicAdd(&st->sum, r0); atomicAdd(&st->sumsq, r1); atomicAdd(&st->klc, r2);
        atomicAdd(&st->nnz, (unsigned long long)(r3 + 0.5));
        if (bad) atomicOr(&st->bad, bad);
    }
}

// host layouts <-> device layouts for the factors.  src double [cols][k] (R's k x cols column-major)
// bad != nullptr: validate on the way in (negative / NaN / Inf entries set `bit`), so the host does not walk k x m doubles
__global__ void factor_in_cellmajor_kernel(const double *__restrict__ src, float *__restrict__ dst, int k, int KP, size_t cols,
                                           unsigned *bad = nullptr, unsigned bit = 0)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cols * (size_t)KP) return;
    const size_t j = t / KP; const int f = (int)(t % KP);
    const double v = (f < k) ? src[j * k + f] : 0.0;
    if (bad && (!(v >= 0.0) || isinf(v))) atomicOr(bad, bit);
    dst[t] = (float)v;
}
__global__ void factor_out_cellmajor_kernel(const float *__restrict__ src, double *__restrict__ dst, int k, int KP, size_t cols)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cols * (size_t)k) return;
    const size_t j = t / k; const int f = (int)(t % k);
    dst[t] = (double)src[j * KP + f];
}
// W: host n x k column-major  W[i + n f]  <->  device [n][KP]
__global__ void factor_in_colmajor_kernel(const double *__restrict__ src, float *__restrict__ dst, int k, int KP, int n,
                                          unsigned *bad = nullptr, unsigned bit = 0)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * KP) return;
    const int i = (int)(t / KP), f = (int)(t % KP);
    const double v = (f < k) ? src[(size_t)f * n + i] : 0.0;
    if (bad && (!(v >= 0.0) || isinf(v))) atomicOr(bad, bit);
    dst[t] = (float)v;
}
__global__ void factor_out_colmajor_kernel(const float *__restrict__ src, double *__restrict__ dst, int k, int KP, int n)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * k) return;
    const int f = (int)(t / n), i = (int)(t % n);
    dst[t] = (double)src[(size_t)i * KP + f];
}

// builds the GramPack the solvers use from the fp64 sums (one block):
//   G = G0 + (p0 - p1) I + p1 11' + tiny I   (update()), D = sqrt(diag G),
//   pack = [ G' = G / (D D') with unit diagonal | G0 | D | 1/D | column sums ], rows/cols >= k zeroed
__global__ void finalize_gram_kernel(const double *__restrict__ G64, float *__restrict__ pack, int k, int KP, double p0,
                                     double p1)
{
    __shared__ float sInv[64];
    const double *colsum64 = G64 + KP * KP;
    for (int t = threadIdx.x; t < KP; t += blockDim.x) {
        double g = 1.0;
        if (t < k) {
            g = G64[t * KP + t];
            if (p0 != p1) g += p0 - p1;
            if (p1 != 0.0) g += p1;
            g += TINY_NUM_D;
        }
        const float d = sqrtf((float)g);
        pack[2 * KP * KP + t] = d;
        pack[2 * KP * KP + KP + t] = 1.f / d;
        sInv[t] = 1.f / d;
        pack[2 * KP * KP + 2 * KP + t] = (t < k) ? (float)colsum64[t] : 0.f;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < KP * KP; t += blockDim.x) {
        const int a = t / KP, b = t % KP;
        float gs = 0.f, g0 = 0.f;
        if (a < k && b < k) {
            g0 = (float)G64[t];
            double g = G64[t];
            if (p1 != 0.0) g += p1;
            gs = (a == b) ? 1.f : (float)g * sInv[a] * sInv[b];
        }
        pack[t] = gs;
        pack[KP * KP + t] = g0;
    }
}

// ----------------------------------------------------------------------------------------------
// sparse views of a dense matrix (needed by the mkl path and by the mkl trace of mse fits)
//   by cell : cptr[m+1], cidx (gene), cval      by gene : gptr[n+1], gidx (cell), gval
// ----------------------------------------------------------------------------------------------
__global__ void count_nz_by_cell_kernel(const float *__restrict__ A, size_t ldA, int n, int m, int *__restrict__ cnt)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < m; j += gridDim.x * wpb) {
        int c = 0;
        for (int g = lane; g < n; g += 32) c += (A[(size_t)j * ldA + g] != 0.f);
        c = (int)warp_sum((float)c);
        if (lane == 0) cnt[j] = c;
    }
}
__global__ void fill_by_cell_kernel(const float *__restrict__ A, size_t ldA, int n, int m, const int *__restrict__ ptr,
                                    int *__restrict__ idx, float *__restrict__ val)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < m; j += gridDim.x * wpb) {
        int o = ptr[j];
        for (int g0 = 0; g0 < n; g0 += 32) {
            const int g = g0 + lane;
            const float v = (g < n) ? A[(size_t)j * ldA + g] : 0.f;
            const unsigned mask = __ballot_sync(0xffffffffu, v != 0.f);
            if (v != 0.f) {
                const int pos = o + __popc(mask & ((1u << lane) - 1u));
                idx[pos] = g; val[pos] = v;
            }
            o += __popc(mask);
        }
    }
}
// counts per (cell chunk, gene); chunk = 1024 cells.  cnt [nchunks][n]
__global__ void count_nz_by_gene_kernel(const float *__restrict__ A, size_t ldA, int n, int m, int chunk,
                                        int *__restrict__ cnt)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int c0 = blockIdx.y * chunk, c1 = min(m, c0 + chunk);
    int c = 0;
    for (int j = c0; j < c1; ++j) c += (A[(size_t)j * ldA + g] != 0.f);
    cnt[(size_t)blockIdx.y * n + g] = c;
}
// cnt [nchunks][n] -> per gene totals tot[n] and in-place exclusive offsets within the gene
__global__ void chunk_offsets_kernel(int *__restrict__ cnt, int nchunks, int n, int *__restrict__ tot)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    int s = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int v = cnt[(size_t)c * n + g];
        cnt[(size_t)c * n + g] = s;
        s += v;
    }
    tot[g] = s;
}
__global__ void fill_by_gene_kernel(const float *__restrict__ A, size_t ldA, int n, int m, int chunk,
                                    const int *__restrict__ off, const int *__restrict__ gptr, int *__restrict__ idx,
                                    float *__restrict__ val)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int c0 = blockIdx.y * chunk, c1 = min(m, c0 + chunk);
    int o = gptr[g] + off[(size_t)blockIdx.y * n + g];
    for (int j = c0; j < c1; ++j) {
        const float v = A[(size_t)j * ldA + g];
        if (v != 0.f) { idx[o] = j; val[o] = v; ++o; }
    }
}
// single block exclusive scan of int counts into ptr[0..len] (len up to a few million; setup only)
__global__ void exclusive_scan_kernel(const int *__restrict__ cnt, int *__restrict__ ptr, int len)
{
    __shared__ int s_part[1024];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < len; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < len) ? cnt[i] : 0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int add = (threadIdx.x >= o) ? s_part[threadIdx.x - o] : 0;
            __syncthreads();
            s_part[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < len) ptr[i] = s_carry + s_part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry += s_part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) ptr[len] = s_carry;
}

// ----------------------------------------------------------------------------------------------
// get_sample_coords (R/swne_plotting.R:70-84): per cell, the n_pull largest scores (ties: lower
// factor index first, as R's order(decreasing=TRUE)), weights h^alpha / sum, weighted mean of factor xy.
// One thread per cell; H double [m][k] (R layout), coords [k][2] col-major -> out [m][2] col-major.
// ----------------------------------------------------------------------------------------------
__global__ void sample_coords_kernel(const double *__restrict__ H, int k, int m, const double *__restrict__ coords,
                                     double alpha, int n_pull, double *__restrict__ out)
{
    extern __shared__ double s_coords[];  // 2k
    for (int t = threadIdx.x; t < 2 * k; t += blockDim.x) s_coords[t] = coords[t];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double *h = H + (size_t)j * k;
    unsigned long long used = 0ull;  // k <= 64
    double sx = 0.0, sy = 0.0, sw = 0.0;
    for (int p = 0; p < n_pull; ++p) {
        int best = -1;
        double bv = 0.0;
        for (int f = 0; f < k; ++f) {
            if (used >> f & 1ull) continue;
            const double v = h[f];
            if (best < 0 || v > bv) { best = f; bv = v; }
        }
        used |= 1ull << best;
        const double w = (alpha == 1.0) ? bv : pow(bv, alpha);
        sw += w; sx += s_coords[best] * w; sy += s_coords[k + best] * w;
    }
    // R: sum(coords * (w / pull.sum)) ; same up to rounding order
    out[j] = sx / sw;
    out[(size_t)m + j] = sy / sw;
}


}  // namespace swne
