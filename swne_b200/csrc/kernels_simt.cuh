// kernels_simt.cuh -- fp32 CUDA-core kernels of the NMF hot path (engine SWNE_ENGINE_SIMT) plus the
// ingest / layout / bookkeeping kernels every engine shares.
//
// Device layouts (all fp32 unless noted):
//   dA  [m][ldA]  cell-major: row j = the n gene values of cell j   (== R's column-major n x m)
//   dWt [n][KP]   gene-major: row g = the k factor loadings of gene g, zero padded to KP
//   dH  [m][KP]   cell-major: row j = the k factor scores of cell j, zero padded to KP
// KP = k rounded up to a multiple of 8 (<= 64).  Gram matrices are KP x KP, symmetric.
//
// What each kernel restates (NNLM restated):
//   contract_cells_kernel  : Wt * A[:,j]      inside update()           (H side)
//   contract_genes_kernel  : H  * At[:,g]     inside update()           (W side, no A.t() copy)
//   gram_kernel            : G = X X'         inside update()
//   scd_ls_solve_kernel    : mu = G h - X a (+L1); scd_ls_update(h, G, mu, 50, 1e-9)
//   scd_kl_kernel          : scd_kl_update (sparse aware: zero entries of A add exactly 0 to a2 and b)
#pragma once
#include "common.cuh"

namespace swne {

// ----------------------------------------------------------------------------------------------
// Gram: G64[KP*KP] += X'X, colsum64[KP] += column sums, X [rows][KP].  256 threads, grid <= 2 x #SM
// ----------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) gram_kernel(const float *__restrict__ X, int rows, double *__restrict__ G64,
                                                   double *__restrict__ colsum64)
{
    // 2 x 2 register blocks of the KP x KP Gram: two 8-byte shared-memory loads feed four FMAs (one entry per thread and two
    // loads per FMA made the kernel shared-memory bound: 0.9 ms for 1.3 M x 40, round 2)
    constexpr int ROWS = 64;
    constexpr int HB = KP / 2;                        // 2 x 2 blocks per dimension
    constexpr int BPT = (HB * HB + 255) / 256;        // blocks per thread
    __shared__ __align__(16) float xs[ROWS][KP + 2];
    float acc[BPT][4];
#pragma unroll
    for (int q = 0; q < BPT; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
    float cs = 0.f;
    // contiguous row range per block; few blocks so that the fp64 atomics at the end stay cheap
    const int per = ((rows + gridDim.x - 1) / gridDim.x + ROWS - 1) / ROWS * ROWS;
    const int row_begin = blockIdx.x * per;
    const int row_end = min(rows, row_begin + per);
    for (int r0 = row_begin; r0 < row_end; r0 += ROWS) {
        const int nr = min(ROWS, row_end - r0);
        __syncthreads();
        for (int t = threadIdx.x; t < ROWS * KP; t += 256) {
            const int r = t / KP, f = t % KP;
            xs[r][f] = (r < nr) ? X[(size_t)(r0 + r) * KP + f] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < BPT; ++q) {
            const int blk = threadIdx.x + q * 256;
            if (blk < HB * HB) {
                const int a0 = 2 * (blk / HB), b0 = 2 * (blk % HB);
#pragma unroll 8
                for (int r = 0; r < ROWS; ++r) {
                    const float2 xa = *reinterpret_cast<const float2 *>(&xs[r][a0]);
                    const float2 xb = *reinterpret_cast<const float2 *>(&xs[r][b0]);
                    acc[q][0] = fmaf(xa.x, xb.x, acc[q][0]); acc[q][1] = fmaf(xa.x, xb.y, acc[q][1]);
                    acc[q][2] = fmaf(xa.y, xb.x, acc[q][2]); acc[q][3] = fmaf(xa.y, xb.y, acc[q][3]);
                }
            }
        }
        if (threadIdx.x < KP) {
            float sacc = 0.f;
#pragma unroll 8
            for (int r = 0; r < ROWS; ++r) sacc += xs[r][threadIdx.x];
            cs += sacc;
        }
    }
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
        const int blk = threadIdx.x + q * 256;
        if (blk < HB * HB) {
            const int a0 = 2 * (blk / HB), b0 = 2 * (blk % HB);
            atomicAdd(&G64[a0 * KP + b0], (double)acc[q][0]); atomicAdd(&G64[a0 * KP + b0 + 1], (double)acc[q][1]);
            atomicAdd(&G64[(a0 + 1) * KP + b0], (double)acc[q][2]); atomicAdd(&G64[(a0 + 1) * KP + b0 + 1], (double)acc[q][3]);
        }
    }
    if (threadIdx.x < KP) atomicAdd(&colsum64[threadIdx.x], (double)cs);
}

// ----------------------------------------------------------------------------------------------
// One FastICA iteration (ica::icafast, alg = "par", fun = "logcosh", alpha = 1; call sites R/nmf.R:59,63) on the whitened
// data Y [rows][KP] (nc sources):   S = Y R,  g = tanh(S),  out[a KP + b] += sum_r y_a g_b,  out[KP^2 + b] += sum_r (1 - g_b^2)
// (the host forms R+ = Y'g / nobs - R diag(mean(1 - g^2)) and orthogonalises it: nc x nc work).  One pass over Y.
// ----------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) ica_step_kernel(const float *__restrict__ Y, int rows, int nc, const float *__restrict__ Rm,
                                                       double *__restrict__ out)
{
    constexpr int ROWS = 32;
    constexpr int PAIRS = (KP * KP + 255) / 256;
    __shared__ float ys[ROWS][KP + 1];
    __shared__ float gs[ROWS][KP + 1];
    __shared__ float sR[KP][KP + 1];
    __shared__ float sg2[KP];
    for (int t = threadIdx.x; t < KP * KP; t += 256) sR[t / KP][t % KP] = Rm[t];
    if (threadIdx.x < KP) sg2[threadIdx.x] = 0.f;
    float acc[PAIRS];
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) acc[q] = 0.f;
    const int per = ((rows + gridDim.x - 1) / gridDim.x + ROWS - 1) / ROWS * ROWS;
    const int row_begin = blockIdx.x * per;
    const int row_end = min(rows, row_begin + per);
    for (int r0 = row_begin; r0 < row_end; r0 += ROWS) {
        const int nr = min(ROWS, row_end - r0);
        __syncthreads();
        for (int t = threadIdx.x; t < ROWS * KP; t += 256) {
            const int r = t / KP, f = t % KP;
            ys[r][f] = (r < nr) ? Y[(size_t)(r0 + r) * KP + f] : 0.f;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < ROWS * KP; t += 256) {
            const int r = t / KP, b = t % KP;
            float g = 0.f;
            if (b < nc && r < nr) {
                float sv = 0.f;
                for (int a = 0; a < nc; ++a) sv = fmaf(ys[r][a], sR[a][b], sv);
                g = tanhf(sv);
                atomicAdd(&sg2[b], 1.f - g * g);
            }
            gs[r][b] = g;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) {
            const int pr = threadIdx.x + q * 256;
            if (pr < KP * KP) {
                const int a = pr / KP, b = pr % KP;
                float sv = 0.f;
#pragma unroll 8
                for (int r = 0; r < ROWS; ++r) sv = fmaf(ys[r][a], gs[r][b], sv);
                acc[q] += sv;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) {
        const int pr = threadIdx.x + q * 256;
        if (pr < KP * KP && pr / KP < nc && pr % KP < nc) atomicAdd(&out[pr], (double)acc[q]);
    }
    if (threadIdx.x < nc) atomicAdd(&out[KP * KP + threadIdx.x], (double)sg2[threadIdx.x]);
}

// ----------------------------------------------------------------------------------------------
// C[m][KP] = A[m][n] * Wt[n][KP]        (H side: c_j = Wt a_j for a tile of 64 cells)
// 128 threads; thread (ty = t/8, tx = t%8) owns cells ty*4..+4 and factors tx*FPT..+FPT
// ----------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(128) contract_cells_kernel(const float *__restrict__ A, size_t ldA, int n, int m,
                                                             const float *__restrict__ Wt, float *__restrict__ C)
{
    constexpr int FPT = KP / 8, TM = 64, TK = 32;
    __shared__ float as[TM][TK + 1];
    __shared__ float ws[TK][KP];
    const int t = threadIdx.x, ty = t >> 3, tx = t & 7, lane = t & 31, wid = t >> 5;
    const int cell0 = blockIdx.x * TM;
    float acc[4][FPT];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int f = 0; f < FPT; ++f) acc[a][f] = 0.f;
    for (int g0 = 0; g0 < n; g0 += TK) {
        __syncthreads();
        // A tile: each warp loads 16 cell rows x 32 genes (128 B coalesced per row)
#pragma unroll 4
        for (int r = wid; r < TM; r += 4) {
            const int cell = cell0 + r, g = g0 + lane;
            as[r][lane] = (cell < m && g < n) ? __ldg(A + (size_t)cell * ldA + g) : 0.f;
        }
        for (int q = t; q < TK * KP; q += 128) {
            const int g = g0 + q / KP;
            ws[q / KP][q % KP] = (g < n) ? Wt[(size_t)g * KP + q % KP] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < TK; ++kk) {
            float av[4], wv[FPT];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = as[ty * 4 + a][kk];
#pragma unroll
            for (int f = 0; f < FPT; ++f) wv[f] = ws[kk][tx * FPT + f];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int f = 0; f < FPT; ++f) acc[a][f] = fmaf(av[a], wv[f], acc[a][f]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int cell = cell0 + ty * 4 + a;
        if (cell < m)
#pragma unroll
            for (int f = 0; f < FPT; ++f) C[(size_t)cell * KP + tx * FPT + f] = acc[a][f];
    }
}

// ----------------------------------------------------------------------------------------------
// B[n][KP] += A[cells]' * H[cells][KP]   (W side: b_g = H At[:,g]); split over cell chunks, fp32 atomics
// grid (gene tiles of 64, cell chunks); 128 threads; thread (ty, tx) owns genes ty*4..+4, factors tx*FPT..
// ----------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(128) contract_genes_kernel(const float *__restrict__ A, size_t ldA, int n, int m,
                                                             const float *__restrict__ H, float *__restrict__ B,
                                                             int cells_per_chunk)
{
    constexpr int FPT = KP / 8, TG = 64, TK = 32;
    __shared__ __align__(16) float as[TK][TG + 4];
    __shared__ __align__(16) float hs[TK][KP];
    const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
    const int gene0 = blockIdx.x * TG;
    const int c_begin = blockIdx.y * cells_per_chunk, c_end = min(m, c_begin + cells_per_chunk);
    float acc[4][FPT];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int f = 0; f < FPT; ++f) acc[a][f] = 0.f;
    for (int c0 = c_begin; c0 < c_end; c0 += TK) {
        __syncthreads();
        // 32 cells x 64 genes: half warp per cell row (256 B contiguous)
        for (int q = t; q < TK * TG; q += 128) {
            const int r = q / TG, gi = q % TG;
            const int cell = c0 + r, g = gene0 + gi;
            as[r][gi] = (cell < c_end && g < n) ? __ldg(A + (size_t)cell * ldA + g) : 0.f;
        }
        for (int q = t; q < TK * KP; q += 128) {
            const int cell = c0 + q / KP;
            hs[q / KP][q % KP] = (cell < c_end) ? H[(size_t)cell * KP + q % KP] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < TK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&as[kk][ty * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            float hv[FPT];
#pragma unroll
            for (int f = 0; f < FPT; ++f) hv[f] = hs[kk][tx * FPT + f];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int f = 0; f < FPT; ++f) acc[a][f] = fmaf(av[a], hv[f], acc[a][f]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int g = gene0 + ty * 4 + a;
        if (g < n)
#pragma unroll
            for (int f = 0; f < FPT; ++f) atomicAdd(&B[(size_t)g * KP + tx * FPT + f], acc[a][f]);
    }
}

// ----------------------------------------------------------------------------------------------
// SCD least squares: one thread per column, h and mu in registers, G broadcast from shared memory.
// Semantics of scd_ls_update kept: coordinate order 0..k-1, clamp at 0, skip when unchanged,
// continue while any coordinate's relative change 2|d|/(new+old+tiny) exceeds rel_tol, cap max_iter,
// warm start from the previous iterate.
// ----------------------------------------------------------------------------------------------
// Unit-diagonal form: with d_k = sqrt(G_kk), h' = d o h, mu' = mu / d and G' = G / (d d') the update
//   h_k <- max(0, h_k - mu_k / G_kk)   becomes   h'_k <- max(0, h'_k - mu'_k),   mu' += delta' * G'[:,k]
// (same iterates up to rounding; the relative-change test is invariant).  It removes one multiply and one
// load from the dependent chain of every coordinate step.
//
// Where G' lives: in the __constant__ bank.  Every thread of a warp needs the same KP floats per coordinate
// step; from shared memory that is 8 broadcast LDS.128 per warp-step and the SM-wide shared-memory pipe
// saturates at ~19 SM-cycles per warp-step (tools/micro/solver_bench.cu, measured on B200).  Constant-bank
// operands arrive through the per-SMSP uniform datapath (LDCU.128 -> FFMA2 with uniform operands) and the
// sweep becomes FMA-issue bound at ~11 SM-cycles per warp-step.
//
// GramPack (device, floats): [0,KP^2) G' unit-diagonal scaled | [KP^2,2KP^2) plain G0 | D[KP] | invD[KP] | colsum[KP]
constexpr int GRAM_KP_MAX = 64;
static __constant__ float c_gram[2 * GRAM_KP_MAX * GRAM_KP_MAX];   // per translation unit: G' then G0, row stride KP

__host__ __device__ constexpr int gram_pack_floats(int KP) { return 2 * KP * KP + 3 * KP; }

// One sweep block: coordinates kb*8 .. kb*8+7, branch free so the compiler can pipeline the shared-memory
// loads of the next coordinate under the FMAs of the current one.  An unchanged coordinate has d == 0 and
// the update is a no-op, which is exactly scd_ls_update's "if (tmp != h)" skip; padded coordinates
// (kk >= k: h = 0, mu >= 0, zero row of G') are no-ops too.  mu is held as float2 pairs and updated with
// the packed FFMA2 of sm_100 (two fp32 FMAs per issue slot).
template <int KP, bool ANY_CHANGE, bool PACKED = true>
__device__ __forceinline__ void scd_ls_block8(float (&h)[KP], float2 (&mu)[KP / 2], int kb, float rel_tol, unsigned &changed)
{
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int kk = kb * 8 + u;
        float4 g[KP / 4];
#pragma unroll
        for (int r = 0; r < KP / 4; ++r)
            g[r] = make_float4(c_gram[kk * KP + 4 * r], c_gram[kk * KP + 4 * r + 1], c_gram[kk * KP + 4 * r + 2],
                               c_gram[kk * KP + 4 * r + 3]);
        const float mk = (kk & 1) ? mu[kk / 2].y : mu[kk / 2].x;
        const float tmp = fmaxf(h[kk] - mk, 0.f);
        const float d = tmp - h[kk];
        const float2 d2 = make_float2(d, d);
        if (PACKED) {
#pragma unroll
            for (int r = 0; r < KP / 4; ++r) {
                mu[2 * r + 0] = __ffma2_rn(d2, make_float2(g[r].x, g[r].y), mu[2 * r + 0]);
                mu[2 * r + 1] = __ffma2_rn(d2, make_float2(g[r].z, g[r].w), mu[2 * r + 1]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < KP / 4; ++r) {
                mu[2 * r + 0].x = fmaf(d, g[r].x, mu[2 * r + 0].x);
                mu[2 * r + 0].y = fmaf(d, g[r].y, mu[2 * r + 0].y);
                mu[2 * r + 1].x = fmaf(d, g[r].z, mu[2 * r + 1].x);
                mu[2 * r + 1].y = fmaf(d, g[r].w, mu[2 * r + 1].y);
            }
        }
        if (ANY_CHANGE) changed |= __float_as_uint(d);   // d is +0 exactly when unchanged
        else changed |= (2.f * fabsf(d) > rel_tol * (tmp + h[kk] + TINY_NUM_F)) ? 1u : 0u;
        h[kk] = tmp;
    }
}

// h, mu in the scaled variables; G' is read from c_gram.  returns #sweeps
template <int KP, bool PACKED = true>
__device__ __forceinline__ int scd_ls_thread(float (&h)[KP], float2 (&mu)[KP / 2], int k, int max_iter, float rel_tol)
{
    int t = 0;
    unsigned changed = 1u;
    // with rel_tol below fp32 resolution "relative change > rel_tol" is the same as "changed at all"
    if (rel_tol < 2.9e-8f) {
        for (; t < max_iter && changed; ++t) {
            changed = 0u;
#pragma unroll
            for (int kb = 0; kb < KP / 8; ++kb)
                if (kb * 8 < k) scd_ls_block8<KP, true, PACKED>(h, mu, kb, rel_tol, changed);
        }
    } else {
        for (; t < max_iter && changed; ++t) {
            changed = 0u;
#pragma unroll
            for (int kb = 0; kb < KP / 8; ++kb)
                if (kb * 8 < k) scd_ls_block8<KP, false, PACKED>(h, mu, kb, rel_tol, changed);
        }
    }
    return t;
}

// Shared-memory variant with NC columns per thread: one broadcast LDS.128 fetch of a Gram row feeds NC
// independent chains (the SM-wide shared-memory pipe, not the FMA pipe, limits the one-column form: ~19.5
// SM-cycles per warp-step against ~14 with two columns; tools/micro/solver_bench2.cu).  Used inside the
// tcgen05 pass kernel, where the compiler funnels constant-bank operands through a single uniform-register
// quad and the constant-bank form loses its advantage.
template <int KP, int NC, bool ANY_CHANGE>
__device__ __forceinline__ void scd_ls_block8_smem(float (&h)[NC][KP], float2 (&mu)[NC][KP / 2], unsigned sg_addr, int kb,
                                                   float rel_tol, unsigned (&changed)[NC])
{
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int kk = kb * 8 + u;
        // asm volatile: plain loads of the loop-invariant G' get hoisted out of the sweep loop and spilled
        float4 g[KP / 4];
#pragma unroll
        for (int r = 0; r < KP / 4; ++r)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(g[r].x), "=f"(g[r].y), "=f"(g[r].z), "=f"(g[r].w)
                         : "r"(sg_addr + (unsigned)((kk * KP + 4 * r) * 4)));
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float mk = (kk & 1) ? mu[c][kk / 2].y : mu[c][kk / 2].x;
            const float tmp = fmaxf(h[c][kk] - mk, 0.f);
            const float d = tmp - h[c][kk];
            const float2 d2 = make_float2(d, d);
#pragma unroll
            for (int r = 0; r < KP / 4; ++r) {
                mu[c][2 * r + 0] = __ffma2_rn(d2, make_float2(g[r].x, g[r].y), mu[c][2 * r + 0]);
                mu[c][2 * r + 1] = __ffma2_rn(d2, make_float2(g[r].z, g[r].w), mu[c][2 * r + 1]);
            }
            if (ANY_CHANGE) changed[c] |= __float_as_uint(d);
            else changed[c] |= (2.f * fabsf(d) > rel_tol * (tmp + h[c][kk] + TINY_NUM_F)) ? 1u : 0u;
            h[c][kk] = tmp;
        }
    }
}

// sweeps[c] = the number of sweeps scd_ls_update would have run on column c (the sweep that finds no change
// included); extra sweeps a converged column sits through while its neighbour finishes are exact no-ops.
template <int KP, int NC>
__device__ __forceinline__ void scd_ls_thread_smem(float (&h)[NC][KP], float2 (&mu)[NC][KP / 2], unsigned sg_addr, int k,
                                                   int max_iter, float rel_tol, int (&sweeps)[NC])
{
    const bool any_change = rel_tol < 2.9e-8f;
#pragma unroll
    for (int c = 0; c < NC; ++c) sweeps[c] = 1;
    unsigned any = 1u;
    for (int t = 0; t < max_iter && any; ++t) {
        unsigned changed[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) changed[c] = 0u;
        if (any_change) {
#pragma unroll
            for (int kb = 0; kb < KP / 8; ++kb)
                if (kb * 8 < k) scd_ls_block8_smem<KP, NC, true>(h, mu, sg_addr, kb, rel_tol, changed);
        } else {
#pragma unroll
            for (int kb = 0; kb < KP / 8; ++kb)
                if (kb * 8 < k) scd_ls_block8_smem<KP, NC, false>(h, mu, sg_addr, kb, rel_tol, changed);
        }
        any = 0u;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (changed[c]) sweeps[c] = min(t + 2, max_iter);
            any |= changed[c];
        }
    }
}

// mu' += sum_q h'_q G'[q,:]  with h'_q = X[q] * D[q]   (G' from the constant bank, q uniform)
template <int KP>
__device__ __forceinline__ void scd_mu_init(float2 (&mu)[KP / 2], const float *__restrict__ xrow, const float *sD, int k)
{
#pragma unroll 1
    for (int q = 0; q < k; ++q) {
        const float hq = xrow[q] * sD[q];
        const float2 h2 = make_float2(hq, hq);
#pragma unroll
        for (int r = 0; r < KP / 4; ++r) {
            mu[2 * r] = __ffma2_rn(h2, make_float2(c_gram[q * KP + 4 * r], c_gram[q * KP + 4 * r + 1]), mu[2 * r]);
            mu[2 * r + 1] = __ffma2_rn(h2, make_float2(c_gram[q * KP + 4 * r + 2], c_gram[q * KP + 4 * r + 3]), mu[2 * r + 1]);
        }
    }
}
// t = G0 x  (plain Gram, second half of c_gram)
template <int KP>
__device__ __forceinline__ void scd_g0_times(float2 (&t)[KP / 2], const float *__restrict__ xrow, int k)
{
#pragma unroll
    for (int r = 0; r < KP / 2; ++r) t[r] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int q = 0; q < k; ++q) {
        const float hq = xrow[q];
        const float2 h2 = make_float2(hq, hq);
        const float *g0 = c_gram + KP * KP + q * KP;
#pragma unroll
        for (int r = 0; r < KP / 4; ++r) {
            t[2 * r] = __ffma2_rn(h2, make_float2(g0[4 * r], g0[4 * r + 1]), t[2 * r]);
            t[2 * r + 1] = __ffma2_rn(h2, make_float2(g0[4 * r + 2], g0[4 * r + 3]), t[2 * r + 1]);
        }
    }
}

// X [cols][KP] in/out (warm start), C [cols][KP] = the contraction (Wt a_j); the Grams come from c_gram
// (uploaded from `pack` by the launcher) and pack's D / invD; l1 = the L1 penalty added to mu.
// blockDim.x = 32 (few columns: one warp per SM keeps the latency-bound W side spread out) or 128.
template <int KP>
__global__ void __launch_bounds__(128) scd_ls_solve_kernel(float *__restrict__ X, const float *__restrict__ C,
                                                           const float *__restrict__ pack, float l1, int k, int cols,
                                                           int max_iter, float rel_tol, unsigned long long *sweeps,
                                                           double *loss_part)
{
    __shared__ float sD[KP], sInvD[KP];
    __shared__ double scratch[32];
    for (int t = threadIdx.x; t < KP; t += blockDim.x) {
        sD[t] = pack[2 * KP * KP + t];
        sInvD[t] = pack[2 * KP * KP + KP + t];
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int it = 0;
    double lp = 0.0;
    if (j < cols) {
        float h[KP];
        float2 mu[KP / 2];
        float *xrow = X + (size_t)j * KP;
        const float4 *hp = reinterpret_cast<const float4 *>(xrow);
        const float4 *cp = reinterpret_cast<const float4 *>(C + (size_t)j * KP);
#pragma unroll
        for (int r = 0; r < KP / 4; ++r) {
            const float4 v = hp[r];
            h[4 * r] = v.x * sD[4 * r]; h[4 * r + 1] = v.y * sD[4 * r + 1];
            h[4 * r + 2] = v.z * sD[4 * r + 2]; h[4 * r + 3] = v.w * sD[4 * r + 3];
            const float4 c = cp[r];
            mu[2 * r] = make_float2((l1 - c.x) * sInvD[4 * r], (l1 - c.y) * sInvD[4 * r + 1]);
            mu[2 * r + 1] = make_float2((l1 - c.z) * sInvD[4 * r + 2], (l1 - c.w) * sInvD[4 * r + 3]);
        }
        scd_mu_init<KP>(mu, xrow, sD, k);
        it = scd_ls_thread<KP>(h, mu, k, max_iter, rel_tol);
        float4 *op = reinterpret_cast<float4 *>(xrow);
#pragma unroll
        for (int r = 0; r < KP / 4; ++r) {
            h[4 * r] *= sInvD[4 * r]; h[4 * r + 1] *= sInvD[4 * r + 1];
            h[4 * r + 2] *= sInvD[4 * r + 2]; h[4 * r + 3] *= sInvD[4 * r + 3];
            op[r] = make_float4(h[4 * r], h[4 * r + 1], h[4 * r + 2], h[4 * r + 3]);
        }
        if (loss_part) {
            // h' G0 h - 2 h' c ; mu is dead, reuse it for t = G0 h  (xrow was just written by this thread)
            scd_g0_times<KP>(mu, xrow, k);
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < KP / 4; ++r) {
                const float4 c = cp[r];
                s = fmaf(h[4 * r + 0], mu[2 * r].x - 2.f * c.x, s);
                s = fmaf(h[4 * r + 1], mu[2 * r].y - 2.f * c.y, s);
                s = fmaf(h[4 * r + 2], mu[2 * r + 1].x - 2.f * c.z, s);
                s = fmaf(h[4 * r + 3], mu[2 * r + 1].y - 2.f * c.w, s);
            }
            lp = (double)s;
        }
    }
    const double tot_it = block_sum((double)it, scratch);
    if (threadIdx.x == 0) atomicAdd(sweeps, (unsigned long long)(tot_it + 0.5));
    if (loss_part) {
        const double tot = block_sum(lp, scratch);
        if (threadIdx.x == 0) atomicAdd(loss_part, tot);
    }
}

// ----------------------------------------------------------------------------------------------
// SCD KL update, sparse aware, one NT-thread block per column (NT sized to the column length so that
// every thread owns one or two nonzeros: the per-coordinate chain is latency bound, not work bound).  (scd_kl_update)
//   ptr/idx/val : the sparse view whose "columns" are the columns being solved
//   X  [cols][KP] in/out;  Y [rows][KP] the fixed factor (row r = factors of entity r);  sumY[KP]
//   order / sorted_cnt : the columns by descending length (device) and their lengths (host), or null
//   err: bit 0 (KL_ERR) add sum a log(ahat+tiny) and sum a^2-2 a ahat;  bit 1 (KL_L2_LATE) the other reading of NNLM's
//        Newton step with an L2 penalty (see kl_column)
// A column with at most `cap` nonzeros is staged once in shared memory -- the gathered rows of Y
// transposed to Yt[f][q] (lanes read consecutive words), A-hat on the support and the values -- so the
// k coordinate passes of every sweep run out of shared memory instead of re-gathering Y through L2.
// Longer columns take the same code on the global arrays (ajt_g is the nnz-sized A-hat scratch).
// ----------------------------------------------------------------------------------------------
constexpr int KL_ERR = 1, KL_L2_LATE = 2;
template <int KP, int NT, bool SM>
__device__ __forceinline__ int kl_column(const int *__restrict__ idx, const float *__restrict__ val,
                                         float *__restrict__ ajt_g, const float *__restrict__ Y,
                                         const float *__restrict__ sumY, float b0, float b1, float b2, int k, int nz,
                                         int max_iter, float rel_tol, int cap, float *Yt, float *aj, float *av, float *h,
                                         float (*part)[2][NT / 32], int err, double &klp, double &sqp)
{
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // stage / initialise A-hat on the support
    for (int q = tid; q < nz; q += NT) {
        const float4 *y4 = reinterpret_cast<const float4 *>(Y + (size_t)idx[q] * KP);
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < KP / 4; ++r) {
            const float4 y = __ldg(y4 + r);
            const float4 hh = *reinterpret_cast<const float4 *>(h + 4 * r);
            s = fmaf(y.x, hh.x, s); s = fmaf(y.y, hh.y, s); s = fmaf(y.z, hh.z, s); s = fmaf(y.w, hh.w, s);
            if (SM) {
                Yt[(4 * r + 0) * cap + q] = y.x; Yt[(4 * r + 1) * cap + q] = y.y;
                Yt[(4 * r + 2) * cap + q] = y.z; Yt[(4 * r + 3) * cap + q] = y.w;
            }
        }
        if (SM) { aj[q] = s; av[q] = val[q]; }
        else ajt_g[q] = s;
    }
    float sumH = 0.f;
    for (int f = 0; f < k; ++f) sumH += h[f];
    int it = 0;
    bool again = true;
    // The A-hat update of coordinate kk (aj += d w_kk) is applied lazily inside the accumulation loop of
    // the next coordinate: one pass over the column per coordinate instead of two, and the update of the
    // last coordinate of the last sweep is never needed.
    float d_prev = 0.f;
    int k_prev = 0;
    float *part_f = &part[0][0][0];
    for (; it < max_iter && again; ++it) {
        again = false;
        for (int kk = 0; kk < k; ++kk) {
            const float hk = h[kk];   // read before the barrier: thread 0 rewrites it after the barrier
            float a2 = 0.f, bb = 0.f;
#pragma unroll 2
            for (int q = tid; q < nz; q += NT) {
                float a = SM ? aj[q] : ajt_g[q];
                if (d_prev != 0.f) {
                    // exact arithmetic keeps A-hat >= 0 (d >= -h_kk); fp32 rounding may not: clamp
                    const float wp = SM ? Yt[k_prev * cap + q] : Y[(size_t)idx[q] * KP + k_prev];
                    a = fmaxf(fmaf(d_prev, wp, a), 0.f);
                    if (SM) aj[q] = a; else ajt_g[q] = a;
                }
                const float w = SM ? Yt[kk * cap + q] : Y[(size_t)idx[q] * KP + kk];
                const float v = SM ? av[q] : val[q];
                // a + 1e-16 is a normal number far below 2^126: the 2-ulp fast division is safe here
                const float mu = __fdividef(w, a + TINY_NUM_F);
                a2 = fmaf(v * mu, mu, a2);
                bb = fmaf(v, mu, bb);
            }
            // warp fold of the pair (a2, bb) in one butterfly: after the first exchange even lanes carry a2,
            // odd lanes bb
            float x = (lane & 1) ? bb : a2;
            const float y = (lane & 1) ? a2 : bb;
            x += __shfl_xor_sync(0xffffffffu, y, 1);
#pragma unroll
            for (int o = 2; o < 32; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            float *pk = part_f + (kk & 1) * 2 * NW;
            if (lane < 2) pk[lane * NW + wid] = x;
            __syncthreads();
            // every warp folds the 2 NW partials with the same butterfly: identical totals in all threads
            x = (lane < 2 * NW) ? pk[lane] : 0.f;
#pragma unroll
            for (int o = NW / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            a2 = __shfl_sync(0xffffffffu, x, 0);
            bb = __shfl_sync(0xffffffffu, x, NW & 31);
            bb -= sumY[kk];
            // NNLM's scd_kl_update as recalled: "a += beta0; b += a*h - ..." -- the L2 penalty is part of
            // the `a` that multiplies h.  The other reading (the exact Newton step: b += a*h BEFORE a += beta0) is kept
            // behind KL_L2_LATE so that a golden vector from real NNLM can decide without a code change; the two differ
            // only when the L2 penalty (alpha[0] / beta[0]) is non-zero.
            float a2h;
            if (err & KL_L2_LATE) { a2h = a2 * hk; a2 += b0; }
            else { a2 += b0; a2h = a2 * hk; }
            bb += a2h - b2 - b1 * (sumH - hk);
            float tmp = bb / (a2 + TINY_NUM_F);
            if (!(tmp > 0.f)) tmp = 0.f;
            d_prev = tmp - hk;
            k_prev = kk;
            if (tmp != hk) {
                again |= (2.f * fabsf(d_prev) > rel_tol * (tmp + hk + TINY_NUM_F));
                sumH += d_prev;
                if (tid == 0) h[kk] = tmp;
            }
        }
        __syncthreads();   // h and the partial-sum slots are reused by the next sweep
    }
    if (err & KL_ERR) {
        // A-hat recomputed from the final h: the incrementally updated values keep fp32 residue where the
        // exact value is 0, and log(ahat + 1e-16) is sensitive exactly there
        for (int q = tid; q < nz; q += NT) {
            float s = 0.f;
            if (SM) {
                for (int f = 0; f < k; ++f) s = fmaf(Yt[f * cap + q], h[f], s);
            } else {
                const float *y = Y + (size_t)idx[q] * KP;
                for (int f = 0; f < k; ++f) s = fmaf(y[f], h[f], s);
            }
            const double a = (double)(SM ? av[q] : val[q]), ah = (double)s;
            klp += a * log(ah + TINY_NUM_D);
            sqp += a * a - 2.0 * a * ah;
        }
    }
    return it;
}

template <int KP, int NT>
__global__ void __launch_bounds__(NT) scd_kl_kernel(const int *__restrict__ ptr, const int *__restrict__ idx,
                                                    const float *__restrict__ val, float *__restrict__ ajt,
                                                    float *__restrict__ X, const float *__restrict__ Y,
                                                    const float *__restrict__ sumY, float b0, float b1, float b2, int k,
                                                    int cols, int max_iter, float rel_tol, unsigned long long *sweeps,
                                                    double *kl_part, double *sq_part, int err, int cap,
                                                    const int *__restrict__ order, int first)
{
    extern __shared__ __align__(16) float kl_dyn[];   // Yt[KP][cap] | aj[cap] | av[cap]
    __shared__ __align__(16) float h[KP];
    __shared__ float part[2][2][NT / 32];
    __shared__ double scratch[32];
    float *Yt = kl_dyn, *aj = kl_dyn + (size_t)KP * cap, *av = aj + cap;
    unsigned long long its = 0;
    double klp = 0.0, sqp = 0.0;
    // `cols` columns of this launch: order[first .. first + cols) (columns sorted by length, so that one launch holds
    // columns of similar length and its staging buffers are sized for them), or first .. first + cols when order is null
    for (int t = blockIdx.x; t < cols; t += gridDim.x) {
        const int j = order ? order[first + t] : first + t;
        __syncthreads();
        if (threadIdx.x < KP) h[threadIdx.x] = X[(size_t)j * KP + threadIdx.x];
        __syncthreads();
        const int b = ptr[j], nz = ptr[j + 1] - b;
        int it;
        if (nz <= cap)
            it = kl_column<KP, NT, true>(idx + b, val + b, ajt + b, Y, sumY, b0, b1, b2, k, nz, max_iter, rel_tol, cap, Yt, aj,
                                         av, h, part, err, klp, sqp);
        else
            it = kl_column<KP, NT, false>(idx + b, val + b, ajt + b, Y, sumY, b0, b1, b2, k, nz, max_iter, rel_tol, cap, Yt, aj,
                                          av, h, part, err, klp, sqp);
        its += (unsigned long long)it;
        if (threadIdx.x < KP) X[(size_t)j * KP + threadIdx.x] = h[threadIdx.x];
    }
    if (threadIdx.x == 0 && its) atomicAdd(sweeps, its);
    if (err & KL_ERR) {
        const double t1 = block_sum(klp, scratch);
        if (threadIdx.x == 0) atomicAdd(kl_part, t1);
        const double t2 = block_sum(sqp, scratch);
        if (threadIdx.x == 0) atomicAdd(sq_part, t2);
    }
}

// ----------------------------------------------------------------------------------------------
// SCD KL update, one WARP per column (scd_kl_update) -- the kernel for columns of up to a few thousand
// nonzeros.  scd_kl_kernel stages the gathered rows of the fixed factor ((KP + 2) floats per nonzero), which caps an SM at
// 3-4 columns in flight, and pays a block-wide reduction per coordinate in every warp (ncu: 65 % issue slots at 137 warp
// instructions of overhead per warp and coordinate).  Here a column costs 12 bytes of shared memory per nonzero (A-hat, the
// value, the row index), the coordinate step reads column kk of the fixed factor from its factor-major copy Yt[kk][row]
// (the rows of a column ascend, so a warp's gather touches a few L1-resident lines), and the reduction is one butterfly
// with no barrier: 16-32 columns per SM in flight and ~3x fewer instructions per nonzero-coordinate.
//   X [cols][KP] in/out;  Y [rows][KP] the fixed factor, Yt [KP][ld] its factor-major copy;  the launcher guarantees
//   nz <= cap for every column of the launch;  err as in scd_kl_kernel
// ----------------------------------------------------------------------------------------------
constexpr int KLW_THREADS = 128;
template <int KP>
__global__ void __launch_bounds__(KLW_THREADS) scd_kl_warp_kernel(
    const int *__restrict__ ptr, const int *__restrict__ idx, const float *__restrict__ val, float *__restrict__ X,
    const float *__restrict__ Y, const float *__restrict__ Yt, size_t ld, const float *__restrict__ sumY, float b0, float b1,
    float b2, int k, int count, int max_iter, float rel_tol, unsigned long long *sweeps, double *kl_part, double *sq_part,
    int err, int cap, const int *__restrict__ order, int first)
{
    extern __shared__ __align__(16) float klw_dyn[];   // per warp: aj[cap] | av[cap] | ix[cap] | h[KP]
    constexpr int WPB = KLW_THREADS / 32;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *aj = klw_dyn + (size_t)wib * (3 * (size_t)cap + KP);
    float *av = aj + cap;
    int *ix = reinterpret_cast<int *>(av + cap);
    float *hs = reinterpret_cast<float *>(ix + cap);
    const int gw = blockIdx.x * WPB + wib, nw = gridDim.x * WPB;
    unsigned long long its = 0;
    double klp = 0.0, sqp = 0.0;
    for (int t = gw; t < count; t += nw) {
        const int j = order ? order[first + t] : first + t;
        const int b = ptr[j], nz = ptr[j + 1] - b;
        __syncwarp();
        for (int f = lane; f < KP; f += 32) hs[f] = X[(size_t)j * KP + f];
        __syncwarp();
        float sumH = 0.f;
        for (int f = 0; f < k; ++f) sumH += hs[f];
        // stage the column: A-hat = (row of Y) . h on the support, the values, the row indices
        for (int q = lane; q < nz; q += 32) {
            const int r = idx[b + q];
            const float4 *y4 = reinterpret_cast<const float4 *>(Y + (size_t)r * KP);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < KP / 4; ++c) {
                const float4 y = __ldg(y4 + c);
                const float4 hh = *reinterpret_cast<const float4 *>(hs + 4 * c);
                s = fmaf(y.x, hh.x, s); s = fmaf(y.y, hh.y, s); s = fmaf(y.z, hh.z, s); s = fmaf(y.w, hh.w, s);
            }
            aj[q] = s; av[q] = val[b + q]; ix[q] = r;
        }
        __syncwarp();
        int it = 0;
        bool again = true;
        float d_prev = 0.f;      // the A-hat update of the previous coordinate, applied on the way into the next one
        int k_prev = 0;
        for (; it < max_iter && again; ++it) {
            again = false;
            for (int kk = 0; kk < k; ++kk) {
                const float hk = hs[kk];
                const float *__restrict__ yk = Yt + (size_t)kk * ld;
                float a2 = 0.f, bb = 0.f;
                if (d_prev != 0.f) {       // warp-uniform
                    const float *__restrict__ yp = Yt + (size_t)k_prev * ld;
#pragma unroll 4
                    for (int q = lane; q < nz; q += 32) {
                        const int r = ix[q];
                        // exact arithmetic keeps A-hat >= 0 (d >= -h_kk); fp32 rounding may not: clamp
                        const float a = fmaxf(fmaf(d_prev, __ldg(yp + r), aj[q]), 0.f);
                        aj[q] = a;
                        const float mu = __fdividef(__ldg(yk + r), a + TINY_NUM_F);
                        const float v = av[q];
                        a2 = fmaf(v * mu, mu, a2);
                        bb = fmaf(v, mu, bb);
                    }
                } else {
#pragma unroll 4
                    for (int q = lane; q < nz; q += 32) {
                        const float mu = __fdividef(__ldg(yk + ix[q]), aj[q] + TINY_NUM_F);
                        const float v = av[q];
                        a2 = fmaf(v * mu, mu, a2);
                        bb = fmaf(v, mu, bb);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                    bb += __shfl_xor_sync(0xffffffffu, bb, o);
                }
                bb -= sumY[kk];
                float a2h;     // the two readings of NNLM's step with an L2 penalty: see kl_column
                if (err & KL_L2_LATE) { a2h = a2 * hk; a2 += b0; }
                else { a2 += b0; a2h = a2 * hk; }
                bb += a2h - b2 - b1 * (sumH - hk);
                float tmp = bb / (a2 + TINY_NUM_F);
                if (!(tmp > 0.f)) tmp = 0.f;
                d_prev = tmp - hk;
                k_prev = kk;
                if (tmp != hk) {
                    again |= (2.f * fabsf(d_prev) > rel_tol * (tmp + hk + TINY_NUM_F));
                    sumH += d_prev;
                    if (lane == 0) hs[kk] = tmp;
                }
                __syncwarp();
            }
        }
        its += (unsigned long long)it;
        if (err & KL_ERR) {
            // A-hat recomputed from the final h (see kl_column)
            for (int q = lane; q < nz; q += 32) {
                const float *y = Y + (size_t)ix[q] * KP;
                float s = 0.f;
                for (int f = 0; f < k; ++f) s = fmaf(__ldg(y + f), hs[f], s);
                const double a = (double)av[q], ah = (double)s;
                klp += a * log(ah + TINY_NUM_D);
                sqp += a * a - 2.0 * a * ah;
            }
        }
        __syncwarp();
        for (int f = lane; f < KP; f += 32) X[(size_t)j * KP + f] = hs[f];
    }
    if (lane == 0 && its) atomicAdd(sweeps, its);
    if (err & KL_ERR) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            klp += __shfl_xor_sync(0xffffffffu, klp, o);
            sqp += __shfl_xor_sync(0xffffffffu, sqp, o);
        }
        if (lane == 0) { atomicAdd(kl_part, klp); atomicAdd(sq_part, sqp); }
    }
}

// ----------------------------------------------------------------------------------------------
// SCD KL update of the gene side when the cells are sharded over ranks (scd_kl_update).
// A gene's row of A is spread over all ranks, so one coordinate step is split at the point where the Newton sums are
// complete:   kl_shard_accum (the rank's part of sum a mu^2, sum a mu on its own nonzeros, with the A-hat update of the
// previous coordinate applied lazily, exactly as kl_column does)  ->  all-reduce of the 2 n partials  ->
// kl_shard_update (the step itself, replicated: every rank gets the same W).  A-hat lives in the nnz-sized scratch of the
// by-gene view.  active[0..cols) : the gene is still sweeping;  active[cols..2 cols) : "again" flag of the running sweep.
// ----------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) kl_shard_init_kernel(const int *__restrict__ ptr, const int *__restrict__ idx,
                                                            float *__restrict__ ajt, const float *__restrict__ X,
                                                            const float *__restrict__ Y, int k, int cols,
                                                            float *__restrict__ dprev, float *__restrict__ sumX,
                                                            int *__restrict__ active)
{
    __shared__ __align__(16) float h[KP];
    for (int j = blockIdx.x; j < cols; j += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < KP) h[threadIdx.x] = X[(size_t)j * KP + threadIdx.x];
        __syncthreads();
        const int b = ptr[j], nz = ptr[j + 1] - b;
        for (int q = threadIdx.x; q < nz; q += 256) {
            const float4 *y4 = reinterpret_cast<const float4 *>(Y + (size_t)idx[b + q] * KP);
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < KP / 4; ++r) {
                const float4 y = __ldg(y4 + r);
                const float4 hh = *reinterpret_cast<const float4 *>(h + 4 * r);
                s = fmaf(y.x, hh.x, s); s = fmaf(y.y, hh.y, s); s = fmaf(y.z, hh.z, s); s = fmaf(y.w, hh.w, s);
            }
            ajt[b + q] = s;
        }
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int f = 0; f < k; ++f) s += h[f];
            sumX[j] = s; dprev[j] = 0.f; active[j] = 1; active[cols + j] = 0;
        }
    }
}

// Yt[f][r] = Y[r][f]: the fixed factor in factor-major order (row stride ld).  A coordinate step only needs column kk of Y,
// and the nonzeros of a by-gene row come in ascending cell order, so the gathers of the step walk one dense vector almost
// in order (a few 128-byte lines per warp) instead of touching one 4 KP-byte row of Y per nonzero.
template <int KP>
__global__ void __launch_bounds__(256) kl_shard_transpose_kernel(const float *__restrict__ Y, int rows, float *__restrict__ Yt,
                                                                 size_t ld)
{
    __shared__ float tile[32][KP + 1];
    const int r0 = blockIdx.x * 32;
    for (int t = threadIdx.x; t < 32 * KP; t += 256) {
        const int r = t / KP, f = t % KP;
        tile[r][f] = (r0 + r < rows) ? Y[(size_t)(r0 + r) * KP + f] : 0.f;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 32 * KP; t += 256) {
        const int f = t / 32, r = t % 32;
        if (r0 + r < rows) Yt[(size_t)f * ld + r0 + r] = tile[r][f];
    }
}

template <int KP>
__global__ void __launch_bounds__(256) kl_shard_accum_kernel(const int *__restrict__ ptr, const int *__restrict__ idx,
                                                             const float *__restrict__ val, float *__restrict__ ajt,
                                                             const float *__restrict__ Yt, size_t ld,
                                                             const float *__restrict__ dprev,
                                                             const int *__restrict__ active, int kk, int kprev, int cols,
                                                             float2 *__restrict__ part)
{
    const float *__restrict__ yk = Yt + (size_t)kk * ld;
    const float *__restrict__ yp = Yt + (size_t)kprev * ld;
    __shared__ float red[2][8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int j = blockIdx.x; j < cols; j += gridDim.x) {
        if (!active[j]) {                       // block-uniform: the gene left the sweeps, its partials are zeros
            if (tid == 0) part[j] = make_float2(0.f, 0.f);
            continue;
        }
        const int b = ptr[j], nz = ptr[j + 1] - b;
        const float d = dprev[j];
        float a2 = 0.f, bb = 0.f;
        for (int q = tid; q < nz; q += 256) {
            const int r = idx[b + q];
            float a = ajt[b + q];
            if (d != 0.f) {
                a = fmaxf(fmaf(d, __ldg(yp + r), a), 0.f);
                ajt[b + q] = a;
            }
            const float w = __ldg(yk + r), v = val[b + q];
            const float mu = __fdividef(w, a + TINY_NUM_F);
            a2 = fmaf(v * mu, mu, a2);
            bb = fmaf(v, mu, bb);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            bb += __shfl_xor_sync(0xffffffffu, bb, o);
        }
        if (lane == 0) { red[0][wid] = a2; red[1][wid] = bb; }
        __syncthreads();
        if (tid == 0) {
            float s2 = 0.f, sb = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) { s2 += red[0][w8]; sb += red[1][w8]; }
            part[j] = make_float2(s2, sb);
        }
        __syncthreads();   // red[] is reused by the next gene of this block
    }
}

// one thread per gene: the Newton step of coordinate kk from the summed partials (same arithmetic as kl_column)
template <int KP>
__global__ void __launch_bounds__(128) kl_shard_update_kernel(const float2 *__restrict__ part, float *__restrict__ X,
                                                              const float *__restrict__ sumY, float *__restrict__ sumX,
                                                              float *__restrict__ dprev, int *__restrict__ active, float b0,
                                                              float b1, float b2, int kk, int cols, float rel_tol, int flags)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols || !active[j]) return;
    const float hk = X[(size_t)j * KP + kk];
    const float2 pt = part[j];
    float a2 = pt.x, bb = pt.y - sumY[kk];
    float a2h;
    if (flags & KL_L2_LATE) { a2h = a2 * hk; a2 += b0; }
    else { a2 += b0; a2h = a2 * hk; }
    const float sH = sumX[j];
    bb += a2h - b2 - b1 * (sH - hk);
    float tmp = bb / (a2 + TINY_NUM_F);
    if (!(tmp > 0.f)) tmp = 0.f;
    const float d = tmp - hk;
    dprev[j] = d;
    if (tmp != hk) {
        if (2.f * fabsf(d) > rel_tol * (tmp + hk + TINY_NUM_F)) active[cols + j] = 1;
        sumX[j] = sH + d;
        X[(size_t)j * KP + kk] = tmp;
    }
}

// end of a sweep: count it for every gene that ran it, carry "again" into "active"; counts[0] += sweeps, counts[1] += genes
// that go on   (a template only because this header is compiled into several translation units)
template <int KP>
__global__ void kl_shard_sweep_end_kernel(int *__restrict__ active, int cols, unsigned long long *__restrict__ sweeps,
                                          unsigned *__restrict__ n_active)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols || !active[j]) return;
    atomicAdd(sweeps, 1ull);
    const int again = active[cols + j];
    active[j] = again;
    active[cols + j] = 0;
    if (again) atomicAdd(n_active, 1u);
}

// A-hat on the support only: kl_part += sum_nz a log(ahat + tiny)   (the mkl trace of an mse fit).
// Works on the dense rows directly (no sparse view to build).  A thread owns one gene and keeps that gene's row of
// Wt in registers; a block walks KL_CELLS cells whose h vectors sit in shared memory (broadcast reads), so the
// only streamed operand is A itself (coalesced 128 B per warp and cell) -- no per-nonzero gather of W rows.
constexpr int KL_CELLS = 128;
template <int KP>
__global__ void __launch_bounds__(256) kl_trace_kernel(const float *__restrict__ A, size_t ldA, int n, int m,
                                                       const float *__restrict__ H, const float *__restrict__ Wt, int k,
                                                       double *kl_part)
{
    __shared__ __align__(16) float sh[KL_CELLS][KP];
    __shared__ double scratch[32];
    const int g = blockIdx.x * 256 + threadIdx.x;
    const int c0 = blockIdx.y * KL_CELLS;
    const int nc = min(KL_CELLS, m - c0);
    for (int t = threadIdx.x; t < nc * KP; t += 256) sh[t / KP][t % KP] = H[(size_t)c0 * KP + t];
    float w[KP];
#pragma unroll
    for (int r = 0; r < KP / 4; ++r) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < n) v = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)g * KP) + r);
        w[4 * r] = v.x; w[4 * r + 1] = v.y; w[4 * r + 2] = v.z; w[4 * r + 3] = v.w;
    }
    __syncthreads();
    double klp = 0.0;
    if (g < n) {
        const float *ap = A + (size_t)c0 * ldA + g;
#pragma unroll 4
        for (int c = 0; c < nc; ++c) {
            const float av = __ldg(ap + (size_t)c * ldA);
            if (av != 0.f) {
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < KP / 4; ++r) {
                    const float4 hh = *reinterpret_cast<const float4 *>(&sh[c][4 * r]);
                    s = fmaf(w[4 * r], hh.x, s); s = fmaf(w[4 * r + 1], hh.y, s);
                    s = fmaf(w[4 * r + 2], hh.z, s); s = fmaf(w[4 * r + 3], hh.w, s);
                }
                // fp32 log (1 ulp): a warp almost always holds a nonzero lane, so a double-precision log here is paid
                // for every entry of A and was 80 % of this kernel; the sum itself stays in double
                klp += (double)av * (double)logf(s + TINY_NUM_F);
            }
        }
    }
    const double t1 = block_sum(klp, scratch);
    if (threadIdx.x == 0) atomicAdd(kl_part, t1);
}

// ----------------------------------------------------------------------------------------------
// dense reconstruction errors of FindNumFactors (R/nmf.R:170-179): sum (ahat - a)^2 and
// sum kl_div(ahat, a) with kl_div(x,y) = x log(x/y) - x + y on x+1e-12, y+1e-12 (R/nmf.R:71-75)
// one block per cell, threads over genes
// ----------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) recon_error_kernel(const float *__restrict__ A, size_t ldA, int n, int m,
                                                          const float *__restrict__ Wt, const float *__restrict__ H, int k,
                                                          double *out2)
{
    __shared__ float sh[KP];
    __shared__ double scratch[32];
    double se = 0.0, kl = 0.0;
    for (int j = blockIdx.x; j < m; j += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < KP) sh[threadIdx.x] = H[(size_t)j * KP + threadIdx.x];
        __syncthreads();
        for (int g = threadIdx.x; g < n; g += blockDim.x) {
            const float4 *w4 = reinterpret_cast<const float4 *>(Wt + (size_t)g * KP);
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < KP / 4; ++r) {
                const float4 w = w4[r];
                s = fmaf(w.x, sh[4 * r], s); s = fmaf(w.y, sh[4 * r + 1], s);
                s = fmaf(w.z, sh[4 * r + 2], s); s = fmaf(w.w, sh[4 * r + 3], s);
            }
            const double a = (double)A[(size_t)j * ldA + g], ah = (double)s;
            se += (ah - a) * (ah - a);
            const double x = ah + 1e-12, y = a + 1e-12;
            kl += x * log(x / y) - x + y;
        }
    }
    const double t1 = block_sum(se, scratch);
    if (threadIdx.x == 0) atomicAdd(&out2[0], t1);
    const double t2 = block_sum(kl, scratch);
    if (threadIdx.x == 0) atomicAdd(&out2[1], t2);
}

// X [rows][KP] <- X * T (T l x l, row-major double, applied in fp32)
template <int KP>
__global__ void right_multiply_kernel(float *X, size_t rows, const float *__restrict__ T, int l)
{
    __shared__ float sT[KP * KP];
    for (int t = threadIdx.x; t < KP * KP; t += blockDim.x) sT[t] = T[t];
    __syncthreads();
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float x[KP], y[KP];
#pragma unroll
    for (int f = 0; f < KP; ++f) { x[f] = X[r * KP + f]; y[f] = 0.f; }
#pragma unroll
    for (int a = 0; a < KP; ++a)
        if (a < l)
#pragma unroll
            for (int b = 0; b < KP; ++b) y[b] = fmaf(x[a], sT[a * KP + b], y[b]);
#pragma unroll
    for (int f = 0; f < KP; ++f) X[r * KP + f] = y[f];
}


}  // namespace swne
