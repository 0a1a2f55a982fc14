// gs_bench.cu -- stand-alone micro-benchmark of the per-tile SCD solver round trip (round 2).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/micro/gs_bench tools/micro/gs_bench.cu
//   tools/micro/gs_bench            (on a B200)
//
// One CTA of 128 threads solves one tile of 128 columns (k = 30 padded to 32), exactly the job of a solver group of
// the fused pass, but with nothing else on the SM: no TMA stream, no streaming MMAs, no drain warps.  G = CTAs
// resident per SM is varied through the grid size (148 * G tiles, 1 wave).  Variants:
//   0  block Gauss-Seidel, 2 blocks of 16, deltas through TMEM (tcgen05.st), A operand of the update MMA from TMEM
//      (= swne::tc_block_gs of pass_tc.cuh, restated here with optional clock64 instrumentation)
//   1  same, deltas through shared memory (st.shared + fence.proxy.async), A operand from smem
//   2  ONE block of 32 per sweep (lazy cross-sub-block corrections in registers), deltas through TMEM
//   3  one block of 32 per sweep, deltas through shared memory
//   4  CUDA-core sweeps, Gram in the constant bank (swne::scd_ls_solve_kernel<32>), 128 threads per CTA
//   5  same with 32 threads per CTA
// Prints per variant and G: total time, SM-cycles per sweep per tile, and for variants 0-3 the cycle breakdown of
// one thread (wait for the MMA, TMEM load, coordinate steps, delta store, barrier, MMA issue).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../swne_b200/csrc/pass_tc.cuh"
#include "../../swne_b200/csrc/kernels_misc.cuh"

namespace swne {
void set_error(const char *fmt, ...) { (void)fmt; }
}
using namespace swne;

#define CK(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } \
    } while (0)

constexpr int NPH = 8;

__device__ __forceinline__ long long clk() { return clock64(); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// A operand [128 rows x 8] tf32, K-major, no swizzle: 8-row x 16 B core matrices, K halves 128 B apart, row groups 256 B
__device__ __forceinline__ uint32_t aop_off(int row, int kq) { return (row >> 3) * 256 + (kq >> 2) * 128 + (row & 7) * 16 + (kq & 3) * 4; }
constexpr int AOP = 4096;   // bytes of one [128 x 8] operand

template <int VAR, bool INSTR, bool ELECT = false, bool FASTD = false>
__device__ __forceinline__ int gs_sweeps(float (&h)[32], uint32_t trow, uint32_t tmu, uint32_t tdelta, uint32_t gop_u, uint32_t sdelta,
                                         uint8_t *sdelta_p, uint32_t mbar, int quarter, int lane, int k, int max_iter, float rel_tol,
                                         long long (&ph)[NPH], int &loops)
{
    constexpr uint32_t ID_TF32 = idesc_tf32(128, 32);
    const bool any_change = rel_tol < 2.9e-8f;
    const int row = quarter * 32 + lane;
    int my_sweeps = 1;
    unsigned any = 1u;
    uint32_t mb_parity = 0;
    bool pending = false;
    long long t0 = 0, t1 = 0;
    loops = 0;
    for (int sweep = 0; sweep < max_iter && any; ++sweep) {
        unsigned changed = 0u;
        ++loops;
        if constexpr (VAR == 0 || VAR == 1) {
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                if (blk * 16 < k) {
                    if (INSTR) t0 = clk();
                    if (pending) { mbar_wait(mbar, mb_parity); mb_parity ^= 1u; pending = false; }
                    tc_fence_after();
                    if (INSTR) { t1 = clk(); ph[0] += t1 - t0; t0 = t1; }
                    float m16[16], d[16];
                    tc_ld16(trow + 16 * blk, m16);
                    if (INSTR) { t1 = clk(); ph[1] += t1 - t0; t0 = t1; }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int kk = 16 * blk + i;
                        if constexpr (FASTD) {
                            // delta = max(h - mu, 0) - h = max(-mu, -h): one op on the dependent chain instead of three
                            d[i] = fmaxf(-m16[i], -h[kk]);
                            const float tmp = h[kk] + d[i];
                            if (any_change) changed |= (tmp != h[kk]) ? 1u : 0u;
                            else changed |= (2.f * fabsf(d[i]) > rel_tol * (tmp + h[kk] + TINY_NUM_F)) ? 1u : 0u;
                            h[kk] = tmp;
                        } else {
                        const float tmp = fmaxf(h[kk] - m16[i], 0.f);
                        d[i] = tmp - h[kk];
                        if (any_change) changed |= __float_as_uint(d[i]);
                        else changed |= (2.f * fabsf(d[i]) > rel_tol * (tmp + h[kk] + TINY_NUM_F)) ? 1u : 0u;
                        h[kk] = tmp;
                        }
#pragma unroll
                        for (int j = i + 1; j < 16; ++j) m16[j] = fmaf(d[i], c_gram[kk * TC_KP + 16 * blk + j], m16[j]);
                    }
                    if (INSTR) { t1 = clk(); ph[2] += t1 - t0; t0 = t1; }
                    if constexpr (VAR == 0) {
                        float hl[32];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            hl[i] = __uint_as_float(__float_as_uint(d[i]) & 0xffffe000u);
                            hl[16 + i] = d[i] - hl[i];
                        }
                        tc_st32(tdelta + ((uint32_t)(quarter * 32) << 16), hl);
                        tc_fence_before();
                    } else {
                        // [ks][hi|lo] operands of 4 KB each
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                            for (int q4 = 0; q4 < 2; ++q4) {
                                float4 hi, lo;
                                float *hp = &hi.x, *lp = &lo.x;
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float dv = d[8 * ks + 4 * q4 + e];
                                    hp[e] = __uint_as_float(__float_as_uint(dv) & 0xffffe000u);
                                    lp[e] = dv - hp[e];
                                }
                                *reinterpret_cast<float4 *>(sdelta_p + (2 * ks) * AOP + aop_off(row, 4 * q4)) = hi;
                                *reinterpret_cast<float4 *>(sdelta_p + (2 * ks + 1) * AOP + aop_off(row, 4 * q4)) = lo;
                            }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    }
                    if (INSTR) { t1 = clk(); ph[3] += t1 - t0; t0 = t1; }
                    named_bar_sync(1, 128);
                    if (INSTR) { t1 = clk(); ph[4] += t1 - t0; t0 = t1; }
                    if (quarter == 0 && (ELECT ? elect_one() : lane == 0)) {
                        tc_fence_after();
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const int b8 = 2 * blk + ks;
                            const uint64_t g_hi = desc_nosw(gop_u + (2 * b8) * TC_GOP, 128, 256);
                            const uint64_t g_lo = desc_nosw(gop_u + (2 * b8 + 1) * TC_GOP, 128, 256);
                            if constexpr (VAR == 0) {
                                tc_mma_tf32_ts(tmu, tdelta + 8 * ks, g_hi, ID_TF32, 1);
                                tc_mma_tf32_ts(tmu, tdelta + 8 * ks, g_lo, ID_TF32, 1);
                                tc_mma_tf32_ts(tmu, tdelta + 16 + 8 * ks, g_hi, ID_TF32, 1);
                            } else {
                                const uint64_t a_hi = desc_nosw(sdelta + (2 * ks) * AOP, 128, 256);
                                const uint64_t a_lo = desc_nosw(sdelta + (2 * ks + 1) * AOP, 128, 256);
                                tc_mma_tf32(tmu, a_hi, g_hi, ID_TF32, 1);
                                tc_mma_tf32(tmu, a_hi, g_lo, ID_TF32, 1);
                                tc_mma_tf32(tmu, a_lo, g_hi, ID_TF32, 1);
                            }
                        }
                        tc_commit(mbar);
                    }
                    pending = true;
                    if (INSTR) { t1 = clk(); ph[5] += t1 - t0; t0 = t1; }
                }
            }
        } else {
            // ---- one block of 32 per sweep: four sub-blocks of 8; the corrections of a sub-block's mu by the deltas of the
            // earlier sub-blocks of this sweep are applied lazily when the sub-block starts
            if (INSTR) t0 = clk();
            if (pending) { mbar_wait(mbar, mb_parity); mb_parity ^= 1u; pending = false; }
            tc_fence_after();
            if (INSTR) { t1 = clk(); ph[0] += t1 - t0; t0 = t1; }
            float d[32];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (8 * s < k) {
                    float m8[8];
                    tc_ld8(trow + 8 * s, m8);
                    if (INSTR) { t1 = clk(); ph[1] += t1 - t0; t0 = t1; }
#pragma unroll
                    for (int i = 0; i < 8 * s; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) m8[j] = fmaf(d[i], c_gram[i * TC_KP + 8 * s + j], m8[j]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int kk = 8 * s + i;
                        const float tmp = fmaxf(h[kk] - m8[i], 0.f);
                        d[kk] = tmp - h[kk];
                        if (any_change) changed |= __float_as_uint(d[kk]);
                        else changed |= (2.f * fabsf(d[kk]) > rel_tol * (tmp + h[kk] + TINY_NUM_F)) ? 1u : 0u;
                        h[kk] = tmp;
#pragma unroll
                        for (int j = i + 1; j < 8; ++j) m8[j] = fmaf(d[kk], c_gram[kk * TC_KP + 8 * s + j], m8[j]);
                    }
                    if (INSTR) { t1 = clk(); ph[2] += t1 - t0; t0 = t1; }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) d[8 * s + i] = 0.f;
                }
            }
            if constexpr (VAR == 2) {
                // TMEM operand columns: [0,32) hi, [32,64) lo
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        hi[i] = __uint_as_float(__float_as_uint(d[16 * half + i]) & 0xffffe000u);
                        lo[i] = d[16 * half + i] - hi[i];
                    }
                    float a8[8], b8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { a8[i] = hi[i]; b8[i] = hi[8 + i]; }
                    tc_st16(tdelta + 16 * half + ((uint32_t)(quarter * 32) << 16), a8, b8);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { a8[i] = lo[i]; b8[i] = lo[8 + i]; }
                    tc_st16(tdelta + 32 + 16 * half + ((uint32_t)(quarter * 32) << 16), a8, b8);
                }
                tc_fence_before();
            } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                    for (int q4 = 0; q4 < 2; ++q4) {
                        float4 hi, lo;
                        float *hp = &hi.x, *lp = &lo.x;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float dv = d[8 * ks + 4 * q4 + e];
                            hp[e] = __uint_as_float(__float_as_uint(dv) & 0xffffe000u);
                            lp[e] = dv - hp[e];
                        }
                        *reinterpret_cast<float4 *>(sdelta_p + (2 * ks) * AOP + aop_off(row, 4 * q4)) = hi;
                        *reinterpret_cast<float4 *>(sdelta_p + (2 * ks + 1) * AOP + aop_off(row, 4 * q4)) = lo;
                    }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            if (INSTR) { t1 = clk(); ph[3] += t1 - t0; t0 = t1; }
            named_bar_sync(1, 128);
            if (INSTR) { t1 = clk(); ph[4] += t1 - t0; t0 = t1; }
            if (quarter == 0 && lane == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (8 * ks < k) {
                        const uint64_t g_hi = desc_nosw(gop_u + (2 * ks) * TC_GOP, 128, 256);
                        const uint64_t g_lo = desc_nosw(gop_u + (2 * ks + 1) * TC_GOP, 128, 256);
                        if constexpr (VAR == 2) {
                            tc_mma_tf32_ts(tmu, tdelta + 8 * ks, g_hi, ID_TF32, 1);
                            tc_mma_tf32_ts(tmu, tdelta + 8 * ks, g_lo, ID_TF32, 1);
                            tc_mma_tf32_ts(tmu, tdelta + 32 + 8 * ks, g_hi, ID_TF32, 1);
                        } else {
                            const uint64_t a_hi = desc_nosw(sdelta + (2 * ks) * AOP, 128, 256);
                            const uint64_t a_lo = desc_nosw(sdelta + (2 * ks + 1) * AOP, 128, 256);
                            tc_mma_tf32(tmu, a_hi, g_hi, ID_TF32, 1);
                            tc_mma_tf32(tmu, a_hi, g_lo, ID_TF32, 1);
                            tc_mma_tf32(tmu, a_lo, g_hi, ID_TF32, 1);
                        }
                    }
                }
                tc_commit(mbar);
            }
            pending = true;
            if (INSTR) { t1 = clk(); ph[5] += t1 - t0; t0 = t1; }
        }
        if (INSTR) t0 = clk();
        any = named_bar_or(1, 128, changed);
        if (INSTR) { t1 = clk(); ph[6] += t1 - t0; }
        if (changed) my_sweeps = min(sweep + 2, max_iter);
    }
    if (pending) { mbar_wait(mbar, mb_parity); pending = false; }
    tc_fence_after();
    return my_sweeps;
}

template <int VAR, bool INSTR, bool ELECT = false, bool FASTD = false>
__global__ void __launch_bounds__(128, 4)
gs_kernel(float *__restrict__ X, const float *__restrict__ C, const float *__restrict__ pack, float l1, int k, int cols,
          int max_iter, float rel_tol, unsigned long long *sweeps, unsigned long long *tile_loops, long long *phases)
{
    constexpr int TMEM_COLS = (VAR == 0) ? 64 : (VAR == 2) ? 128 : 32;
    extern __shared__ uint8_t dyn_raw[];
    uint8_t *dyn = (uint8_t *)(((uintptr_t)dyn_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *gops = dyn;                       // 8 KB
    uint8_t *sdel = dyn + 8 * TC_GOP;          // up to 8 operands of 4 KB
    __shared__ float sD[2 * TC_KP];
    __shared__ __align__(8) uint64_t mbar_store;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, quarter = warp;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const uint32_t mbar = smem_u32(&mbar_store);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fill_gops(gops, pack, threadIdx.x, blockDim.x);
    if (threadIdx.x < 2 * TC_KP) sD[threadIdx.x] = pack[2 * TC_KP * TC_KP + threadIdx.x];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const float *sInvD = sD + TC_KP;
    const int j = blockIdx.x * 128 + threadIdx.x;
    const bool valid = j < cols;
    float *xrow = X + (size_t)(valid ? j : 0) * TC_KP;
    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
    float h[32];
    {
        float mv[32];
        float2 mu[16];
        if (valid) {
            const float4 *cp = reinterpret_cast<const float4 *>(C + (size_t)j * TC_KP);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 v = reinterpret_cast<const float4 *>(xrow)[r];
                h[4 * r] = v.x * sD[4 * r]; h[4 * r + 1] = v.y * sD[4 * r + 1];
                h[4 * r + 2] = v.z * sD[4 * r + 2]; h[4 * r + 3] = v.w * sD[4 * r + 3];
                const float4 c = cp[r];
                mu[2 * r] = make_float2((l1 - c.x) * sInvD[4 * r], (l1 - c.y) * sInvD[4 * r + 1]);
                mu[2 * r + 1] = make_float2((l1 - c.z) * sInvD[4 * r + 2], (l1 - c.w) * sInvD[4 * r + 3]);
            }
            scd_mu_init<TC_KP>(mu, xrow, sD, k);
        } else {
#pragma unroll
            for (int f = 0; f < 32; ++f) h[f] = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) mu[r] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) { mv[2 * r] = mu[r].x; mv[2 * r + 1] = mu[r].y; }
        tc_st32(trow, mv);
    }
    long long ph[NPH];
#pragma unroll
    for (int i = 0; i < NPH; ++i) ph[i] = 0;
    int loops = 0;
    const long long tb = clk();
    const int my_sweeps = gs_sweeps<VAR, INSTR, ELECT, FASTD>(h, trow, tmem, tmem + 32, smem_u32(gops), smem_u32(sdel), sdel, mbar, quarter, lane, k,
                                                max_iter, rel_tol, ph, loops);
    ph[7] = clk() - tb;
    if (valid) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
            reinterpret_cast<float4 *>(xrow)[r] = make_float4(h[4 * r] * sInvD[4 * r], h[4 * r + 1] * sInvD[4 * r + 1],
                                                              h[4 * r + 2] * sInvD[4 * r + 2], h[4 * r + 3] * sInvD[4 * r + 3]);
    }
    int sw = valid ? my_sweeps : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sw += __shfl_xor_sync(0xffffffffu, sw, o);
    if (lane == 0) atomicAdd(sweeps, (unsigned long long)sw);
    if (threadIdx.x == 0) atomicAdd(tile_loops, (unsigned long long)loops);
    if (threadIdx.x == 0 && phases) {
        for (int i = 0; i < NPH; ++i) atomicAdd((unsigned long long *)&phases[i], (unsigned long long)ph[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

// CUDA-core sweeps with plain FFMA and constant-bank operands (no LDCU -> UR funnel): thread per column
template <int NT>
__global__ void __launch_bounds__(NT) simt_cbank_kernel(float *__restrict__ X, const float *__restrict__ C, const float *__restrict__ pack,
                                                        float l1, int k, int cols, int max_iter, float rel_tol, unsigned long long *sweeps)
{
    constexpr int KP = 32;
    __shared__ float sD[KP], sInvD[KP];
    for (int t = threadIdx.x; t < KP; t += blockDim.x) { sD[t] = pack[2 * KP * KP + t]; sInvD[t] = pack[2 * KP * KP + KP + t]; }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    float h[KP];
    float2 mu[KP / 2];
    float *xrow = X + (size_t)j * KP;
    const float4 *cp = reinterpret_cast<const float4 *>(C + (size_t)j * KP);
#pragma unroll
    for (int r = 0; r < KP / 4; ++r) {
        const float4 v = reinterpret_cast<const float4 *>(xrow)[r];
        h[4 * r] = v.x * sD[4 * r]; h[4 * r + 1] = v.y * sD[4 * r + 1]; h[4 * r + 2] = v.z * sD[4 * r + 2]; h[4 * r + 3] = v.w * sD[4 * r + 3];
        const float4 c = cp[r];
        mu[2 * r] = make_float2((l1 - c.x) * sInvD[4 * r], (l1 - c.y) * sInvD[4 * r + 1]);
        mu[2 * r + 1] = make_float2((l1 - c.z) * sInvD[4 * r + 2], (l1 - c.w) * sInvD[4 * r + 3]);
    }
    scd_mu_init<KP>(mu, xrow, sD, k);
    const int it = scd_ls_thread<KP, false>(h, mu, k, max_iter, rel_tol);
#pragma unroll
    for (int r = 0; r < KP / 4; ++r)
        reinterpret_cast<float4 *>(xrow)[r] = make_float4(h[4 * r] * sInvD[4 * r], h[4 * r + 1] * sInvD[4 * r + 1], h[4 * r + 2] * sInvD[4 * r + 2],
                                                          h[4 * r + 3] * sInvD[4 * r + 3]);
    atomicAdd(sweeps, (unsigned long long)it);
}

// CUDA-core sweeps, G' broadcast from shared memory (LDS.128), thread per column
template <int NT>
__global__ void __launch_bounds__(NT) simt_smem_kernel(float *__restrict__ X, const float *__restrict__ C, const float *__restrict__ pack,
                                                       float l1, int k, int cols, int max_iter, float rel_tol, unsigned long long *sweeps)
{
    constexpr int KP = 32;
    __shared__ __align__(16) float sG[KP * KP];
    __shared__ float sD[KP], sInvD[KP];
    for (int t = threadIdx.x; t < KP * KP; t += blockDim.x) sG[t] = pack[t];
    for (int t = threadIdx.x; t < KP; t += blockDim.x) { sD[t] = pack[2 * KP * KP + t]; sInvD[t] = pack[2 * KP * KP + KP + t]; }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    float h[1][KP];
    float2 mu[1][KP / 2];
    float *xrow = X + (size_t)j * KP;
    const float4 *cp = reinterpret_cast<const float4 *>(C + (size_t)j * KP);
#pragma unroll
    for (int r = 0; r < KP / 4; ++r) {
        const float4 v = reinterpret_cast<const float4 *>(xrow)[r];
        h[0][4 * r] = v.x * sD[4 * r]; h[0][4 * r + 1] = v.y * sD[4 * r + 1]; h[0][4 * r + 2] = v.z * sD[4 * r + 2]; h[0][4 * r + 3] = v.w * sD[4 * r + 3];
        const float4 c = cp[r];
        mu[0][2 * r] = make_float2((l1 - c.x) * sInvD[4 * r], (l1 - c.y) * sInvD[4 * r + 1]);
        mu[0][2 * r + 1] = make_float2((l1 - c.z) * sInvD[4 * r + 2], (l1 - c.w) * sInvD[4 * r + 3]);
    }
    scd_mu_init<KP>(mu[0], xrow, sD, k);
    int sw[1];
    scd_ls_thread_smem<KP, 1>(h, mu, (unsigned)__cvta_generic_to_shared(sG), k, max_iter, rel_tol, sw);
#pragma unroll
    for (int r = 0; r < KP / 4; ++r)
        reinterpret_cast<float4 *>(xrow)[r] = make_float4(h[0][4 * r] * sInvD[4 * r], h[0][4 * r + 1] * sInvD[4 * r + 1], h[0][4 * r + 2] * sInvD[4 * r + 2],
                                                          h[0][4 * r + 3] * sInvD[4 * r + 3]);
    atomicAdd(sweeps, (unsigned long long)sw[0]);
}

struct Problem {
    int cols;
    std::vector<float> X, C, pack;
};

static Problem make_problem(int cols, int k, unsigned seed)
{
    // G = W'W of a sparse non-negative W (3000 x k), as in a W-side Gram of an NMF iterate; columns: c = G0 h* + noise with a
    // sparse non-negative h*, start from a positive random h (many sweeps needed, like the first outer iterations)
    const int KP = 32, n = 3000;
    std::mt19937 rng(seed);
    std::gamma_distribution<float> gam(0.4f, 1.f);
    std::uniform_real_distribution<float> uni(0.f, 1.f);
    std::normal_distribution<float> nor(0.f, 1.f);
    std::vector<double> G64(KP * KP + KP, 0.0);
    std::vector<float> W((size_t)n * KP, 0.f);
    // correlated factors: a shared component makes the Gram ill-conditioned (slow coordinate descent)
    for (int i = 0; i < n; ++i) {
        const float shared = gam(rng);
        for (int f = 0; f < k; ++f) W[(size_t)i * KP + f] = (uni(rng) < 0.5f ? gam(rng) : 0.f) + 0.6f * shared;
    }
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < k; ++a) {
            const double wa = W[(size_t)i * KP + a];
            G64[KP * KP + a] += wa;
            for (int b = 0; b < k; ++b) G64[a * KP + b] += wa * W[(size_t)i * KP + b];
        }
    Problem p;
    p.cols = cols;
    p.pack.assign(gram_pack_floats(KP), 0.f);
    std::vector<float> inv(KP);
    for (int t = 0; t < KP; ++t) {
        const double g = (t < k) ? G64[t * KP + t] + 1e-16 : 1.0;
        const float d = sqrtf((float)g);
        p.pack[2 * KP * KP + t] = d;
        p.pack[2 * KP * KP + KP + t] = 1.f / d;
        inv[t] = 1.f / d;
    }
    for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b) {
            p.pack[a * KP + b] = (a == b) ? 1.f : (float)G64[a * KP + b] * inv[a] * inv[b];
            p.pack[KP * KP + a * KP + b] = (float)G64[a * KP + b];
        }
    p.X.assign((size_t)cols * KP, 0.f);
    p.C.assign((size_t)cols * KP, 0.f);
    std::vector<float> hs(KP);
    for (int j = 0; j < cols; ++j) {
        for (int f = 0; f < KP; ++f) hs[f] = (f < k && uni(rng) < 0.4f) ? gam(rng) : 0.f;
        for (int a = 0; a < k; ++a) {
            double c = 0;
            for (int b = 0; b < k; ++b) c += G64[a * KP + b] * hs[b];
            c *= 1.0 + 0.3 * nor(rng);
            p.C[(size_t)j * KP + a] = (float)std::max(c, 0.0);
            p.X[(size_t)j * KP + a] = 0.5f * gam(rng) + 0.01f;
        }
    }
    return p;
}

int main(int argc, char **argv)
{
    const int k = 30, KP = 32;
    const int max_iter = argc > 1 ? atoi(argv[1]) : 50;
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    const double ghz = prop.clockRate * 1e-6;
    printf("device %s, %d SMs, %.3f GHz nominal\n", prop.name, sms, ghz);
    const int dyn = 8 * TC_GOP + 8 * AOP + 1024;
    CK(cudaFuncSetAttribute(gs_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute(gs_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute(gs_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute(gs_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute(gs_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute(gs_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute(gs_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute(gs_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute((gs_kernel<0, false, true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute((gs_kernel<0, true, true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute((gs_kernel<0, false, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CK(cudaFuncSetAttribute((gs_kernel<0, true, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));

    const int Gs[] = {1, 2, 3, 4, 4};
    const int maxcols = 128 * sms * 4;
    Problem P = make_problem(maxcols, k, 7);
    float *dX, *dX0, *dC, *dpack, *dXref;
    unsigned long long *dcnt;
    long long *dph;
    CK(cudaMalloc(&dX, sizeof(float) * (size_t)maxcols * KP));
    CK(cudaMalloc(&dX0, sizeof(float) * (size_t)maxcols * KP));
    CK(cudaMalloc(&dXref, sizeof(float) * (size_t)maxcols * KP));
    CK(cudaMalloc(&dC, sizeof(float) * (size_t)maxcols * KP));
    CK(cudaMalloc(&dpack, sizeof(float) * P.pack.size()));
    CK(cudaMalloc(&dcnt, 2 * sizeof(unsigned long long)));
    CK(cudaMalloc(&dph, NPH * sizeof(long long)));
    CK(cudaMemcpy(dX0, P.X.data(), sizeof(float) * P.X.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, P.C.data(), sizeof(float) * P.C.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dpack, P.pack.data(), sizeof(float) * P.pack.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(c_gram, P.pack.data(), sizeof(float) * 2 * KP * KP));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<float> ref((size_t)maxcols * KP), got((size_t)maxcols * KP);

    auto run = [&](int var, bool instr, int cols, float *ms, unsigned long long *cnt, long long *ph) {
        CK(cudaMemcpy(dX, dX0, sizeof(float) * (size_t)cols * KP, cudaMemcpyDeviceToDevice));
        CK(cudaMemset(dcnt, 0, 2 * sizeof(unsigned long long)));
        CK(cudaMemset(dph, 0, NPH * sizeof(long long)));
        const int tiles = (cols + 127) / 128;
        CK(cudaEventRecord(e0));
        switch (var * 2 + (instr ? 1 : 0)) {
#define LAUNCH(V, I) gs_kernel<V, I><<<tiles, 128, dyn>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt, dcnt + 1, dph)
            case 0: LAUNCH(0, false); break;
            case 1: LAUNCH(0, true); break;
            case 2: LAUNCH(1, false); break;
            case 3: LAUNCH(1, true); break;
            case 4: LAUNCH(2, false); break;
            case 5: LAUNCH(2, true); break;
            case 6: LAUNCH(3, false); break;
            case 7: LAUNCH(3, true); break;
            case 8: scd_ls_solve_kernel<32><<<tiles, 128>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt, nullptr); break;
            case 10: scd_ls_solve_kernel<32><<<(cols + 31) / 32, 32>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt, nullptr); break;
            case 12: simt_cbank_kernel<128><<<tiles, 128>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt); break;
            case 14: simt_cbank_kernel<32><<<(cols + 31) / 32, 32>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt); break;
            case 16: simt_smem_kernel<128><<<tiles, 128>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt); break;
            case 18: simt_smem_kernel<32><<<(cols + 31) / 32, 32>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt); break;
            case 20: gs_kernel<0, false, true, false><<<tiles, 128, dyn>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt, dcnt + 1, dph); break;
            case 21: gs_kernel<0, true, true, false><<<tiles, 128, dyn>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt, dcnt + 1, dph); break;
            case 22: gs_kernel<0, false, true, true><<<tiles, 128, dyn>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt, dcnt + 1, dph); break;
            case 23: gs_kernel<0, true, true, true><<<tiles, 128, dyn>>>(dX, dC, dpack, 0.f, k, cols, max_iter, -1.f, dcnt, dcnt + 1, dph); break;
            default: printf("bad variant\n"); exit(1);
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(ms, e0, e1));
        CK(cudaMemcpy(cnt, dcnt, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ph, dph, NPH * sizeof(long long), cudaMemcpyDeviceToHost));
    };

    // reference result: the CUDA-core solver on the largest problem
    {
        float ms; unsigned long long cnt[2]; long long ph[NPH];
        run(4, false, maxcols, &ms, cnt, ph);
        CK(cudaMemcpy(ref.data(), dX, sizeof(float) * ref.size(), cudaMemcpyDeviceToHost));
        printf("reference (CUDA-core sweeps): mean sweeps per column %.2f\n", (double)cnt[0] / maxcols);
    }
    const char *names[] = {"0 blk16 tmem-delta", "1 blk16 smem-delta", "2 blk32 tmem-delta", "3 blk32 smem-delta", "4 simt 128thr", "5 simt 32thr", "6 simt-cbank 128thr", "7 simt-cbank 32thr", "8 simt-smem 128thr", "9 simt-smem 32thr", "10 blk16 elect", "11 blk16 elect fastd"};
    for (int var = 0; var < 12; ++var) {
        for (int gi = 0; gi < 4; ++gi) {
            const int G = Gs[gi];
            const int cols = 128 * sms * G;
            float ms = 0, best = 1e30f;
            unsigned long long cnt[2];
            long long ph[NPH];
            for (int rep = 0; rep < 3; ++rep) { run(var, false, cols, &ms, cnt, ph); best = std::min(best, ms); }
            CK(cudaMemcpy(got.data(), dX, sizeof(float) * (size_t)cols * KP, cudaMemcpyDeviceToHost));
            double num = 0, den = 0;
            for (size_t i = 0; i < (size_t)cols * KP; ++i) { const double e = (double)got[i] - ref[i]; num += e * e; den += (double)ref[i] * ref[i]; }
            const double mean_sw = (double)cnt[0] / cols;
            // tile loops: variants 0-3 count loop iterations per tile; simt: assume max_iter per warp
            const double loops_per_tile = (var < 4 || var >= 10) ? (double)cnt[1] / (cols / 128) : (double)max_iter;
            printf("var %-20s G=%d  %8.1f us  mean sweeps/col %.2f  tile loops %.1f  us/sweep %.3f  rel diff vs simt %.2e\n", names[var], G,
                   best * 1e3, mean_sw, loops_per_tile, best * 1e3 / loops_per_tile, sqrt(num / den));
            if (var < 4 || var >= 10) {
                run(var, true, cols, &ms, cnt, ph);
                const double tl = (double)cnt[1];
                printf("      instrumented %8.1f us; cycles per sweep (thread 0 of each tile, mean): wait %.0f  ld %.0f  steps %.0f  store %.0f  bar %.0f  mma-issue %.0f  bar.or %.0f  total %.0f\n",
                       ms * 1e3, ph[0] / tl, ph[1] / tl, ph[2] / tl, ph[3] / tl, ph[4] / tl, ph[5] / tl, ph[6] / tl, ph[7] / tl);
            }
        }
    }
    // W-side size: 3000 columns
    for (int var = 0; var < 12; ++var) {
        float ms = 0, best = 1e30f;
        unsigned long long cnt[2];
        long long ph[NPH];
        for (int rep = 0; rep < 3; ++rep) { run(var, false, 3000, &ms, cnt, ph); best = std::min(best, ms); }
        printf("W-side 3000 cols: var %-20s %8.1f us  mean sweeps/col %.2f\n", names[var], best * 1e3, (double)cnt[0] / 3000);
    }
    return 0;
}
