// contract_tc.cuh -- tcgen05 contractions for 32 < k <= 64 (the "wide" variant of SWNE_ENGINE_TC).
//
// The fused pass of pass_tc.cuh keeps a tile's mu', the solver operands and the phase-3 accumulators in TMEM at once;
// that budget (512 columns) only works for a 32-column factor block.  For wider factorizations (cfg5: k = 40,
// FindNumFactors above k = 32) the two skinny contractions of update() still go to the tensor cores, un-fused:
//
//   tcx_cells_kernel   C = A Wt    (cells x 64)   M = 128 cells, N = 64, K = genes        -> global C (fp32)
//   tcx_genes_kernel   B += A' H   (genes x 64)   M = 128 genes (MN-major view), N = 64, K = cells -> red.add into B
//
// and the per-column NNLS solves run on the CUDA-core kernels (scd_ls_solve_kernel<64>).  Operands are the same
// split-fp16 pairs (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) on the same pre-split copy of A; W and H are
// re-split every iteration with exact power-of-two scales taken from their maxima, so there is no range guess.
// A is streamed twice per iteration (once by each kernel), like the fused pass.
//
// Reference semantics restated: the two products of update() in NNLM, call sites
// /root/reference/R/nmf.R:117,129.
#pragma once
#include "pass_tc.cuh"

namespace swne {

constexpr int TCX_KP = 64;
constexpr int TCX_MT_PER_PASS = 3;                        // 128-gene accumulators (64 TMEM columns each) per sub-pass, two regions = 384 columns
                                                          // (2 per sub-pass = 256 columns would let solve CTAs sit next to a genes CTA, but
                                                          // the extra sub-passes re-read the H images: measured slower, round 2)
constexpr int TCX_STAGES = 3;                             // 3 x (32 KB A + 16 KB W): leaves shared memory for co-resident solve CTAs (slab pipeline)
constexpr int TCX_W_STAGE = 2 * TCX_KP * TC_CHUNK * 2;   // 16 KB: rows 0..63 W_hi, 64..127 W_lo, 64 genes each
constexpr int TCX_H_SUB = TCX_KP * 64 * 2;               // 8 KB: [64 factors][64 cells] fp16, K-major
constexpr int TCX_H_TILE = 4 * TCX_H_SUB;                // hi cells 0..63, hi cells 64..127, lo, lo
constexpr int TCX_THREADS_CELLS = 6 * 32;                // producer, MMA, 4 epilogue warps
constexpr int TCX_THREADS_GENES = 8 * 32;                // producer, MMA, H loader, (idle), 4 drain warps

struct TcxSmemCells {
    static constexpr int a_hi = 0;
    static constexpr int a_lo = a_hi + TCX_STAGES * TC_A_STAGE;
    static constexpr int w = a_lo + TCX_STAGES * TC_A_STAGE;
    static constexpr int bars = w + TCX_STAGES * TCX_W_STAGE;
    static constexpr int n_bars = 2 * TCX_STAGES + 4;
    static constexpr int misc = bars + n_bars * 8;
    static constexpr int total = misc + 64;
};
struct TcxSmemGenes {
    static constexpr int a_hi = 0;
    static constexpr int a_lo = a_hi + TCX_STAGES * TC_A_STAGE;
    static constexpr int h = a_lo + TCX_STAGES * TC_A_STAGE;
    static constexpr int bars = h + 2 * TCX_H_TILE;
    static constexpr int n_bars = 2 * TCX_STAGES + 8;
    static constexpr int misc = bars + n_bars * 8;
    static constexpr int total = misc + 64;
};

struct TcxParams {
    int n, m;
    int ld;                // row stride (= padded factor count KP, a multiple of 8, <= 64) of C, B, Wt and H
    int n_chunks;          // 64-gene chunks of the split copy (n_pad / 64)
    int n_mtiles;          // ceil(n / 128)
    int tiles_total;       // number of 128-cell tiles this launch covers
    int tile0;             // first of them (a launch may cover a slab of the cells)
    float *C;              // [m][ld] fp32 out (cells kernel)
    float *B;              // [n][ld] fp32, accumulated with red.add (genes kernel; zeroed by the caller)
    const uint8_t *Himg;   // [tiles_total][TCX_H_TILE] split images of H (genes kernel)
    const int *x_exp;      // device: exponent the W (cells kernel) or H (genes kernel) operand was scaled by
    int a_exp;
    int regions;           // genes kernel: accumulator regions in TMEM.  2 (512 columns): the drain of one gene sub-pass overlaps the
                           // MMAs of the next; 1 (256 columns): leaves TMEM for two co-resident solve CTAs (slab pipeline), the
                           // drain bubble is negligible when a sub-pass covers many cells
};

__device__ __forceinline__ void tcx_my_tiles(int tiles_total, int tile0, int &t_begin, int &t_count)
{
    const int base = tiles_total / (int)gridDim.x, rem = tiles_total % (int)gridDim.x;
    t_begin = tile0 + (int)blockIdx.x * base + min((int)blockIdx.x, rem);
    t_count = base + ((int)blockIdx.x < rem ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------
// C = A Wt.  warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue (TMEM -> global), two accumulators
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TCX_THREADS_CELLS, 1)
tcx_cells_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                 const __grid_constant__ CUtensorMap map_w, const TcxParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t_begin, n_tiles;
    tcx_my_tiles(p.tiles_total, p.tile0, t_begin, n_tiles);

    const uint32_t bar0 = sbase + TcxSmemCells::bars;
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (TCX_STAGES + s); };
    auto CFULL = [&](int b) { return bar0 + 8u * (2 * TCX_STAGES + b); };
    auto CEMPTY = [&](int b) { return bar0 + 8u * (2 * TCX_STAGES + 2 + b); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcxSmemCells::misc);

    // rows of a partial tile that TMA never writes only feed C rows that are not stored, but keep them finite
    for (int i = threadIdx.x; i < (2 * TCX_STAGES * TC_A_STAGE) / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem + TcxSmemCells::a_hi)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < TCX_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(CFULL(b), 1); mbar_init(CEMPTY(b), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        int stage = 0, phase = 0;
        for (int i = 0; i < n_tiles; ++i) {
            const int cell0 = (t_begin + i) * TC_TILE;
            const int nbox = (min(TC_TILE, p.m - cell0) + TC_GROUP - 1) / TC_GROUP;
            for (int ch = 0; ch < p.n_chunks; ++ch) {
                if (lane == 0) {
                    mbar_wait(EMPTY(stage), phase ^ 1);
                    mbar_expect_tx(FULL(stage), nbox * 2 * TC_BOX_BYTES + TCX_W_STAGE);
                }
                __syncwarp();
                if (lane < 2 * nbox) {
                    const int b = lane >> 1;
                    tma_load_2d(sbase + ((lane & 1) ? TcxSmemCells::a_lo : TcxSmemCells::a_hi) + stage * TC_A_STAGE + b * TC_BOX_BYTES,
                                (lane & 1) ? &map_alo : &map_ahi, ch * TC_CHUNK, cell0 + b * TC_GROUP, FULL(stage));
                } else if (lane == 2 * nbox) {
                    tma_load_2d(sbase + TcxSmemCells::w + stage * TCX_W_STAGE, &map_w, ch * TC_CHUNK, 0, FULL(stage));
                }
                if (++stage == TCX_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t ID = idesc_f16(128, TCX_KP, 0, 0, 0);
            int stage = 0, phase = 0;
            for (int i = 0; i < n_tiles; ++i) {
                const int b = i & 1;
                mbar_wait(CEMPTY(b), ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t dC = tmem + b * TCX_KP;
                for (int ch = 0; ch < p.n_chunks; ++ch) {
                    mbar_wait(FULL(stage), phase);
                    tc_fence_after();
                    const uint32_t ahi = sbase + TcxSmemCells::a_hi + stage * TC_A_STAGE;
                    const uint32_t alo = sbase + TcxSmemCells::a_lo + stage * TC_A_STAGE;
                    const uint32_t ws = sbase + TcxSmemCells::w + stage * TCX_W_STAGE;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t da_hi = desc_sw128(ahi + ks * 32, 16, 1024);
                        const uint64_t dw_hi = desc_sw128(ws + ks * 32, 16, 1024);
                        tc_mma_f16(dC, da_hi, dw_hi, ID, (ch | ks) != 0);
                        tc_mma_f16(dC, da_hi, desc_sw128(ws + TCX_KP * 128 + ks * 32, 16, 1024), ID, 1);
                        tc_mma_f16(dC, desc_sw128(alo + ks * 32, 16, 1024), dw_hi, ID, 1);
                    }
                    tc_commit(EMPTY(stage));
                    if (++stage == TCX_STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(CFULL(b));
            }
        }
    } else {
        const int quarter = warp & 3;
        const float descale = exp2f(-(float)(p.a_exp + *p.x_exp));
        for (int i = 0; i < n_tiles; ++i) {
            const int b = i & 1;
            mbar_wait(CFULL(b), (i >> 1) & 1);
            tc_fence_after();
            float c0[32], c1[32];
            const uint32_t trow = tmem + b * TCX_KP + ((uint32_t)(quarter * 32) << 16);
            tc_ld32(trow, c0);
            tc_ld32(trow + 32, c1);
            tc_fence_before();
            mbar_arrive(CEMPTY(b));
            const int cell = (t_begin + i) * TC_TILE + quarter * 32 + lane;
            if (cell < p.m) {
                float4 *dst = reinterpret_cast<float4 *>(p.C + (size_t)cell * p.ld);
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (4 * r < p.ld)
                        dst[r] = make_float4(c0[4 * r] * descale, c0[4 * r + 1] * descale, c0[4 * r + 2] * descale, c0[4 * r + 3] * descale);
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (32 + 4 * r < p.ld)
                        dst[8 + r] = make_float4(c1[4 * r] * descale, c1[4 * r + 1] * descale, c1[4 * r + 2] * descale, c1[4 * r + 3] * descale);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
    }
}

// ------------------------------------------------------------------------------------------------
// B += A' H.  warp 0 TMA producer (64 cells x 128 genes per stage), warp 1 MMA issuer, warp 2 H image loader,
// warps 4..7 drain.  Genes are covered in sub-passes of 3 x 128 (two TMEM regions of 3 x 64 columns: the drain of
// one sub-pass overlaps the MMAs of the next); each sub-pass streams those genes of all of this CTA's cells.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TCX_THREADS_GENES, 1)
tcx_genes_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo, const TcxParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t_begin, n_tiles;
    tcx_my_tiles(p.tiles_total, p.tile0, t_begin, n_tiles);
    const int n_pass = (p.n_mtiles + TCX_MT_PER_PASS - 1) / TCX_MT_PER_PASS;
    const int regions = p.regions;
    const uint32_t tmem_cols = regions == 2 ? 512u : 256u;

    const uint32_t bar0 = sbase + TcxSmemGenes::bars;
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (TCX_STAGES + s); };
    constexpr int B2 = 2 * TCX_STAGES;
    auto HFULL = [&](int b) { return bar0 + 8u * (B2 + b); };
    auto HEMPTY = [&](int b) { return bar0 + 8u * (B2 + 2 + b); };
    auto DFULL = [&](int r) { return bar0 + 8u * (B2 + 4 + r); };
    auto DEMPTY = [&](int r) { return bar0 + 8u * (B2 + 6 + r); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcxSmemGenes::misc);

    // cells of a partial tile that TMA never writes meet zero rows of the H image: they must be finite
    for (int i = threadIdx.x; i < (2 * TCX_STAGES * TC_A_STAGE) / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem + TcxSmemGenes::a_hi)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < TCX_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(HFULL(b), 1); mbar_init(HEMPTY(b), 1); }
        for (int r = 0; r < 2; ++r) { mbar_init(DFULL(r), 1); mbar_init(DEMPTY(r), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // unit u accumulates into region u % regions; its barriers' phase advances every `regions` units
    auto reg_of = [&](int u) { return regions == 2 ? (u & 1) : 0; };
    auto ph_of = [&](int u) { return regions == 2 ? ((u >> 1) & 1) : (u & 1); };
    auto tmem_d = [&](int u) { return tmem + reg_of(u) * (TCX_MT_PER_PASS * TCX_KP); };

    if (warp == 0) {
        int stage = 0, phase = 0;
        for (int u = 0; u < n_pass; ++u) {
            const int mt0 = u * TCX_MT_PER_PASS, mt1 = min(p.n_mtiles, mt0 + TCX_MT_PER_PASS);
            for (int i = 0; i < n_tiles; ++i) {
                const int cell0 = (t_begin + i) * TC_TILE;
                const int nbox = (min(TC_TILE, p.m - cell0) + TC_GROUP - 1) / TC_GROUP;
                for (int mt = mt0; mt < mt1; ++mt) {
                    for (int hc = 0; hc < 2; ++hc) {
                        const int nb = min(2, nbox - 2 * hc);
                        if (nb <= 0) break;
                        if (lane == 0) {
                            mbar_wait(EMPTY(stage), phase ^ 1);
                            mbar_expect_tx(FULL(stage), nb * 4 * TC_BOX_BYTES);
                        }
                        __syncwarp();
                        {
                            const int hl = lane & 1, gh = (lane >> 1) & 1, cg = lane >> 2;
                            if (cg < nb) {
                                const int off = stage * TC_A_STAGE + gh * 2 * TC_BOX_BYTES + cg * TC_BOX_BYTES;
                                tma_load_2d(sbase + (hl ? TcxSmemGenes::a_lo : TcxSmemGenes::a_hi) + off, hl ? &map_alo : &map_ahi,
                                            (2 * mt + gh) * TC_CHUNK, cell0 + (2 * hc + cg) * TC_GROUP, FULL(stage));
                            }
                        }
                        if (++stage == TCX_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t ID = idesc_f16(128, TCX_KP, 1, 0, 0);
            int stage = 0, phase = 0, use = 0;
            for (int u = 0; u < n_pass; ++u) {
                const int mt0 = u * TCX_MT_PER_PASS, mt1 = min(p.n_mtiles, mt0 + TCX_MT_PER_PASS);
                const int r = reg_of(u);
                mbar_wait(DEMPTY(r), ph_of(u) ^ 1);
                tc_fence_after();
                for (int i = 0; i < n_tiles; ++i, ++use) {
                    const int hb = use & 1;
                    const int nbox = (min(TC_TILE, p.m - (t_begin + i) * TC_TILE) + TC_GROUP - 1) / TC_GROUP;
                    mbar_wait(HFULL(hb), (use >> 1) & 1);
                    tc_fence_after();
                    const uint32_t hs = sbase + TcxSmemGenes::h + hb * TCX_H_TILE;
                    for (int mt = mt0; mt < mt1; ++mt) {
                        const uint32_t dD = tmem_d(u) + (mt - mt0) * TCX_KP;
                        for (int hc = 0; hc < 2; ++hc) {
                            if (nbox - 2 * hc <= 0) break;
                            mbar_wait(FULL(stage), phase);
                            tc_fence_after();
                            const uint32_t ahi = sbase + TcxSmemGenes::a_hi + stage * TC_A_STAGE;
                            const uint32_t alo = sbase + TcxSmemGenes::a_lo + stage * TC_A_STAGE;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                // A operand, MN-major: 64 genes contiguous (128 B), 8 cells per swizzle atom (SBO 1024),
                                // the second 64-gene atom 8 KB further (LBO)
                                const uint64_t da_hi = desc_sw128(ahi + ks * 2048, 2 * TC_BOX_BYTES, 1024);
                                const uint64_t da_lo = desc_sw128(alo + ks * 2048, 2 * TC_BOX_BYTES, 1024);
                                const uint32_t hoff = hc * TCX_H_SUB + ks * 32;
                                const uint64_t dh_hi = desc_sw128(hs + hoff, 16, 1024);
                                const uint64_t dh_lo = desc_sw128(hs + 2 * TCX_H_SUB + hoff, 16, 1024);
                                tc_mma_f16(dD, da_hi, dh_hi, ID, (i | hc | ks) != 0);
                                tc_mma_f16(dD, da_hi, dh_lo, ID, 1);
                                tc_mma_f16(dD, da_lo, dh_hi, ID, 1);
                            }
                            tc_commit(EMPTY(stage));
                            if (++stage == TCX_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                    tc_commit(HEMPTY(hb));
                }
                tc_commit(DFULL(r));
            }
        }
    } else if (warp == 2) {
        int use = 0;
        for (int u = 0; u < n_pass; ++u) {
            for (int i = 0; i < n_tiles; ++i, ++use) {
                const int hb = use & 1;
                if (lane == 0) {
                    mbar_wait(HEMPTY(hb), ((use >> 1) & 1) ^ 1);
                    mbar_expect_tx(HFULL(hb), TCX_H_TILE);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     sbase + TcxSmemGenes::h + hb * TCX_H_TILE),
                                 "l"(p.Himg + (size_t)(t_begin + i) * TCX_H_TILE), "r"(TCX_H_TILE), "r"(HFULL(hb))
                                 : "memory");
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3;
        const float descale = exp2f(-(float)(p.a_exp + *p.x_exp));
        for (int u = 0; u < n_pass; ++u) {
            const int mt0 = u * TCX_MT_PER_PASS, mt1 = min(p.n_mtiles, mt0 + TCX_MT_PER_PASS);
            mbar_wait(DFULL(reg_of(u)), ph_of(u));
            tc_fence_after();
            for (int mt = mt0; mt < mt1; ++mt) {
                const int gene = mt * 128 + quarter * 32 + lane;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float d[32];
                    tc_ld32(tmem_d(u) + (mt - mt0) * TCX_KP + half * 32 + ((uint32_t)(quarter * 32) << 16), d);
                    if (gene < p.n && n_tiles > 0) {
                        float *dst = p.B + (size_t)gene * p.ld + half * 32;
#pragma unroll
                        for (int r = 0; r < 8; ++r)
                            if (half * 32 + 4 * r < p.ld)
                                red_add_v4(dst + 4 * r, d[4 * r] * descale, d[4 * r + 1] * descale, d[4 * r + 2] * descale,
                                       d[4 * r + 3] * descale);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(DEMPTY(reg_of(u)));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
    }
}

// ------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int tcx_exp_of(int max_bits)
{
    const float mx = __int_as_float(max_bits);
    int e = 0, x = 0;
    if (mx > 0.f) {
        frexpf(mx, &e);
        x = 14 - e;
    }
    return max(-100, min(100, x));
}
// Wt fp32 [n][ld] -> Wst fp16 [128][n_pad] (rows 0..63 hi, 64..127 lo), scaled by 2^w_exp with max -> ~2^14
__global__ void tcx_split_w_kernel(const float *__restrict__ Wt, int ld, int n, int n_pad, const int *__restrict__ max_bits,
                                   int *__restrict__ w_exp_out, __half *__restrict__ Wst)
{
    const int w_exp = tcx_exp_of(*max_bits);
    if (blockIdx.x == 0 && threadIdx.x == 0) *w_exp_out = w_exp;
    const float scale = exp2f((float)w_exp);
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_pad) return;
#pragma unroll 4
    for (int f = 0; f < TCX_KP; ++f) {
        const float x = (g < n && f < ld) ? Wt[(size_t)g * ld + f] * scale : 0.f;
        const __half h = __float2half_rn(x);
        Wst[(size_t)f * n_pad + g] = h;
        Wst[(size_t)(TCX_KP + f) * n_pad + g] = __float2half_rn(x - __half2float(h));
    }
}
// H fp32 [m][ld] -> per 128-cell tile the smem-ready split image: four K-major [64 factors][64 cells] sub-tiles
// (hi cells 0..63, hi cells 64..127, lo, lo) with 128 B rows in the 128B-swizzle pattern.  One block per tile.
__global__ void __launch_bounds__(128) tcx_split_h_kernel(const float *__restrict__ H, int ld, int m, const int *__restrict__ max_bits,
                                                          int *__restrict__ h_exp_out, uint8_t *__restrict__ Himg, int tile0)
{
    const int h_exp = tcx_exp_of(*max_bits);
    if (blockIdx.x == 0 && threadIdx.x == 0) *h_exp_out = h_exp;
    const float scale = exp2f((float)h_exp);
    const int lc = threadIdx.x;
    const int tile = tile0 + blockIdx.x;
    const int cell = tile * TC_TILE + lc;
    uint8_t *img = Himg + (size_t)tile * TCX_H_TILE + (lc >> 6) * TCX_H_SUB;
    const int cc = lc & 63;
    const float4 *row = reinterpret_cast<const float4 *>(H + (size_t)min(cell, m - 1) * ld);
#pragma unroll 4
    for (int r = 0; r < TCX_KP / 4; ++r) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cell < m && 4 * r < ld) v = row[r];
        const float xs[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int f = 4 * r + q;
            const __half hi = __float2half_rn(xs[q]);
            const __half lo = __float2half_rn(xs[q] - __half2float(hi));
            const int off = f * 128 + ((((cc >> 3) ^ (f & 7)) << 4) | ((cc & 7) << 1));
            *reinterpret_cast<__half *>(img + off) = hi;
            *reinterpret_cast<__half *>(img + 2 * TCX_H_SUB + off) = lo;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core block Gauss-Seidel solve for up to 64 coordinates (the k <= 32 form is tc_block_gs in pass_tc.cuh).
// One CTA of 4 warps per 128 columns; mu' of the tile lives in TMEM columns [0,64), the delta operand in [64,96).
// c_gram must hold G' and G0 at row stride 64 (tcx_repack_gram_kernel); gsrc is the same data in global memory.
//   X [cols][ld] in/out, C [cols][ld] the contraction, dvec = D[ld] | invD[ld] of the GramPack.
// NB = number of 16-coordinate blocks in use (ceil(k / 16)): blocks past k are never touched.
// ------------------------------------------------------------------------------------------------
constexpr int TCX_GOP = 2048;   // one [64 x 8] tf32 operand of G' (per 8-row block, hi / lo)

__global__ void tcx_repack_gram_kernel(const float *__restrict__ pack, int ld, float *__restrict__ out)
{
    // out[0][64][64] = G' (unit diagonal scaled), out[1][64][64] = G0; zero outside ld x ld
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 2 * 64 * 64; t += gridDim.x * blockDim.x) {
        const int which = t >> 12, r = (t >> 6) & 63, c = t & 63;
        out[t] = (r < ld && c < ld) ? pack[which * ld * ld + r * ld + c] : 0.f;
    }
}

template <int NB>
__global__ void __launch_bounds__(128)
tcx_solve_kernel(float *__restrict__ X, const float *__restrict__ C, const float *__restrict__ gsrc,
                 const float *__restrict__ dvec, int ld, float l1, int k, int cols, int max_iter, float rel_tol,
                 unsigned long long *sweeps, double *loss_part)
{
    constexpr int KW = 16 * NB;
    __shared__ __align__(1024) uint8_t gops[4 * NB * TCX_GOP];
    __shared__ float sD[64], sInvD[64];
    __shared__ __align__(8) uint64_t mbar_store;
    __shared__ uint32_t tmem_slot;
    __shared__ double scratch[32];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), quarter = warp;     // warp-uniform for the compiler
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const uint32_t mbar = smem_u32(&mbar_store);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // tf32 operands of G': op[b8][hi|lo][n][kq] = split(G'[8 b8 + kq][n]), [64 x 8] K-major, no swizzle
    for (int t = threadIdx.x; t < 2 * NB * 64 * 8; t += blockDim.x) {
        const int b8 = t >> 9, n = (t >> 3) & 63, kq = t & 7;
        const float v = gsrc[(8 * b8 + kq) * 64 + n];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const int off = (n >> 3) * 256 + (kq >> 2) * 128 + (n & 7) * 16 + (kq & 3) * 4;
        *reinterpret_cast<float *>(gops + (2 * b8) * TCX_GOP + off) = hi;
        *reinterpret_cast<float *>(gops + (2 * b8 + 1) * TCX_GOP + off) = v - hi;
    }
    if (threadIdx.x < 64) {
        sD[threadIdx.x] = threadIdx.x < ld ? dvec[threadIdx.x] : 1.f;
        sInvD[threadIdx.x] = threadIdx.x < ld ? dvec[ld + threadIdx.x] : 1.f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tmu = tmem, tdelta = tmem + 64;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t trow = tmu + lane_off;
    const uint32_t gop_u = smem_u32(gops);
    constexpr uint32_t ID_TF32 = idesc_tf32(128, 64);
    const int j = blockIdx.x * 128 + threadIdx.x;
    const bool valid = j < cols;
    float *xrow = X + (size_t)(valid ? j : 0) * ld;
    const float *crow = C + (size_t)(valid ? j : 0) * ld;

    uint32_t mb_parity = 0;
    bool pending = false;
    // rank-16 update Mu += D G'[16 blk .. 16 blk + 15, :] of the deltas sitting in the operand columns
    auto issue_update = [&](int blk) {
        tc_fence_before();
        named_bar_sync(1, 128);
        if (warp == 0) {
            // elect.sync in warp-uniform control flow: the six UTCHMMA issue back to back (under `threadIdx.x == 0` each is
            // wrapped in an ELECT / BRA.U.ANY loop, ~80 cycles per instruction; tools/micro/gs_bench.cu)
            if (elect_one_sync()) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const int b8 = 2 * blk + ks;
                    const uint64_t g_hi = desc_nosw(gop_u + (2 * b8) * TCX_GOP, 128, 256);
                    const uint64_t g_lo = desc_nosw(gop_u + (2 * b8 + 1) * TCX_GOP, 128, 256);
                    tc_mma_tf32_ts(tmu, tdelta + 8 * ks, g_hi, ID_TF32, 1);
                    tc_mma_tf32_ts(tmu, tdelta + 8 * ks, g_lo, ID_TF32, 1);
                    tc_mma_tf32_ts(tmu, tdelta + 16 + 8 * ks, g_hi, ID_TF32, 1);
                }
                tc_commit(mbar);
            }
            __syncwarp();
        }
        pending = true;
    };
    auto wait_update = [&]() {
        if (pending) { mbar_wait(mbar, mb_parity); mb_parity ^= 1u; pending = false; }
        tc_fence_after();
    };
    auto put_deltas = [&](const float (&d)[16]) {
        float hl[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            hl[i] = __uint_as_float(__float_as_uint(d[i]) & 0xffffe000u);
            hl[16 + i] = d[i] - hl[i];
        }
        tc_st32(tdelta + lane_off, hl);
    };

    // ---- h' = D o x; mu' = (l1 - c) / D + G' h'  (the G' h' term through the same rank-16 updates, from h' = 0)
    float h[KW];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float mv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int f = 16 * b + i;
            const bool in = valid && f < ld;
            h[f] = in ? xrow[f] * sD[f] : 0.f;
            mv[i] = in ? (l1 - crow[f]) * sInvD[f] : 0.f;
        }
        {
            uint32_t r[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(mv[i]);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
                    trow + 16 * b),
                "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                : "memory");
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b * 16 < k) {
            wait_update();
            float d[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = h[16 * b + i];
            put_deltas(d);
            issue_update(b);
        }
    }

    // ---- sweeps
    const bool any_change = rel_tol < 2.9e-8f;
    int my_sweeps = 1;
    unsigned any = 1u;
    for (int sweep = 0; sweep < max_iter && any; ++sweep) {
        unsigned changed = 0u;
#pragma unroll
        for (int blk = 0; blk < NB; ++blk) {
            if (blk * 16 < k) {
                wait_update();
                float m16[16], d[16];
                tc_ld16(trow + 16 * blk, m16);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int kk = 16 * blk + i;
                    const float tmp = fmaxf(h[kk] - m16[i], 0.f);
                    d[i] = tmp - h[kk];
                    if (any_change) changed |= __float_as_uint(d[i]);
                    else changed |= (2.f * fabsf(d[i]) > rel_tol * (tmp + h[kk] + TINY_NUM_F)) ? 1u : 0u;
                    h[kk] = tmp;
#pragma unroll
                    for (int jj = i + 1; jj < 16; ++jj) m16[jj] = fmaf(d[i], c_gram[kk * 64 + 16 * blk + jj], m16[jj]);
                }
                put_deltas(d);
                issue_update(blk);
            }
        }
        any = named_bar_or(1, 128, changed);
        if (changed) my_sweeps = min(sweep + 2, max_iter);
    }
    wait_update();

    double lp = 0.0;
    if (valid) {
#pragma unroll
        for (int f = 0; f < KW; ++f) {
            h[f] *= sInvD[f];
            if (f < ld) xrow[f] = h[f];
        }
        if (loss_part) {
            // h' G0 h - 2 h' c
            float s = 0.f;
#pragma unroll 1
            for (int q = 0; q < k; ++q) {
                float t = 0.f;
                const float *g0 = c_gram + 64 * 64 + q * 64;
#pragma unroll
                for (int f = 0; f < KW; ++f) t = fmaf(g0[f], h[f], t);
                s = fmaf(xrow[q], t - 2.f * crow[q], s);
            }
            lp = (double)s;
        }
    }
    const double tot_it = block_sum(valid ? (double)my_sweeps : 0.0, scratch);
    if (threadIdx.x == 0) atomicAdd(sweeps, (unsigned long long)(tot_it + 0.5));
    if (loss_part) {
        const double tot = block_sum(lp, scratch);
        if (threadIdx.x == 0) atomicAdd(loss_part, tot);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
    }
}

struct TcxState {
    int n_pad = 0, tiles_total = 0;
    __half *Wst = nullptr;       // [128][n_pad]
    uint8_t *Himg = nullptr;     // [tiles_total][TCX_H_TILE]
    int *ints = nullptr;         // [0] W max bits, [1] w_exp, [2 + 2 s] H max bits, [3 + 2 s] h_exp of slab slot s (0..3)
    float *gram64 = nullptr;     // G' and G0 repacked to row stride 64 (source of c_gram and of the tf32 operands)
    CUtensorMap map_w;
    bool attr = false;
};

static inline void tcx_release(TcxState &x)
{
    if (x.Wst) cudaFree(x.Wst);
    if (x.Himg) cudaFree(x.Himg);
    if (x.ints) cudaFree(x.ints);
    if (x.gram64) cudaFree(x.gram64);
    x = TcxState();
}

// shapes the un-fused kernels handle: any k up to 64 (the row stride of C, B, W, H is k rounded up to 8)
static inline bool tcx_supported(int n, int m, int k)
{
    return k >= 1 && k <= TCX_KP && n >= 128 && m >= 128;
}

// tc must be prepared (split copy of A + its tensor maps)
static inline void tcx_prepare(TcxState &x, const TcMatrix &tc)
{
    const int tiles_total = (tc.m + TC_TILE - 1) / TC_TILE;
    if (x.Wst && x.n_pad == tc.n_pad && x.tiles_total == tiles_total) return;
    tcx_release(x);
    x.n_pad = tc.n_pad; x.tiles_total = tiles_total;
    CUDA_CHECK(cudaMalloc(&x.Wst, (size_t)2 * TCX_KP * tc.n_pad * sizeof(__half)));
    CUDA_CHECK(cudaMalloc(&x.Himg, (size_t)tiles_total * TCX_H_TILE));
    CUDA_CHECK(cudaMalloc(&x.ints, 16 * sizeof(int)));
    CUDA_CHECK(cudaMalloc(&x.gram64, 2 * 64 * 64 * sizeof(float)));
    tc_make_map(&x.map_w, x.Wst, (uint64_t)tc.n_pad, 2 * TCX_KP, (uint64_t)tc.n_pad * sizeof(__half), TC_CHUNK, 2 * TCX_KP);
    CUDA_CHECK(cudaFuncSetAttribute(tcx_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcxSmemCells::total + 1024));
    CUDA_CHECK(cudaFuncSetAttribute(tcx_genes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcxSmemGenes::total + 1024));
    // The slab pipeline wants solve CTAs (25 KB of shared memory each) resident NEXT TO a contraction CTA.  An SM's shared
    // memory carve-out is fixed while CTAs are resident, and by default it is the smallest one that fits the first kernel
    // (164 KB for 145 KB): ask for the largest carve-out so that the co-resident CTAs fit.
    CUDA_CHECK(cudaFuncSetAttribute(tcx_cells_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_CHECK(cudaFuncSetAttribute(tcx_genes_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_CHECK(cudaFuncSetAttribute(tcx_solve_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_CHECK(cudaFuncSetAttribute(tcx_solve_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
}

static inline TcxParams tcx_params(const TcxState &x, const TcMatrix &tc, int ld, int tile0, int ntiles)
{
    TcxParams p;
    p.n = tc.n; p.m = tc.m; p.ld = ld;
    p.n_chunks = tc.n_pad / TC_CHUNK;
    p.n_mtiles = (tc.n + 127) / 128;
    p.tiles_total = ntiles; p.tile0 = tile0;
    p.C = nullptr; p.B = nullptr; p.Himg = x.Himg; p.x_exp = nullptr;
    p.a_exp = tc.a_exp;
    p.regions = 2;
    return p;
}

// W changed: its split image (max -> exponent -> stacked fp16 operand).  launches: 2
static inline void tcx_split_w(TcxState &x, const TcMatrix &tc, int ld, const float *Wt, cudaStream_t st)
{
    CUDA_CHECK(cudaMemsetAsync(x.ints, 0, sizeof(int), st));
    tc_absmax_kernel<<<32, 256, 0, st>>>(Wt, (size_t)tc.n * ld, x.ints);
    tcx_split_w_kernel<<<(tc.n_pad + 127) / 128, 128, 0, st>>>(Wt, ld, tc.n, tc.n_pad, x.ints, x.ints + 1, x.Wst);
    CUDA_CHECK(cudaGetLastError());
}
// C[cells of the slab][ld] = A Wt for the tiles [tile0, tile0 + ntiles); the split image of W must be current.  launches: 1
static inline void tcx_cells_slab(TcxState &x, const TcMatrix &tc, int ld, float *C, int tile0, int ntiles, cudaStream_t st)
{
    TcxParams p = tcx_params(x, tc, ld, tile0, ntiles);
    p.C = C; p.x_exp = x.ints + 1;
    const int grid = std::min(tc.sm_count, ntiles);
    tcx_cells_kernel<<<grid, TCX_THREADS_CELLS, TcxSmemCells::total + 1024, st>>>(tc.map_ahi, tc.map_alo, x.map_w, p);
    CUDA_CHECK(cudaGetLastError());
}
// B[n][ld] += A' H over the cells of the slab (B is NOT zeroed here).  slot: which pair of x.ints holds this slab's max / exponent
// (slabs in flight on different streams use different slots: 0..3).  launches: 3
static inline void tcx_genes_slab(TcxState &x, const TcMatrix &tc, int ld, const float *H, float *B, int tile0, int ntiles, int slot,
                                  cudaStream_t st, int regions = 2)
{
    int *mx = x.ints + 2 + 2 * slot;
    const int cell0 = tile0 * TC_TILE, cells = std::min(tc.m - cell0, ntiles * TC_TILE);
    CUDA_CHECK(cudaMemsetAsync(mx, 0, sizeof(int), st));
    tc_absmax_kernel<<<148, 256, 0, st>>>(H + (size_t)cell0 * ld, (size_t)cells * ld, mx);
    tcx_split_h_kernel<<<ntiles, 128, 0, st>>>(H, ld, tc.m, mx, mx + 1, x.Himg, tile0);
    TcxParams p = tcx_params(x, tc, ld, tile0, ntiles);
    p.B = B; p.x_exp = mx + 1; p.regions = regions;
    const int grid = std::min(tc.sm_count, ntiles);
    tcx_genes_kernel<<<grid, TCX_THREADS_GENES, TcxSmemGenes::total + 1024, st>>>(tc.map_ahi, tc.map_alo, p);
    CUDA_CHECK(cudaGetLastError());
}

// C[m][ld] = A Wt.  launches: 3
static inline void tcx_contract_cells(TcxState &x, const TcMatrix &tc, int ld, const float *Wt, float *C, cudaStream_t st)
{
    tcx_split_w(x, tc, ld, Wt, st);
    tcx_cells_slab(x, tc, ld, C, 0, x.tiles_total, st);
}
// B[n][ld] = A' H (B zeroed here).  launches: 4
static inline void tcx_contract_genes(TcxState &x, const TcMatrix &tc, int ld, const float *H, float *B, cudaStream_t st)
{
    CUDA_CHECK(cudaMemsetAsync(B, 0, sizeof(float) * (size_t)tc.n * ld, st));
    tcx_genes_slab(x, tc, ld, H, B, 0, x.tiles_total, 0, st);
}

// per-column NNLS solves (H side: cols = cells, loss partial wanted; W side: cols = genes).  pack: GramPack at row stride ld.
// tcx_solve_prepare puts the Grams where the kernel reads them (once per half step); tcx_solve_run may then be launched on
// several column ranges (streams) of the same half step.
static inline void tcx_solve_prepare(TcxState &x, const float *pack, int ld, cudaStream_t st)
{
    tcx_repack_gram_kernel<<<8, 1024, 0, st>>>(pack, ld, x.gram64);
    CUDA_CHECK(cudaMemcpyToSymbolAsync(c_gram, x.gram64, sizeof(float) * 2 * 64 * 64, 0, cudaMemcpyDeviceToDevice, st));
}
static inline void tcx_solve_run(TcxState &x, float *X, const float *C, const float *pack, int ld, float l1, int k, int cols, int max_iter,
                                 float rel_tol, unsigned long long *sweeps, double *loss_part, cudaStream_t st)
{
    const float *dvec = pack + 2 * ld * ld;
    const int grid = (cols + 127) / 128;
    if (k <= 48)
        tcx_solve_kernel<3><<<grid, 128, 0, st>>>(X, C, x.gram64, dvec, ld, l1, k, cols, max_iter, rel_tol, sweeps, loss_part);
    else
        tcx_solve_kernel<4><<<grid, 128, 0, st>>>(X, C, x.gram64, dvec, ld, l1, k, cols, max_iter, rel_tol, sweeps, loss_part);
    CUDA_CHECK(cudaGetLastError());
}
static inline void tcx_solve(TcxState &x, float *X, const float *C, const float *pack, int ld, float l1, int k, int cols,
                             int max_iter, float rel_tol, unsigned long long *sweeps, double *loss_part, cudaStream_t st)
{
    tcx_solve_prepare(x, pack, ld, st);
    tcx_solve_run(x, X, C, pack, ld, l1, k, cols, max_iter, rel_tol, sweeps, loss_part, st);
}

}  // namespace swne
