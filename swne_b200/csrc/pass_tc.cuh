// pass_tc.cuh -- tcgen05 engine (SWNE_ENGINE_TC) of the mse path: one persistent, warp-specialised kernel
// per outer iteration that streams the cells x genes matrix through TMA-staged shared-memory tiles twice:
//
//   phase 1  C = A Wt          (cells x k)   M = 128 cells, K = genes; per tile the accumulator (TMEM) goes
//            straight to the solver warps: thread-per-cell SCD NNLS (scd_ls_update), H written back;
//   phase 3  B += A' H         (genes x k)   M = 128 genes (MN-major view of the SAME kind of tile), K = cells;
//            accumulators for half of the genes live in TMEM (12 x 32 columns), so phase 3 runs once per gene
//            half, each time reading only that half of every cell row; drained with vector red.add to global B.
//
// Operands are "split fp16": x = hi + lo with hi = fp16(x), lo = fp16(x - hi), after a power-of-two scaling,
// and every product is hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM (the fp32-emulating scheme the
// north star allows; error ~2^-21).  A is stored pre-split (2 + 2 bytes per entry = the 4 bytes of fp32), so
// the algorithmic HBM traffic is unchanged and no SIMT split pass touches the streamed tile.
//
// Why two reads of A per iteration instead of one: the per-cell NNLS solve has a latency of 10^4..10^5
// cycles; keeping a cell's 12 KB row on chip (smem or L2) for that long for enough cells in flight to keep the
// CUDA cores busy needs > 200 MB chip-wide.  See DESIGN.md "pass structure".
//
// Reference semantics restated: update() / scd_ls_update of NNLM, call sites
// /root/reference/R/nmf.R:117,129.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels_simt.cuh"

namespace swne {

constexpr int TC_KP = 32;             // padded factor count handled by this engine (k <= 32)
constexpr int TC_TILE = 128;          // cells per tile (MMA M in phase 1, K in phase 3)
constexpr int TC_CHUNK = 64;          // genes per stage (128 B of fp16: one swizzle-128B row)
constexpr int TC_STAGES = 4;
constexpr int TC_GROUP = 32;          // cells per TMA box / work granule
constexpr int TC_A_STAGE = TC_TILE * TC_CHUNK * 2;      // 16 KB
constexpr int TC_W_STAGE = 64 * TC_CHUNK * 2;           // 8 KB  (rows 0..31 W_hi, 32..63 W_lo)
constexpr int TC_BOX_BYTES = TC_GROUP * TC_CHUNK * 2;   // 4 KB
constexpr int TC_H_SUB = TC_KP * 64 * 2;                // 4 KB: [32 factors][64 cells] fp16, K-major
constexpr int TC_H_TILE = 4 * TC_H_SUB;                 // hi k-half 0/1, lo k-half 0/1
constexpr int TC_MT_PER_PASS = 12;                      // 128-gene accumulators per gene sub-pass (12 * 32 = 384 TMEM columns)
constexpr int TC_THREADS = 64 + 256;

struct TcSmem {
    // offsets from the 1024-aligned base
    static constexpr int a_hi = 0;
    static constexpr int a_lo = a_hi + TC_STAGES * TC_A_STAGE;
    static constexpr int w = a_lo + TC_STAGES * TC_A_STAGE;
    static constexpr int h = w + TC_STAGES * TC_W_STAGE;
    static constexpr int g = h + 2 * TC_H_TILE;
    static constexpr int g0 = g + TC_KP * TC_KP * 4;
    static constexpr int invd = g0 + TC_KP * TC_KP * 4;
    static constexpr int bars = invd + TC_KP * 4;
    static constexpr int n_bars = 2 * TC_STAGES + 2 + 2 + 2 + 2 + 2;
    static constexpr int misc = bars + n_bars * 8;
    static constexpr int total = misc + 64;
};

struct TcPassParams {
    int n, m, k;
    int n_chunks;            // gene chunks per tile (even, >= ceil(n/64))
    int n_mtiles;            // ceil(n/128)
    int mode;                // 0: phase 1 + solve + phase 3; 1: phase 3 only (A H' of the given H)
    const float *G, *G0;     // penalised / plain Gram of W (KP x KP)
    float l1;
    int max_iter;
    float rel_tol;
    float *H;                // [m][32] fp32 in/out
    float *B;                // [n][32] fp32, accumulated with red.add (zeroed by the caller)
    FitScalars *scal;
    const int *w_exp;        // device: W was scaled by 2^w_exp before the split
    int a_exp;               // A was scaled by 2^a_exp before the split
    int groups_total;        // ceil(m / 32)
};

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): f16 x f16 -> f32
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major)
{
    return (1u << 4) /* C = F32 */ | (0u << 7) /* A = F16 */ | (0u << 10) /* B = F16 */ | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------
// the pass kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_pass_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
               const __grid_constant__ CUtensorMap map_w, const TcPassParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- my cells: a contiguous range of 32-cell groups
    const int gpc = p.groups_total / gridDim.x, grem = p.groups_total % gridDim.x;
    const int g_begin = blockIdx.x * gpc + min((int)blockIdx.x, grem);
    const int g_count = gpc + ((int)blockIdx.x < grem ? 1 : 0);
    const int cell_begin = g_begin * TC_GROUP;
    const int cell_end = min(p.m, (g_begin + g_count) * TC_GROUP);
    const int n_tiles = (cell_end - cell_begin + TC_TILE - 1) / TC_TILE;
    const int n_pass = (p.n_mtiles + TC_MT_PER_PASS - 1) / TC_MT_PER_PASS;

    // ---- barriers
    const uint32_t bar0 = sbase + TcSmem::bars;
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (TC_STAGES + s); };
    auto CFULL = [&](int b) { return bar0 + 8u * (2 * TC_STAGES + b); };
    auto CEMPTY = [&](int b) { return bar0 + 8u * (2 * TC_STAGES + 2 + b); };
    auto HFULL = [&](int b) { return bar0 + 8u * (2 * TC_STAGES + 4 + b); };
    auto HEMPTY = [&](int b) { return bar0 + 8u * (2 * TC_STAGES + 6 + b); };
    const uint32_t DFULL = bar0 + 8u * (2 * TC_STAGES + 8), DEMPTY = bar0 + 8u * (2 * TC_STAGES + 9);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcSmem::misc);
    int *s_hmax = reinterpret_cast<int *>(smem + TcSmem::misc + 8);

    // zero the A stages once: rows of a partial tile that TMA never writes must hold finite values
    for (int i = threadIdx.x; i < (2 * TC_STAGES * TC_A_STAGE) / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem + TcSmem::a_hi)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(CFULL(b), 1); mbar_init(CEMPTY(b), 128);
            mbar_init(HFULL(b), 256); mbar_init(HEMPTY(b), 1);
        }
        mbar_init(DFULL, 1); mbar_init(DEMPTY, 256);
        *s_hmax = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (warp >= 2) {
        float *sG = reinterpret_cast<float *>(smem + TcSmem::g), *sG0 = reinterpret_cast<float *>(smem + TcSmem::g0);
        float *sInvD = reinterpret_cast<float *>(smem + TcSmem::invd);
        for (int t = threadIdx.x - 64; t < TC_KP * TC_KP; t += 256) { sG[t] = p.G[t]; sG0[t] = p.G0[t]; }
        for (int t = threadIdx.x - 64; t < TC_KP; t += 256) sInvD[t] = (t < p.k) ? 1.f / p.G[t * TC_KP + t] : 0.f;
    }
    // generic-proxy zero fill must be visible to the async proxy (TMA / UMMA reads)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_c = tmem;             // columns [0,128): two C accumulators of 64 columns
    const uint32_t tmem_d = tmem + 128;       // columns [128,512): twelve 128 x 32 accumulators

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            int stage = 0, phase = 0;
            if (p.mode == 0) {
                for (int t = 0; t < n_tiles; ++t) {
                    const int cell0 = cell_begin + t * TC_TILE;
                    const int nbox = (min(TC_TILE, cell_end - cell0) + TC_GROUP - 1) / TC_GROUP;
                    for (int ch = 0; ch < p.n_chunks; ++ch) {
                        mbar_wait(EMPTY(stage), phase ^ 1);
                        mbar_expect_tx(FULL(stage), nbox * 2 * TC_BOX_BYTES + TC_W_STAGE);
                        for (int b = 0; b < nbox; ++b) {
                            tma_load_2d(sbase + TcSmem::a_hi + stage * TC_A_STAGE + b * TC_BOX_BYTES, &map_ahi, ch * TC_CHUNK,
                                        cell0 + b * TC_GROUP, FULL(stage));
                            tma_load_2d(sbase + TcSmem::a_lo + stage * TC_A_STAGE + b * TC_BOX_BYTES, &map_alo, ch * TC_CHUNK,
                                        cell0 + b * TC_GROUP, FULL(stage));
                        }
                        tma_load_2d(sbase + TcSmem::w + stage * TC_W_STAGE, &map_w, ch * TC_CHUNK, 0, FULL(stage));
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            for (int ps = 0; ps < n_pass; ++ps) {
                const int mt0 = ps * TC_MT_PER_PASS, mt1 = min(p.n_mtiles, mt0 + TC_MT_PER_PASS);
                for (int t = 0; t < n_tiles; ++t) {
                    const int cell0 = cell_begin + t * TC_TILE;
                    const int nbox = (min(TC_TILE, cell_end - cell0) + TC_GROUP - 1) / TC_GROUP;
                    for (int ch = 2 * mt0; ch < 2 * mt1; ++ch) {
                        mbar_wait(EMPTY(stage), phase ^ 1);
                        mbar_expect_tx(FULL(stage), nbox * 2 * TC_BOX_BYTES);
                        for (int b = 0; b < nbox; ++b) {
                            tma_load_2d(sbase + TcSmem::a_hi + stage * TC_A_STAGE + b * TC_BOX_BYTES, &map_ahi, ch * TC_CHUNK,
                                        cell0 + b * TC_GROUP, FULL(stage));
                            tma_load_2d(sbase + TcSmem::a_lo + stage * TC_A_STAGE + b * TC_BOX_BYTES, &map_alo, ch * TC_CHUNK,
                                        cell0 + b * TC_GROUP, FULL(stage));
                        }
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            constexpr uint32_t ID_P1_N64 = idesc_f16(128, 64, 0, 0);
            constexpr uint32_t ID_P1_N32 = idesc_f16(128, 32, 0, 0);
            constexpr uint32_t ID_P3 = idesc_f16(128, 32, 1, 0);
            int stage = 0, phase = 0;
            if (p.mode == 0) {
                for (int t = 0; t < n_tiles; ++t) {
                    const int b = t & 1;
                    mbar_wait(CEMPTY(b), ((t >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t dC = tmem_c + b * 64;
                    for (int ch = 0; ch < p.n_chunks; ++ch) {
                        mbar_wait(FULL(stage), phase);
                        tc_fence_after();
                        const uint32_t ahi = sbase + TcSmem::a_hi + stage * TC_A_STAGE;
                        const uint32_t alo = sbase + TcSmem::a_lo + stage * TC_A_STAGE;
                        const uint32_t ws = sbase + TcSmem::w + stage * TC_W_STAGE;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t dw = desc_sw128(ws + ks * 32, 16, 1024);
                            tc_mma_f16(dC, desc_sw128(ahi + ks * 32, 16, 1024), dw, ID_P1_N64, (ch | ks) != 0);
                            tc_mma_f16(dC, desc_sw128(alo + ks * 32, 16, 1024), dw, ID_P1_N32, 1);
                        }
                        tc_commit(EMPTY(stage));
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(CFULL(b));
                }
            }
            int use = 0;  // H tile buffer uses
            for (int ps = 0; ps < n_pass; ++ps) {
                const int mt0 = ps * TC_MT_PER_PASS, mt1 = min(p.n_mtiles, mt0 + TC_MT_PER_PASS);
                mbar_wait(DEMPTY, (ps & 1) ^ 1);
                tc_fence_after();
                for (int t = 0; t < n_tiles; ++t, ++use) {
                    const int hb = use & 1;
                    mbar_wait(HFULL(hb), (use >> 1) & 1);
                    tc_fence_after();
                    const uint32_t hs = sbase + TcSmem::h + hb * TC_H_TILE;
                    for (int mt = mt0; mt < mt1; ++mt) {
                        mbar_wait(FULL(stage), phase);
                        mbar_wait(FULL(stage + 1), phase);
                        tc_fence_after();
                        const uint32_t ahi = sbase + TcSmem::a_hi + stage * TC_A_STAGE;
                        const uint32_t alo = sbase + TcSmem::a_lo + stage * TC_A_STAGE;
                        const uint32_t dD = tmem_d + (mt - mt0) * 32;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            // A operand, MN-major: 64 genes contiguous (128 B), 8 cells per swizzle atom (SBO 1024),
                            // second 64-gene atom in the next stage (LBO = stage stride)
                            const uint64_t da_hi = desc_sw128(ahi + ks * 2048, TC_A_STAGE, 1024);
                            const uint64_t da_lo = desc_sw128(alo + ks * 2048, TC_A_STAGE, 1024);
                            const uint32_t hoff = (ks >> 2) * TC_H_SUB + (ks & 3) * 32;
                            const uint64_t dh_hi = desc_sw128(hs + hoff, 16, 1024);
                            const uint64_t dh_lo = desc_sw128(hs + 2 * TC_H_SUB + hoff, 16, 1024);
                            tc_mma_f16(dD, da_hi, dh_hi, ID_P3, (t | ks) != 0);
                            tc_mma_f16(dD, da_hi, dh_lo, ID_P3, 1);
                            tc_mma_f16(dD, da_lo, dh_hi, ID_P3, 1);
                        }
                        tc_commit(EMPTY(stage));
                        tc_commit(EMPTY(stage + 1));
                        stage += 2;
                        if (stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(HEMPTY(hb));
                }
                tc_commit(DFULL);
            }
        }
    } else {
        // =========================== solver / loader / drain warps ===========================
        const int sw = warp - 2;                 // 0..7
        const int grp = sw >> 2;                 // solver group: handles tiles t with (t & 1) == grp
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const float *sG = reinterpret_cast<const float *>(smem + TcSmem::g);
        const float *sG0 = reinterpret_cast<const float *>(smem + TcSmem::g0);
        const float *sInvD = reinterpret_cast<const float *>(smem + TcSmem::invd);
        float hmax = 0.f;
        if (p.mode == 0) {
            const float descale = exp2f(-(float)(p.a_exp + *p.w_exp));
            double lp = 0.0;
            unsigned long long its = 0;
            for (int t = grp; t < n_tiles; t += 2) {
                const int b = t & 1;
                mbar_wait(CFULL(b), (t >> 1) & 1);
                tc_fence_after();
                float c[32], mu[32];
                const uint32_t taddr = tmem_c + b * 64 + ((uint32_t)(quarter * 32) << 16);
                tc_ld32(taddr, c);
                tc_ld32(taddr + 32, mu);
                tc_fence_before();
                mbar_arrive(CEMPTY(b));
                const int cell = cell_begin + t * TC_TILE + quarter * 32 + lane;
                if (cell < cell_end) {
                    float h[32];
                    float *hrow = p.H + (size_t)cell * TC_KP;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const float4 v = reinterpret_cast<const float4 *>(hrow)[r];
                        h[4 * r] = v.x; h[4 * r + 1] = v.y; h[4 * r + 2] = v.z; h[4 * r + 3] = v.w;
                    }
#pragma unroll
                    for (int f = 0; f < 32; ++f) {
                        c[f] = (c[f] + mu[f]) * descale;
                        mu[f] = p.l1 - c[f];
                    }
#pragma unroll 1
                    for (int q = 0; q < p.k; ++q) {
                        const float hq = hrow[q];
                        const float4 *g4 = reinterpret_cast<const float4 *>(sG + q * TC_KP);
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const float4 g = g4[r];
                            mu[4 * r + 0] = fmaf(hq, g.x, mu[4 * r + 0]);
                            mu[4 * r + 1] = fmaf(hq, g.y, mu[4 * r + 1]);
                            mu[4 * r + 2] = fmaf(hq, g.z, mu[4 * r + 2]);
                            mu[4 * r + 3] = fmaf(hq, g.w, mu[4 * r + 3]);
                        }
                    }
                    its += scd_ls_thread<TC_KP>(h, mu, sG, sInvD, p.k, p.max_iter, p.rel_tol);
#pragma unroll
                    for (int r = 0; r < 8; ++r)
                        reinterpret_cast<float4 *>(hrow)[r] = make_float4(h[4 * r], h[4 * r + 1], h[4 * r + 2], h[4 * r + 3]);
                    // loss partial h' G0 h - 2 h' c  (mu reused for G0 h)
#pragma unroll
                    for (int r = 0; r < 32; ++r) mu[r] = 0.f;
#pragma unroll 1
                    for (int q = 0; q < p.k; ++q) {
                        const float hq = hrow[q];
                        const float4 *g4 = reinterpret_cast<const float4 *>(sG0 + q * TC_KP);
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const float4 g = g4[r];
                            mu[4 * r + 0] = fmaf(hq, g.x, mu[4 * r + 0]);
                            mu[4 * r + 1] = fmaf(hq, g.y, mu[4 * r + 1]);
                            mu[4 * r + 2] = fmaf(hq, g.z, mu[4 * r + 2]);
                            mu[4 * r + 3] = fmaf(hq, g.w, mu[4 * r + 3]);
                        }
                    }
                    float s = 0.f;
#pragma unroll
                    for (int f = 0; f < 32; ++f) {
                        s = fmaf(h[f], mu[f] - 2.f * c[f], s);
                        hmax = fmaxf(hmax, h[f]);
                    }
                    lp += (double)s;
                }
            }
            lp = warp_sum(lp);
            double itd = warp_sum((double)its);
            if (lane == 0) {
                atomicAdd(&p.scal->loss_part, lp);
                atomicAdd(&p.scal->sweeps_h, (unsigned long long)(itd + 0.5));
            }
        } else {
            // phase 3 only: the scale comes from a max scan over my cells' H rows
            for (int i = threadIdx.x - 64; i < (cell_end - cell_begin) * TC_KP; i += 256)
                hmax = fmaxf(hmax, p.H[(size_t)cell_begin * TC_KP + i]);
        }
        // ---- per-CTA H scale: a power of two that brings the largest entry of my cells to ~2^14
        atomicMax(s_hmax, __float_as_int(hmax));
        __threadfence_block();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float hm = __int_as_float(*s_hmax);
        int h_exp = 0;
        if (hm > 0.f) {
            int e;
            frexpf(hm, &e);  // hm = f * 2^e, f in [0.5, 1)
            h_exp = 14 - e;
        }
        h_exp = max(-100, min(100, h_exp));
        const float hscale = exp2f((float)h_exp);
        const float bdescale = exp2f(-(float)(p.a_exp + h_exp));

        // ---- phase 3: H tile loader (all 256 threads) and accumulator drain
        const int ltid = threadIdx.x - 64;
        const int lcell = ltid & 127, fhalf = ltid >> 7;   // this thread converts factors fhalf*16 .. +16 of one cell
        int use = 0;
        for (int ps = 0; ps < n_pass; ++ps) {
            const int mt0 = ps * TC_MT_PER_PASS, mt1 = min(p.n_mtiles, mt0 + TC_MT_PER_PASS);
            for (int t = 0; t < n_tiles; ++t, ++use) {
                const int hb = use & 1;
                mbar_wait(HEMPTY(hb), ((use >> 1) & 1) ^ 1);
                const int cell = cell_begin + t * TC_TILE + lcell;
                float v[16];
                if (cell < cell_end) {
                    const float4 *src = reinterpret_cast<const float4 *>(p.H + (size_t)cell * TC_KP + fhalf * 16);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float4 x = src[r];
                        v[4 * r] = x.x; v[4 * r + 1] = x.y; v[4 * r + 2] = x.z; v[4 * r + 3] = x.w;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 16; ++r) v[r] = 0.f;
                }
                // K-major [factor][cell] fp16 sub-tiles of 64 cells, 128 B rows, swizzle-128B applied by hand
                uint8_t *tile = smem + TcSmem::h + hb * TC_H_TILE + (lcell >> 6) * TC_H_SUB;
                const int cc = lcell & 63;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int f = fhalf * 16 + r;
                    const float x = v[r] * hscale;
                    const __half hi = __float2half_rn(x);
                    const __half lo = __float2half_rn(x - __half2float(hi));
                    const int off = f * 128 + ((((cc >> 3) ^ (f & 7)) << 4) | ((cc & 7) << 1));
                    *reinterpret_cast<__half *>(tile + off) = hi;
                    *reinterpret_cast<__half *>(tile + 2 * TC_H_SUB + off) = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(HFULL(hb));
            }
            // ---- drain this gene sub-pass: D (128 genes x 32) -> red.add into B
            mbar_wait(DFULL, ps & 1);
            tc_fence_after();
            for (int mt = mt0 + grp; mt < mt1; mt += 2) {
                float d[32];
                tc_ld32(tmem_d + (mt - mt0) * 32 + ((uint32_t)(quarter * 32) << 16), d);
                const int gene = mt * 128 + quarter * 32 + lane;
                if (gene < p.n) {
                    float *dst = p.B + (size_t)gene * TC_KP;
#pragma unroll
                    for (int r = 0; r < 8; ++r)
                        red_add_v4(dst + 4 * r, d[4 * r] * bdescale, d[4 * r + 1] * bdescale, d[4 * r + 2] * bdescale,
                                   d[4 * r + 3] * bdescale);
                }
            }
            tc_fence_before();
            mbar_arrive(DEMPTY);
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------
// A (fp32 [m][ldA]) -> scaled split fp16 pair [m][n_pad]
__global__ void tc_split_a_kernel(const float *__restrict__ A, size_t ldA, int n, int m, int n_pad, float scale,
                                  __half *__restrict__ hi, __half *__restrict__ lo)
{
    for (size_t j = blockIdx.x; j < (size_t)m; j += gridDim.x)
        for (int g = threadIdx.x; g < n_pad; g += blockDim.x) {
            const float x = (g < n) ? A[j * ldA + g] * scale : 0.f;
            const __half h = __float2half_rn(x);
            hi[j * n_pad + g] = h;
            lo[j * n_pad + g] = __float2half_rn(x - __half2float(h));
        }
}
__global__ void tc_absmax_kernel(const float *__restrict__ X, size_t count, int *out_bits)
{
    float mx = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        mx = fmaxf(mx, fabsf(X[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_int(mx));
}
// Wt fp32 [n][32] -> Wst fp16 [64][n_pad] (rows 0..31 hi, 32..63 lo), scaled by 2^w_exp with max -> ~2^14
__global__ void tc_split_w_kernel(const float *__restrict__ Wt, int n, int n_pad, const int *__restrict__ max_bits,
                                  int *__restrict__ w_exp_out, __half *__restrict__ Wst)
{
    const float mx = __int_as_float(*max_bits);
    int w_exp = 0;
    if (mx > 0.f) {
        int e;
        frexpf(mx, &e);
        w_exp = 14 - e;
    }
    w_exp = max(-100, min(100, w_exp));
    if (blockIdx.x == 0 && threadIdx.x == 0) *w_exp_out = w_exp;
    const float scale = exp2f((float)w_exp);
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_pad) return;
#pragma unroll 4
    for (int f = 0; f < TC_KP; ++f) {
        const float x = (g < n) ? Wt[(size_t)g * TC_KP + f] * scale : 0.f;
        const __half h = __float2half_rn(x);
        Wst[(size_t)f * n_pad + g] = h;
        Wst[(size_t)(TC_KP + f) * n_pad + g] = __float2half_rn(x - __half2float(h));
    }
}

struct TcMatrix {
    bool ready = false;
    int n = 0, m = 0, n_pad = 0, a_exp = 0, sm_count = 148;
    __half *Ahi = nullptr, *Alo = nullptr, *Wst = nullptr;
    int *dev_ints = nullptr;   // [0] W max bits, [1] w_exp, [2] scratch max bits
    CUtensorMap map_ahi, map_alo, map_w;
    const float *src = nullptr;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled tc_encode_fn()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            throw CudaError{cudaErrorNotSupported, __FILE__, __LINE__};
        }
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

static inline void tc_make_map(CUtensorMap *map, void *base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner,
                               uint32_t box_outer)
{
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = tc_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        throw CudaError{cudaErrorInvalidValue, __FILE__, __LINE__};
    }
}

static inline bool tc_supported(int n, int m, int k)
{
    // shape limits of this engine: k <= 32; smaller problems than one wave gain nothing
    return k >= 1 && k <= TC_KP && n >= 64 && m >= 128;
}

static inline void tc_release(TcMatrix &tc)
{
    if (tc.Ahi) cudaFree(tc.Ahi);
    if (tc.Alo) cudaFree(tc.Alo);
    if (tc.Wst) cudaFree(tc.Wst);
    if (tc.dev_ints) cudaFree(tc.dev_ints);
    tc = TcMatrix();
}

// builds the split copy of A and the tensor maps (once per resident matrix)
static inline void tc_prepare(TcMatrix &tc, const float *dA, size_t ldA, int n, int m, double a_max, int sm_count, cudaStream_t st)
{
    if (tc.ready && tc.src == dA && tc.n == n && tc.m == m) return;
    tc_release(tc);
    tc.n = n; tc.m = m; tc.sm_count = sm_count; tc.src = dA;
    tc.n_pad = round_up(n, 64);
    const size_t cnt = (size_t)m * tc.n_pad;
    CUDA_CHECK(cudaMalloc(&tc.Ahi, cnt * sizeof(__half)));
    CUDA_CHECK(cudaMalloc(&tc.Alo, cnt * sizeof(__half)));
    CUDA_CHECK(cudaMalloc(&tc.Wst, (size_t)64 * tc.n_pad * sizeof(__half)));
    CUDA_CHECK(cudaMalloc(&tc.dev_ints, 4 * sizeof(int)));
    CUDA_CHECK(cudaMemsetAsync(tc.dev_ints, 0, 4 * sizeof(int), st));
    tc.a_exp = 0;
    if (a_max > 0) {
        int e;
        frexp(a_max, &e);
        tc.a_exp = std::max(-100, std::min(100, 14 - e));
    }
    tc_split_a_kernel<<<std::min(m, sm_count * 16), 256, 0, st>>>(dA, ldA, n, m, tc.n_pad, (float)ldexp(1.0, tc.a_exp), tc.Ahi, tc.Alo);
    CUDA_CHECK(cudaGetLastError());
    const uint64_t pitch = (uint64_t)tc.n_pad * sizeof(__half);
    tc_make_map(&tc.map_ahi, tc.Ahi, (uint64_t)tc.n_pad, (uint64_t)m, pitch, TC_CHUNK, TC_GROUP);
    tc_make_map(&tc.map_alo, tc.Alo, (uint64_t)tc.n_pad, (uint64_t)m, pitch, TC_CHUNK, TC_GROUP);
    tc_make_map(&tc.map_w, tc.Wst, (uint64_t)tc.n_pad, 64, pitch, TC_CHUNK, 64);
    CUDA_CHECK(cudaFuncSetAttribute(tc_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::total + 1024));
    CUDA_CHECK(cudaStreamSynchronize(st));
    tc.ready = true;
}

// W changed: re-split it (W max -> exponent -> stacked fp16 operand)
static inline void tc_update_w(TcMatrix &tc, const float *Wt, cudaStream_t st)
{
    CUDA_CHECK(cudaMemsetAsync(tc.dev_ints, 0, sizeof(int), st));
    tc_absmax_kernel<<<32, 256, 0, st>>>(Wt, (size_t)tc.n * TC_KP, tc.dev_ints);
    tc_split_w_kernel<<<(tc.n_pad + 127) / 128, 128, 0, st>>>(Wt, tc.n, tc.n_pad, tc.dev_ints, tc.dev_ints + 1, tc.Wst);
    CUDA_CHECK(cudaGetLastError());
}

// one pass.  mode 0: H update (phase 1 + solve) followed by B = A H' of the new H; mode 1: B = A H' only.
// B must be zeroed by the caller.  launches: 1 (+3 for the W split in mode 0)
static inline void tc_pass(TcMatrix &tc, int mode, const float *Wt, float *H, const float *G, const float *G0, float l1, int k,
                           int max_iter, float rel_tol, float *B, FitScalars *scal, cudaStream_t st)
{
    if (mode == 0) tc_update_w(tc, Wt, st);
    TcPassParams p;
    p.n = tc.n; p.m = tc.m; p.k = k;
    p.n_mtiles = (tc.n + 127) / 128;
    p.n_chunks = 2 * p.n_mtiles;
    p.mode = mode;
    p.G = G; p.G0 = G0; p.l1 = l1; p.max_iter = max_iter; p.rel_tol = rel_tol;
    p.H = H; p.B = B; p.scal = scal;
    p.w_exp = tc.dev_ints + 1;
    p.a_exp = tc.a_exp;
    p.groups_total = (tc.m + TC_GROUP - 1) / TC_GROUP;
    const int grid = std::min(tc.sm_count, p.groups_total);
    tc_pass_kernel<<<grid, TC_THREADS, TcSmem::total + 1024, st>>>(tc.map_ahi, tc.map_alo, tc.map_w, p);
    CUDA_CHECK(cudaGetLastError());
}
static inline int tc_pass_launches(int mode) { return mode == 0 ? 4 : 1; }

}  // namespace swne
