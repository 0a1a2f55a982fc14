// pass_tc.cuh -- tcgen05 engine (SWNE_ENGINE_TC) of the mse path: one persistent, warp-specialised kernel
// per outer iteration that streams the cells x genes matrix through TMA-staged shared-memory tiles twice:
//
//   phase 1  C = A Wt          (cells x k)   M = 128 cells, K = genes; the accumulator of a tile (TMEM) is turned in
//            place into the mu' of the per-cell SCD NNLS solve (scd_ls_update), which runs as a tensor-core block
//            Gauss-Seidel on the TMEM-resident tile; H is written back;
//   phase 3  B += A' H         (genes x k)   M = 128 genes (MN-major view of the SAME kind of tile), K = cells;
//            accumulators for 3 x 128 genes at a time live in TMEM (two regions), so phase 3 runs in 8 gene
//            sub-passes, each reading only those genes of every cell row; drained with vector red.add to global B.
//
// Operands are "split fp16": x = hi + lo with hi = fp16(x), lo = fp16(x - hi), after a power-of-two scaling,
// and every product is hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM (the fp32-emulating scheme the
// north star allows; error ~2^-21).  A is stored pre-split (2 + 2 bytes per entry = the 4 bytes of fp32), so
// the algorithmic HBM traffic is unchanged and no SIMT split pass touches the streamed tile.
//
// Why two reads of A per iteration instead of one: the per-cell solve runs 35..50 sweeps in the iterations that
// matter; keeping a cell's 12 KB row on chip (smem or L2) for that long for enough cells in flight needs > 200 MB
// chip-wide.  See DESIGN.md section 5.
//
// Reference semantics restated: update() / scd_ls_update of NNLM, call sites
// /root/reference/R/nmf.R:117,129.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <vector>

#include "common.cuh"
#include "kernels_simt.cuh"

namespace swne {

constexpr int TC_KP = 32;             // padded factor count handled by this engine (k <= 32)
constexpr int TC_TILE = 128;          // cells per tile (MMA M in phase 1, K in phase 3)
constexpr int TC_CHUNK = 64;          // genes per stage (128 B of fp16: one swizzle-128B row)
constexpr int TC_STAGES = 4;                  // A stages of 16 KB hi + 16 KB lo (+ 8 KB of W in phase 1)
constexpr int TC_GROUP = 32;          // cells per TMA box / work granule
constexpr int TC_A_STAGE = TC_TILE * TC_CHUNK * 2;      // 16 KB
constexpr int TC_W_STAGE = 64 * TC_CHUNK * 2;           // 8 KB  (rows 0..31 W_hi, 32..63 W_lo)
constexpr int TC_BOX_BYTES = TC_GROUP * TC_CHUNK * 2;   // 4 KB
constexpr int TC_H_SUB = TC_KP * 64 * 2;                // 4 KB: [32 factors][64 cells] 16-bit, K-major
constexpr int TC_H_TILE = 4 * TC_H_SUB;                 // hi cells 0..63, hi cells 64..127, lo, lo (all fp16)
constexpr int TC_CBUF = 5;                              // C / mu' accumulators (32 TMEM columns each) in flight
// TMEM map (512 columns): [0,160) five C / mu' tiles, [160,192) the H H' accumulator of the CTA, [192,320) the delta
// operands of the 4 solver groups (16 hi + 16 lo columns each), [320,512) two phase-3 accumulator regions of 3 x 32.
constexpr int TC_TGROUPS = 4;                           // solver groups (4 warps each) of the fused pass
constexpr int TC_GRAM_COL = 160;
constexpr int TC_DELTA_COL = 192;
constexpr int TC_D_COL = 320;
constexpr int TC_MT_PER_PASS = 3;                       // 128-gene accumulators per phase-3 unit (3 * 32 TMEM columns, two regions)
// device-side state of a fit (ints): scales and counters that used to be separate memset / absmax / copy launches
//   [0] max(W) float bits of the running W-side kernel   [1] fp16 scale exponent of the current split image of W
//   [2,3] max(H) float bits of the even / odd iteration   [4] fp16 overflow flag (H images)
//   [5] ticket counter of the W-side kernel   [6] ticket counter of the pass kernel
constexpr int TC_ST_WMAX = 0, TC_ST_WEXP = 1, TC_ST_HMAX = 2, TC_ST_OVF = 4, TC_ST_CNT_W = 5, TC_ST_CNT_P = 6, TC_ST_INTS = 8;
constexpr int TC_GS = TC_KP * TC_KP + TC_KP;            // doubles of one Gram + sums buffer
constexpr int TC_MAX_TILES = 128;                       // tiles per CTA (one mbarrier each)
constexpr int TC_GOP = 1024;                            // one [32 x 8] tf32 operand of G' (per block, hi / lo)

struct TcSmem {
    // offsets from the 1024-aligned base
    static constexpr int a_hi = 0;
    static constexpr int a_lo = a_hi + TC_STAGES * TC_A_STAGE;
    static constexpr int w = a_lo + TC_STAGES * TC_A_STAGE;
    // The two H tile images of phase 3 overlay the W stages of phase 1 (4 x 8 KB = 2 x 16 KB): the MMA / TMA warps run
    // phase 1 to completion before phase 3 starts, so the two are never live together (the loader waits for P1DONE).
    // This keeps the CTA at ~170 KB of shared memory, i.e. under the 196 KB carve-out, which leaves 60 KB of L1 for the
    // solver's local-memory traffic (with 28 KB of L1 its spills missed half of the time, round 2 ncu).
    static constexpr int h = w;
    static_assert(2 * TC_H_TILE <= TC_STAGES * TC_W_STAGE, "H images must fit the W stage area");
    static constexpr int gops = w + TC_STAGES * TC_W_STAGE;       // 8 tf32 operands of G' (8 KB)
    static constexpr int dvec = gops + 8 * TC_GOP;       // sqrt(diag), 1/sqrt(diag)
    static constexpr int bars = dvec + 2 * TC_KP * 4;
    static constexpr int n_bars = 2 * TC_STAGES + 2 * TC_CBUF + 2 + 2 + 4 + TC_TGROUPS + 1 + TC_MAX_TILES;
    static constexpr int misc = bars + n_bars * 8;       // [0] TMEM base, [4] last-CTA flag, [64,192) row sums of H (fp32)
    static constexpr int total = misc + 256;
};

struct TcPassParams {
    int n, m, k;
    int n_chunks;            // gene chunks per tile in phase 1 (even, >= ceil(n/64))
    int n_mtiles;            // ceil(n/128)
    int mode;                // 0: phase 1 + solve + phase 3; 1: phase 3 only (A H' of the given H)
    int par;                 // iteration parity: which half of the double-buffered state this launch writes
    int finalize;            // 1: the last CTA turns H H' into the GramPack of the next W update (single GPU)
    const float *pack;       // GramPack of W'W (D / invD / G' operands read here; the triangular part sits in c_gram)
    float *pack_next;        // GramPack of H H' + alpha penalties, written by the last CTA (finalize)
    double pen_next0, pen_next1;   // alpha[0], alpha[1] of that GramPack
    float l1;
    float pen_diag, pen_all; // G = G0 + pen_diag I + pen_all 11' (+ tiny I): the L2 / angle penalties folded into the Gram
    int max_iter;
    float rel_tol;
    float *H;                // [m][32] fp32 in/out
    float *Cscr;             // [m][32] fp32 scratch: the contraction C = A W, kept for the loss after the solve
    float *B;                // [n][32] fp32, accumulated with red.add (zeroed by the W-side kernel / the caller)
    uint8_t *Himg;           // [grid][tiles_per_cta][TC_H_TILE] split 16-bit images of the H tiles
    int tiles_per_cta;
    FitScalars *scal;
    int *state;              // TC_ST_* ints
    double *G64h;            // [2][TC_GS]: H H' and row sums of H, accumulated into half `par`; half par^1 is zeroed
    int a_exp;               // A was scaled by 2^a_exp before the split
    int groups_total;        // ceil(m / 32)
    unsigned long long *dbg; // optional timeline (SWNE_TC_TIMELINE): 64 globaltimer slots per CTA
    long long *clk;          // optional solver clock ring (SWNE_TC_CLOCKS): [cta][group][TC_CLK_RECS][8] clock64 stamps
};
constexpr int TC_CLK_RECS = 1024;   // records per (CTA, solver group): one per block update / per tile

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TC_MARK(slot) do { if (p.dbg) p.dbg[(size_t)blockIdx.x * 64 + (slot)] = gtimer(); } while (0)
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
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major, int b_is_bf16 = 0)
{
    return (1u << 4) /* C = F32 */ | (0u << 7) /* A = F16 */ | ((uint32_t)b_is_bf16 << 10) /* B = F16 | BF16 */ |
           ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// no-swizzle K-major descriptor: 8-row x 16 B core matrices, LBO = distance of the two K halves, SBO = 8-row groups
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N)
{
    return (1u << 4) /* C = F32 */ | (2u << 7) /* A = TF32 */ | (2u << 10) /* B = TF32 */ | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from TMEM (lane = row, one 32-bit column per K element)
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&a)[8], const float (&b)[8])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(a[4])), "r"(__float_as_uint(a[5])), "r"(__float_as_uint(a[6])), "r"(__float_as_uint(a[7])),
        "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])), "r"(__float_as_uint(b[2])), "r"(__float_as_uint(b[3])),
        "r"(__float_as_uint(b[4])), "r"(__float_as_uint(b[5])), "r"(__float_as_uint(b[6])), "r"(__float_as_uint(b[7]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_st16_nowait(uint32_t taddr, const float (&a)[8], const float (&b)[8])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(a[4])), "r"(__float_as_uint(a[5])), "r"(__float_as_uint(a[6])), "r"(__float_as_uint(a[7])),
        "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])), "r"(__float_as_uint(b[2])), "r"(__float_as_uint(b[3])),
        "r"(__float_as_uint(b[4])), "r"(__float_as_uint(b[5])), "r"(__float_as_uint(b[6])), "r"(__float_as_uint(b[7]))
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const float (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned named_bar_or(int id, int nthreads, unsigned pred)
{
    unsigned out;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "bar.red.or.pred p, %2, %3, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(out)
        : "r"(pred), "r"(id), "r"(nthreads)
        : "memory");
    return out;
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Tensor-core block Gauss-Seidel for one tile of 128 columns (4 warps = the 4 TMEM lane quarters).
//   mu' of the tile sits in TMEM ([128 lanes x 32 columns] at L.tmu, this thread's row at trow); h' in registers.
//   Per block of 16 coordinates: read the 16 mu' values, run the 16 sequential coordinate steps with the triangular
//   in-block corrections in registers, write the 16 deltas (tf32 hi | lo) to the TMEM operand columns and let one
//   elected thread issue the 3xTF32 rank-16 update  Mu += D G'[block rows, :]  (A from TMEM, B = the G' operands in
//   smem).  Exactly the deferred rank-1 updates of scd_ls_update: entries outside the block are not read before the
//   block ends.  Converged columns ride along with delta = 0.
//   The same "push" (gs_push) also builds mu' = G' h' + ... at the start of a solve and G' h' for the loss at its end,
//   so the O(k^2) matrix-vector products of a column never run on the CUDA cores.
// Round-2 measurements behind this form (tools/micro/gs_bench.cu, profiles/r2_gs_bench_*.txt): a tcgen05.mma issued under
// `if (lane == 0)` compiles to an ELECT/BRA.U.ANY loop per instruction (~80 cycles each); under elect.sync in warp-uniform
// control flow the six UTCHMMA issue back to back.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}

struct GsLink {
    uint32_t tmu;       // TMEM address of the tile's mu' accumulator (column base, lane 0)
    uint32_t tdelta;    // TMEM operand columns of this group: 16 hi + 16 lo
    uint32_t gop_u;     // shared-memory address of the tf32 operands of G'
    uint32_t mbar;      // "block MMA done" barrier of this group
    uint32_t parity;
    bool pending;
    int bar_id;         // named barrier of the group's 128 threads
    int quarter;        // TMEM lane quarter of this warp
};

__device__ __forceinline__ void gs_wait(GsLink &L)
{
    if (L.pending) { mbar_wait(L.mbar, L.parity); L.parity ^= 1u; L.pending = false; }
    tc_fence_after();
}

// Mu (+)= V G'[16 BLK .. 16 BLK + 15, :] for the 16 values v of every column of the tile.  OVERWRITE: the tile is
// replaced instead of accumulated (only with BLK == 0).  The previous push must have been waited for (gs_wait).
template <int BLK, bool OVERWRITE>
__device__ __forceinline__ void gs_push(GsLink &L, const float (&h)[32], const float (&d)[16], bool from_h)
{
    constexpr uint32_t ID_TF32 = idesc_tf32(128, 32);
    {
        // columns [0,16): hi (tf32-exact), [16,32): lo -- stored in two halves: 16 live registers instead of 32 next to h'
        const uint32_t tq = L.tdelta + ((uint32_t)(L.quarter * 32) << 16);
        float a8[8], b8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a8[i] = __uint_as_float(__float_as_uint(from_h ? h[16 * BLK + i] : d[i]) & 0xffffe000u);
            b8[i] = __uint_as_float(__float_as_uint(from_h ? h[16 * BLK + 8 + i] : d[8 + i]) & 0xffffe000u);
        }
        tc_st16_nowait(tq, a8, b8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a8[i] = (from_h ? h[16 * BLK + i] : d[i]) - a8[i];
            b8[i] = (from_h ? h[16 * BLK + 8 + i] : d[8 + i]) - b8[i];
        }
        tc_st16_nowait(tq + 16, a8, b8);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    named_bar_sync(L.bar_id, 128);
    if (L.quarter == 0) {
        if (elect_one_sync()) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int b8 = 2 * BLK + ks;     // 8-row block of G'
                const uint64_t g_hi = desc_nosw(L.gop_u + (2 * b8) * TC_GOP, 128, 256);
                const uint64_t g_lo = desc_nosw(L.gop_u + (2 * b8 + 1) * TC_GOP, 128, 256);
                tc_mma_tf32_ts(L.tmu, L.tdelta + 8 * ks, g_hi, ID_TF32, (OVERWRITE && ks == 0) ? 0 : 1);
                tc_mma_tf32_ts(L.tmu, L.tdelta + 8 * ks, g_lo, ID_TF32, 1);
                tc_mma_tf32_ts(L.tmu, L.tdelta + 16 + 8 * ks, g_hi, ID_TF32, 1);
            }
            tc_commit(L.mbar);
        }
        __syncwarp();
    }
    L.pending = true;
}

// mu' += G' h'  (start of a solve; mu' holds (l1 - c) / d) or  mu' = G' h'  (OVERWRITE: for the loss)
template <bool OVERWRITE>
__device__ __forceinline__ void gs_push_h(GsLink &L, const float (&h)[32], int k)
{
    const float none[16] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    gs_wait(L);
    gs_push<0, OVERWRITE>(L, h, none, true);
    if (16 < k) {
        gs_wait(L);
        gs_push<1, false>(L, h, none, true);
    }
}

template <int BLK>
__device__ __forceinline__ void gs_block(float (&h)[32], uint32_t trow, GsLink &L, bool any_change, float rel_tol, unsigned &changed,
                                         long long *clk, int *clk_n)
{
    long long *rec = nullptr;
#ifdef TC_SOLVER_CLOCKS
    if (clk && *clk_n < TC_CLK_RECS) { rec = clk + 8 * (*clk_n)++; rec[0] = clock64(); rec[7] = 1; }
#endif
    gs_wait(L);
    if (rec) rec[1] = clock64();
    float m16[16], d[16];
    tc_ld16(trow + 16 * BLK, m16);
    if (rec) rec[2] = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int kk = 16 * BLK + i;
        const float tmp = fmaxf(h[kk] - m16[i], 0.f);
        d[i] = tmp - h[kk];
        if (any_change) changed |= __float_as_uint(d[i]);
        else changed |= (2.f * fabsf(d[i]) > rel_tol * (tmp + h[kk] + TINY_NUM_F)) ? 1u : 0u;
        h[kk] = tmp;
#pragma unroll
        for (int j = i + 1; j < 16; ++j) m16[j] = fmaf(d[i], c_gram[kk * TC_KP + 16 * BLK + j], m16[j]);
    }
    if (rec) rec[3] = clock64();
    gs_push<BLK, false>(L, h, d, false);
    if (rec) { rec[4] = rec[5] = rec[6] = clock64(); }
}

// the sweeps of scd_ls_update on the tile; returns this column's sweep count
__device__ __forceinline__ int tc_block_gs(float (&h)[32], uint32_t trow, GsLink &L, int k, int max_iter, float rel_tol,
                                           long long *clk = nullptr, int *clk_n = nullptr)
{
    const bool any_change = rel_tol < 2.9e-8f;
    int my_sweeps = 1;
    unsigned any = 1u;
    for (int sweep = 0; sweep < max_iter && any; ++sweep) {
        unsigned changed = 0u;
        gs_block<0>(h, trow, L, any_change, rel_tol, changed, clk, clk_n);
        if (16 < k) gs_block<1>(h, trow, L, any_change, rel_tol, changed, clk, clk_n);
        any = named_bar_or(L.bar_id, 128, changed);
        if (changed) my_sweeps = min(sweep + 2, max_iter);
    }
    return my_sweeps;
}

// tf32 operands of G' for the block updates: op[blk8][hi|lo][n][kq] = split(G'[8 blk8 + kq][n]), [32 x 8] K-major, no
// swizzle: 8-row x 16 B core matrices, the two K halves 128 B apart, 8-row groups 256 B apart
__device__ __forceinline__ void tc_fill_gops(uint8_t *gops, const float *__restrict__ pack, int tid, int nthreads)
{
    for (int t = tid; t < 4 * 32 * 8; t += nthreads) {
        const int blk = t >> 8, n = (t >> 3) & 31, kq = t & 7;
        const float v = pack[(8 * blk + kq) * TC_KP + n];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const int off = (n >> 3) * 256 + (kq >> 2) * 128 + (n & 7) * 16 + (kq & 3) * 4;
        *reinterpret_cast<float *>(gops + (2 * blk) * TC_GOP + off) = hi;
        *reinterpret_cast<float *>(gops + (2 * blk + 1) * TC_GOP + off) = v - hi;
    }
}

// fp16 scale exponent of a factor whose largest entry was `max_bits` (float bits) one iteration ago: that maximum maps
// to [8, 16), which leaves a factor 4096 of growth before fp16 overflows (flagged, never silent) while entries down to
// 1e-9 of the maximum keep their value in the hi + lo pair
__device__ __forceinline__ int tc_exp_of(int max_bits)
{
    const float mx = __int_as_float(max_bits);
    if (!(mx > 0.f)) return 0;
    int e;
    frexpf(mx, &e);
    return max(-100, min(100, 4 - e));
}

// GramPack from the fp64 sums, by one whole CTA (finalize_gram_kernel of kernels_misc.cuh restated as a device function;
// update() of NNLM: G = G0 + (p0 - p1) I + p1 11' + tiny I).  G64 is read through L2 (other CTAs' atomics).
__device__ __forceinline__ void tc_finalize_gram(const double *G64, float *pack, int k, double p0, double p1, float *sInv)
{
    constexpr int KP = TC_KP;
    for (int t = threadIdx.x; t < KP; t += blockDim.x) {
        double g = 1.0;
        if (t < k) {
            g = __ldcg(G64 + t * KP + t);
            if (p0 != p1) g += p0 - p1;
            if (p1 != 0.0) g += p1;
            g += TINY_NUM_D;
        }
        const float d = sqrtf((float)g);
        pack[2 * KP * KP + t] = d;
        pack[2 * KP * KP + KP + t] = 1.f / d;
        sInv[t] = 1.f / d;
        pack[2 * KP * KP + 2 * KP + t] = (t < k) ? (float)__ldcg(G64 + KP * KP + t) : 0.f;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < KP * KP; t += blockDim.x) {
        const int a = t / KP, b = t % KP;
        float gs = 0.f, g0 = 0.f;
        if (a < k && b < k) {
            double g = __ldcg(G64 + t);
            g0 = (float)g;
            if (p1 != 0.0) g += p1;
            gs = (a == b) ? 1.f : (float)g * sInv[a] * sInv[b];
        }
        pack[t] = gs;
        pack[KP * KP + t] = g0;
    }
}

// "last CTA done": every CTA calls this once after its global writes; true in exactly one CTA, after all the others'
// writes are visible.  The counter is reset for the next launch.
__device__ __forceinline__ bool tc_last_cta(int *counter, int *s_flag)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(counter, 1);
        const int last = (ticket == (int)gridDim.x - 1) ? 1 : 0;
        if (last) { *counter = 0; __threadfence(); }
        *s_flag = last;
    }
    __syncthreads();
    return *s_flag != 0;
}

// ------------------------------------------------------------------------------------------------
// The W side of one outer iteration in ONE kernel (one CTA of 4 warps per 128 genes):
//   solve the 128 columns with the tensor-core sweeps (X [cols][32] in/out, C = A H' the contraction, pack = GramPack of
//   H H', c_gram its triangular coefficients), write W, add this tile's W'W, column sums and max into the device state,
//   zero the consumed rows of C (the pass accumulates the next A H' into them).  The GramPack of the H update and the
//   split-fp16 image of W (exact power-of-two scale from the global max -- a scale guessed from the previous iteration would
//   cost the small entries of W five bits, measured as 2.5x the W/H error) are built by tc_wfinish_kernel right behind it.
// Together the two replace, per iteration: finalize_gram, tc_wsolve, zero_fill, gram(W), finalize_gram, absmax, split_w and
// three memsets (round 1: 0.15 ms of serial launches).
// ------------------------------------------------------------------------------------------------
struct TcWsideParams {
    float *X;                 // Wt [n][32]
    float *C;                 // B [n][32] = A H' (all-reduced); rows zeroed after they are read
    const float *pack;        // GramPack of H H' (alpha penalties)
    float *pack_next;         // GramPack of W'W (beta penalties), written by CTA 0 of tc_wfinish_kernel
    double pen_next0, pen_next1;
    float l1;
    int k, cols, n_pad, max_iter;
    float rel_tol;
    __half *Wst;              // [64][n_pad] split image
    int *state;
    int par;
    double *G64w;             // [2][TC_GS]
    float *partial;           // [grid][TC_GS] per-CTA partial W'W and column sums
    FitScalars *scal;
    unsigned long long *dbg;  // optional timeline (SWNE_TC_TIMELINE): 16 globaltimer slots per CTA
};
#define TCW_MARK(slot) do { if (p.dbg && threadIdx.x == 0) p.dbg[(size_t)blockIdx.x * 16 + (slot)] = gtimer(); } while (0)

constexpr int TC_WSIDE_THREADS = 256;     // warps 0-3 solve (one per TMEM lane quarter), warps 4-7 help with the Gram and the tail
__global__ void __launch_bounds__(TC_WSIDE_THREADS, 1) tc_wside_kernel(const TcWsideParams p)
{
    __shared__ __align__(1024) uint8_t gops[8 * TC_GOP];
    __shared__ float sD[2 * TC_KP];
    __shared__ __align__(16) float ws[128][TC_KP + 2];     // the tile's W rows (stride 34: 8-byte aligned pairs); reused as the split staging
    __shared__ __align__(8) uint64_t mbar_store;
    __shared__ uint32_t tmem_slot;
    __shared__ int s_flag;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31, quarter = warp & 3;
    TCW_MARK(0);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const uint32_t mbar = smem_u32(&mbar_store);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fill_gops(gops, p.pack, threadIdx.x, blockDim.x);
    if (threadIdx.x < 2 * TC_KP) sD[threadIdx.x] = p.pack[2 * TC_KP * TC_KP + threadIdx.x];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const float *sInvD = sD + TC_KP;
    TCW_MARK(1);
    if (warp < 4) {
        const int j = blockIdx.x * 128 + threadIdx.x;
        const bool valid = j < p.cols;
        float *xrow = p.X + (size_t)(valid ? j : 0) * TC_KP;
        const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
        GsLink L{tmem, tmem + 32, smem_u32(gops), mbar, 0u, false, 1, quarter};
        float h[32];
        {
            // mu'_0 = (l1 - c) / d; the G' h' part is added by the tensor core
            float mv[32];
            if (valid) {
                float4 *cp = reinterpret_cast<float4 *>(p.C + (size_t)j * TC_KP);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float4 v = reinterpret_cast<const float4 *>(xrow)[r];
                    h[4 * r] = v.x * sD[4 * r]; h[4 * r + 1] = v.y * sD[4 * r + 1];
                    h[4 * r + 2] = v.z * sD[4 * r + 2]; h[4 * r + 3] = v.w * sD[4 * r + 3];
                    const float4 c = cp[r];
                    cp[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    mv[4 * r] = (p.l1 - c.x) * sInvD[4 * r]; mv[4 * r + 1] = (p.l1 - c.y) * sInvD[4 * r + 1];
                    mv[4 * r + 2] = (p.l1 - c.z) * sInvD[4 * r + 2]; mv[4 * r + 3] = (p.l1 - c.w) * sInvD[4 * r + 3];
                }
            } else {
#pragma unroll
                for (int f = 0; f < 32; ++f) { h[f] = 0.f; mv[f] = 0.f; }
            }
            tc_st32(trow, mv);
        }
        gs_push_h<false>(L, h, p.k);
        TCW_MARK(2);
        const int my_sweeps = tc_block_gs(h, trow, L, p.k, p.max_iter, p.rel_tol);
        gs_wait(L);
        TCW_MARK(3);
        // ---- W row, the tile's maximum and rows for the Gram
        float wmax = 0.f;
#pragma unroll
        for (int f = 0; f < 32; ++f) {
            h[f] *= sInvD[f];
            ws[threadIdx.x][f] = h[f];
            wmax = fmaxf(wmax, h[f]);
        }
        if (valid) {
#pragma unroll
            for (int r = 0; r < 8; ++r) reinterpret_cast<float4 *>(xrow)[r] = make_float4(h[4 * r], h[4 * r + 1], h[4 * r + 2], h[4 * r + 3]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        int sw = valid ? my_sweeps : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sw += __shfl_xor_sync(0xffffffffu, sw, o);
        if (lane == 0) {
            atomicMax(p.state + TC_ST_WMAX, __float_as_int(wmax));
            atomicAdd(&p.scal->sweeps_w, (unsigned long long)sw);
        }
        tc_fence_before();
    }
    __syncthreads();
    {
        // W'W of the 128 rows: thread e owns entries e, e + 256, ... of the 32 x 32 block (a = entry / 32 is warp-uniform).
        // Every CTA leaves its partial sums in its own slot (no atomics: 24 CTAs finishing together on the same 1056
        // addresses cost 6 us); the last CTA adds the slots up in fp64.
        float *slot = p.partial + (size_t)blockIdx.x * TC_GS;
        // 2 x 2 register blocks: thread t owns rows a0, a0 + 1 and columns b0, b0 + 1 of the 32 x 32 Gram (two 8-byte
        // shared-memory loads feed four FMAs; the one-entry-per-thread form was bound by the shared-memory pipe: 6 us)
        const int a0 = 2 * (threadIdx.x >> 4), b0 = 2 * (threadIdx.x & 15);
        float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f, csum = 0.f;
#pragma unroll 8
        for (int r = 0; r < 128; ++r) {
            const float2 wa = *reinterpret_cast<const float2 *>(&ws[r][a0]);
            const float2 wb = *reinterpret_cast<const float2 *>(&ws[r][b0]);
            g00 = fmaf(wa.x, wb.x, g00); g01 = fmaf(wa.x, wb.y, g01);
            g10 = fmaf(wa.y, wb.x, g10); g11 = fmaf(wa.y, wb.y, g11);
        }
        slot[a0 * TC_KP + b0] = g00; slot[a0 * TC_KP + b0 + 1] = g01;
        slot[(a0 + 1) * TC_KP + b0] = g10; slot[(a0 + 1) * TC_KP + b0 + 1] = g11;
        // column sums: thread (part, f) adds 16 rows, the first warp folds the 8 parts
        {
            const int f = threadIdx.x & 31, part = threadIdx.x >> 5;
            for (int r = 16 * part; r < 16 * part + 16; ++r) csum += ws[r][f];
            __syncthreads();
            float *cpart = &ws[0][0];          // the rows are consumed: reuse the first 256 floats
            cpart[part * 32 + f] = csum;
            __syncthreads();
            if (threadIdx.x < TC_KP) {
                float cs = 0.f;
#pragma unroll
                for (int q = 0; q < TC_WSIDE_THREADS / 32; ++q) cs += cpart[q * 32 + threadIdx.x];
                slot[TC_KP * TC_KP + threadIdx.x] = cs;
            }
        }
    }
    TCW_MARK(4);
    __syncthreads();
    TCW_MARK(8);
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
    }
}

// The W side's second, tiny kernel (one CTA per 128 genes, right behind tc_wside_kernel on the stream): CTA 0 adds up the
// per-CTA partial sums into W'W (fp64) and builds the GramPack of the H update; every CTA writes the split-fp16 image of
// its 128 genes with the exact power-of-two scale of max(W).  As the serial tail of the solve kernel's last CTA this
// took 45 us (L2 round trips of a single CTA); spread over the CTAs it is a few microseconds.
__global__ void __launch_bounds__(1024, 1) tc_wfinish_kernel(const TcWsideParams p, int n_slots)
{
    __shared__ __align__(16) __half stg[64 * 128];
    __shared__ float sInv[64];
    // ---- split image of this CTA's genes: Wst [64][n_pad] (rows 0..31 hi, 32..63 lo), max(W) -> [2^13, 2^14)
    int w_exp = 0;
    {
        const float mx = __int_as_float(p.state[TC_ST_WMAX]);
        if (mx > 0.f) {
            int e;
            frexpf(mx, &e);
            w_exp = max(-100, min(100, 14 - e));
        }
    }
    const float scale = exp2f((float)w_exp);
    {
        const int gi = threadIdx.x >> 3, fq = threadIdx.x & 7;    // gene within the tile, which 4 factors
        const int g = blockIdx.x * 128 + gi;
        const float4 v = (g < p.cols) ? reinterpret_cast<const float4 *>(p.X + (size_t)g * TC_KP)[fq] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int f = 4 * fq + e;
            const float x = w4[e] * scale;
            const __half hi = __float2half_rn(x);
            stg[f * 128 + gi] = hi;
            stg[(TC_KP + f) * 128 + gi] = __float2half_rn(x - __half2float(hi));
        }
        __syncthreads();
        // one 16-byte store per thread: row = thread / 16, 8 genes per piece
        const int row = threadIdx.x >> 4, piece = threadIdx.x & 15;
        const int g0 = blockIdx.x * 128 + 8 * piece;
        if (g0 < p.n_pad)
            *reinterpret_cast<uint4 *>(p.Wst + (size_t)row * p.n_pad + g0) = *reinterpret_cast<const uint4 *>(stg + row * 128 + 8 * piece);
    }
    if (blockIdx.x != 0) return;
    // ---- CTA 0: W'W + column sums of the whole W (fp64, one entry per thread), the GramPack of the H update, state
    double *G = p.G64w + (size_t)p.par * TC_GS;
    for (int t = threadIdx.x; t < TC_GS; t += blockDim.x) {
        double acc = 0.0;
        for (int c0 = 0; c0 < n_slots; c0 += 8) {
            float v8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v8[u] = (c0 + u < n_slots) ? p.partial[(size_t)(c0 + u) * TC_GS + t] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += (double)v8[u];
        }
        const int a = t >> 5, b = t & 31;
        G[t] = (t < TC_KP * TC_KP) ? ((a < p.k && b < p.k) ? acc : 0.0) : ((t - TC_KP * TC_KP < p.k) ? acc : 0.0);
    }
    __threadfence();
    __syncthreads();
    tc_finalize_gram(G, p.pack_next, p.k, p.pen_next0, p.pen_next1, sInv);
    if (threadIdx.x == 0) {
        p.state[TC_ST_WEXP] = w_exp;
        p.state[TC_ST_HMAX + p.par] = 0;
        p.scal->loss_part = 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// the pass kernel (768 threads, one CTA per SM)
//
//   warp 0       TMA producer (one elected lane): phase 1 streams (A_hi, A_lo, W) chunks of every tile, phase 3 streams
//                (A_hi, A_lo) for one third-of-an-eighth of the genes at a time
//   warp 1       MMA issuer (one elected lane): phase 1 C[t] = A W' into one of 5 TMEM accumulators; phase 3 D += A' H,
//                and H H' of every solved tile from the same shared-memory image (one more accumulator)
//   warp 2       H-tile loader: copies the split 16-bit image of a solved tile from global scratch to smem
//   warps 4-7    drain: phase-3 accumulators -> red.add into B
//   warps 8-23   4 solver groups x 4 warps: tensor-core block Gauss-Seidel on tiles g, g+4, ...
// Registers are redistributed with setmaxnreg (launch: 80 per thread): service warps 40, drain 56, solvers 96 --
// with 80 the solver spilled 24 of its 32 h' values to local memory inside the sweep (round 1).
// Phase 3 of tile t only needs tile t's solve, so the second stream over A overlaps the remaining solves.
// ------------------------------------------------------------------------------------------------
constexpr int TC_REGS_SERVICE = 40, TC_REGS_DRAIN = 56, TC_REGS_SOLVER = 96;
static_assert(128 * TC_REGS_SERVICE + 128 * TC_REGS_DRAIN + 512 * TC_REGS_SOLVER <= 768 * 80, "register pool of the pass kernel");

__global__ void __launch_bounds__(768, 1)
tc_pass_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
               const __grid_constant__ CUtensorMap map_w, const TcPassParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    // broadcast from lane 0: the compiler then knows the warp index (and everything derived from it: the role, the
    // solver group, the TMEM lane quarter, barrier addresses) is warp-uniform and keeps it in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    // ---- my cells: a contiguous range of 32-cell groups
    const int gpc = p.groups_total / gridDim.x, grem = p.groups_total % gridDim.x;
    const int g_begin = blockIdx.x * gpc + min((int)blockIdx.x, grem);
    const int g_count = gpc + ((int)blockIdx.x < grem ? 1 : 0);
    const int cell_begin = g_begin * TC_GROUP;
    const int cell_end = min(p.m, (g_begin + g_count) * TC_GROUP);
    const int n_tiles = (cell_end - cell_begin + TC_TILE - 1) / TC_TILE;
    const int n_pass = (p.n_mtiles + TC_MT_PER_PASS - 1) / TC_MT_PER_PASS;
    uint8_t *my_img = p.Himg + (size_t)blockIdx.x * p.tiles_per_cta * TC_H_TILE;

    // ---- barriers
    const uint32_t bar0 = sbase + TcSmem::bars;
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (TC_STAGES + s); };
    auto CFULL = [&](int b) { return bar0 + 8u * (2 * TC_STAGES + b); };
    auto CEMPTY = [&](int b) { return bar0 + 8u * (2 * TC_STAGES + TC_CBUF + b); };
    constexpr int B2 = 2 * TC_STAGES + 2 * TC_CBUF;
    auto HFULL = [&](int b) { return bar0 + 8u * (B2 + b); };
    auto HEMPTY = [&](int b) { return bar0 + 8u * (B2 + 2 + b); };
    auto DFULL = [&](int r) { return bar0 + 8u * (B2 + 4 + r); };
    auto DEMPTY = [&](int r) { return bar0 + 8u * (B2 + 6 + r); };
    auto MBG = [&](int g) { return bar0 + 8u * (B2 + 8 + g); };            // "block MMA of group g done"
    const uint32_t P1DONE = bar0 + 8u * (B2 + 8 + TC_TGROUPS);              // all phase-1 MMAs complete (W stages free)
    auto TDONE = [&](int t) { return bar0 + 8u * (B2 + 8 + TC_TGROUPS + 1 + t); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcSmem::misc);
    int *s_flag = reinterpret_cast<int *>(smem + TcSmem::misc + 4);
    float *s_hsum = reinterpret_cast<float *>(smem + TcSmem::misc + 64);
    if (threadIdx.x < TC_KP) s_hsum[threadIdx.x] = 0.f;

    // zero the A stages once: rows of a partial tile that TMA never writes must hold finite values
    for (int i = threadIdx.x; i < (2 * TC_STAGES * TC_A_STAGE) / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem + TcSmem::a_hi)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int b = 0; b < TC_CBUF; ++b) { mbar_init(CFULL(b), 1); mbar_init(CEMPTY(b), 128); }
        for (int b = 0; b < 2; ++b) { mbar_init(HFULL(b), 1); mbar_init(HEMPTY(b), 1); }
        for (int r = 0; r < 2; ++r) { mbar_init(DFULL(r), 1); mbar_init(DEMPTY(r), 128); }
        for (int t = 0; t < n_tiles; ++t) mbar_init(TDONE(t), 128);
        for (int g = 0; g < TC_TGROUPS; ++g) mbar_init(MBG(g), 1);
        mbar_init(P1DONE, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    float *sD = reinterpret_cast<float *>(smem + TcSmem::dvec);   // D[32] | invD[32]
    if (p.mode == 0) {
        tc_fill_gops(smem + TcSmem::gops, p.pack, threadIdx.x, blockDim.x);
        if (threadIdx.x < 2 * TC_KP) sD[threadIdx.x] = p.pack[2 * TC_KP * TC_KP + threadIdx.x];
    }
    // generic-proxy zero fill must be visible to the async proxy (TMA / UMMA)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) TC_MARK(0);
    const uint32_t tmem_c = tmem;                       // columns [0,192): six C / mu' accumulators
    // phase-3 unit u (= wave * n_pass + gene sub-pass) accumulates into region u & 1 (3 x 32 columns each): the drain
    // of one unit overlaps the MMAs of the next
    auto tmem_d = [&](int u) { return tmem + TC_D_COL + (u & 1) * (TC_MT_PER_PASS * 32); };
    // phase 3 takes the tiles in waves (B += A'H is linear in the tiles and the drain is an atomic add), so the
    // second stream over A starts on the tiles whose solves are done while the last tiles are still being solved
    // Waves: pairs of tiles (a wave can only start when its tiles are solved; every wave drains all accumulators once
    // more, i.e. adds one more round of red.add traffic to L2).  Long CTAs (> 8 tiles) take everything but their last
    // 4 tiles as one leading wave.
    const int w_base = (n_tiles > 8) ? n_tiles - 4 : 0;                   // tiles [0, w_base): the leading wave
    const int n_waves = (p.mode != 0 || n_tiles <= 2) ? 1 : (w_base > 0 ? 1 : 0) + (n_tiles - w_base + 1) / 2;
    auto wave_lo = [&](int w) {
        if (n_waves == 1) return w == 0 ? 0 : n_tiles;
        if (w_base > 0) { if (w == 0) return 0; --w; }
        return min(n_tiles, w_base + 2 * w);
    };

    // fp16 scale of the H images: the previous max(H) maps to [8, 16), which leaves a factor 4096 of growth before fp16
    // overflows (flagged, never silent) while entries down to 1e-9 of the max keep their value
    const int h_exp = tc_exp_of(p.state[TC_ST_HMAX + (p.par ^ 1)]);

    if (warp >= 8) {
        // =========================== solver warps ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_SOLVER));
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const float descale = exp2f(-(float)(p.a_exp + p.state[TC_ST_WEXP]));
        const float hscale = exp2f((float)h_exp);
        const float *sInvD = sD + TC_KP;
        // Group g (4 warps = the 4 lane quarters of a tile) solves tiles g, g+4, ...  mu' of the tile lives in its
        // TMEM accumulator (the C tile is rewritten in place).
        const int grp = (warp - 8) >> 2;
        const int lc = quarter * 32 + lane;
        GsLink L{tmem_c, tmem + TC_DELTA_COL + 32 * grp, sbase + TcSmem::gops, MBG(grp), 0u, false, 1 + grp, quarter};
        // optional clock ring of this group (thread 0 of the group writes it)
#ifdef TC_SOLVER_CLOCKS
        long long *clk = (p.clk && quarter == 0 && lane == 0) ? p.clk + ((size_t)blockIdx.x * TC_TGROUPS + grp) * TC_CLK_RECS * 8 : nullptr;
#else
        long long *clk = nullptr;   // build with -DTC_SOLVER_CLOCKS for the SWNE_TC_CLOCKS ring (tools/solver_clocks.py)
#endif
        int clk_n = 0;
        for (int t = grp; t < n_tiles; t += TC_TGROUPS) {
            const int b = t % TC_CBUF;
            const int cell = cell_begin + t * TC_TILE + lc;
            const bool valid = cell < cell_end;
            float *hrow = p.H + (size_t)(valid ? cell : cell_begin) * TC_KP;
            L.tmu = tmem_c + b * 32;
            const uint32_t trow = L.tmu + ((uint32_t)(quarter * 32) << 16);
            long long *trec = nullptr;
#ifdef TC_SOLVER_CLOCKS
            if (clk && clk_n < TC_CLK_RECS) { trec = clk + 8 * clk_n++; trec[0] = clock64(); trec[7] = 2; trec[6] = t; }
#endif
            if (lane == 0 && quarter == 0 && t < 8) TC_MARK(8 + t);
            float h[32];
            if (p.mode == 0) {
                mbar_wait(CFULL(b), (t / TC_CBUF) & 1);
                tc_fence_after();
                if (trec) trec[1] = clock64();
                {
                    // mu'_0 = (l1 - c) / d   (c = the contraction, kept in the scratch row for the loss); h' = d o h
                    float cv[32];
                    tc_ld32(trow, cv);
                    if (valid) {
                        float *crow = p.Cscr + (size_t)cell * TC_KP;
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const float4 v = reinterpret_cast<const float4 *>(hrow)[r];
                            h[4 * r] = v.x * sD[4 * r]; h[4 * r + 1] = v.y * sD[4 * r + 1];
                            h[4 * r + 2] = v.z * sD[4 * r + 2]; h[4 * r + 3] = v.w * sD[4 * r + 3];
                            const float c0 = cv[4 * r] * descale, c1 = cv[4 * r + 1] * descale, c2 = cv[4 * r + 2] * descale,
                                        c3 = cv[4 * r + 3] * descale;
                            reinterpret_cast<float4 *>(crow)[r] = make_float4(c0, c1, c2, c3);
                            cv[4 * r] = (p.l1 - c0) * sInvD[4 * r]; cv[4 * r + 1] = (p.l1 - c1) * sInvD[4 * r + 1];
                            cv[4 * r + 2] = (p.l1 - c2) * sInvD[4 * r + 2]; cv[4 * r + 3] = (p.l1 - c3) * sInvD[4 * r + 3];
                        }
                    } else {
#pragma unroll
                        for (int f = 0; f < 32; ++f) { h[f] = 0.f; cv[f] = 0.f; }
                    }
                    tc_st32(trow, cv);
                }
                // mu' += G' h' on the tensor core
                gs_push_h<false>(L, h, p.k);
                if (trec) trec[2] = clock64();
                const int my_sweeps = tc_block_gs(h, trow, L, p.k, p.max_iter, p.rel_tol, clk, &clk_n);
                if (trec) { trec[3] = clock64(); trec[5] = my_sweeps; }
                // loss partial h' G0 h - 2 h' c:  T = G' h' freshly on the tensor core (the accumulator is free now),
                // G0 h = d o T - (p0 - p1) h - p1 sum(h)   (G = G0 + (p0 - p1) I + p1 11' is what G' was scaled from)
                gs_push_h<true>(L, h, p.k);
                gs_wait(L);
                {
                    float T[32];
                    tc_ld32(trow, T);
                    tc_fence_before();
                    mbar_arrive(CEMPTY(b));          // the accumulator goes back to the MMA warp
                    float s_tile = 0.f;
                    if (valid) {
                        float hs = 0.f;
#pragma unroll
                        for (int f = 0; f < 32; ++f) { h[f] *= sInvD[f]; hs += h[f]; }
                        const float *crow = p.Cscr + (size_t)cell * TC_KP;
                        float s = 0.f;
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            reinterpret_cast<float4 *>(hrow)[r] = make_float4(h[4 * r], h[4 * r + 1], h[4 * r + 2], h[4 * r + 3]);
                            const float4 cq = reinterpret_cast<const float4 *>(crow)[r];
                            const float cc[4] = {cq.x, cq.y, cq.z, cq.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int f = 4 * r + e;
                                const float g0h = fmaf(sD[f], T[f], -fmaf(p.pen_diag, h[f], p.pen_all * hs));
                                s = fmaf(h[f], g0h - 2.f * cc[e], s);
                            }
                        }
                        s_tile = s;
                    }
                    // per tile: one loss / sweep-count atomic per warp (nothing is carried across tiles in registers)
                    const float sw = warp_sum(s_tile);
                    int sc = valid ? my_sweeps : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
                    if (lane == 0) {
                        atomicAdd(&p.scal->loss_part, (double)sw);
                        atomicAdd(&p.scal->sweeps_h, (unsigned long long)sc);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) v = reinterpret_cast<const float4 *>(hrow)[r];
                    h[4 * r] = v.x; h[4 * r + 1] = v.y; h[4 * r + 2] = v.z; h[4 * r + 3] = v.w;
                }
            }
            // ---- split fp16 image of this cell's column of the tile: x = h * 2^h_exp, hi = fp16(x), lo = fp16(x - hi);
            // K-major [factor][cell] sub-tiles of 64 cells, 128 B rows, swizzle-128B by hand
            uint8_t *img = my_img + (size_t)t * TC_H_TILE + (lc >> 6) * TC_H_SUB;
            const int cc = lc & 63;
            float hmax = 0.f;
            bool ovf = false;
#pragma unroll
            for (int f = 0; f < 32; ++f) {
                hmax = fmaxf(hmax, h[f]);
                // row sums of H (penalties, the mkl trace): one shared-memory reduction per factor
                const float fs = warp_sum(h[f]);
                if (lane == (f & 31)) atomicAdd(&s_hsum[f], fs);
                float x = h[f] * hscale;
                if (x > 60000.f) { x = 60000.f; ovf = true; }
                const __half hi = __float2half_rn(x);
                const __half lo = __float2half_rn(x - __half2float(hi));
                const int off = f * 128 + ((((cc >> 3) ^ (f & 7)) << 4) | ((cc & 7) << 1));
                *reinterpret_cast<__half *>(img + off) = hi;
                *reinterpret_cast<__half *>(img + 2 * TC_H_SUB + off) = lo;
            }
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
            mbar_arrive(TDONE(t));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) hmax = fmaxf(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
            if (lane == 0) atomicMax(p.state + TC_ST_HMAX + p.par, __float_as_int(hmax));
            if (__any_sync(0xffffffffu, ovf) && lane == 0) atomicOr(p.state + TC_ST_OVF, 1);
            if (trec) trec[4] = clock64();
            if (lane == 0 && quarter == 0 && t < 8) TC_MARK(16 + t);
        }
    } else if (warp >= 4) {
        // =========================== drain warps ===========================
        // D (128 genes x 32 factors per M-tile) -> descale -> red.add into B
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_DRAIN));
        const int quarter = warp & 3;
        const float bdescale = exp2f(-(float)(p.a_exp + h_exp));
        for (int u = 0; u < n_waves * n_pass; ++u) {
            const int ps = u % n_pass;
            const int mt0 = ps * TC_MT_PER_PASS, mt1 = min(p.n_mtiles, mt0 + TC_MT_PER_PASS);
            mbar_wait(DFULL(u & 1), (u >> 1) & 1);
            tc_fence_after();
            for (int mt = mt0; mt < mt1; ++mt) {
                float d[32];
                tc_ld32(tmem_d(u) + (mt - mt0) * 32 + ((uint32_t)(quarter * 32) << 16), d);
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
            mbar_arrive(DEMPTY(u & 1));
            if (quarter == 0 && lane == 0 && u < 8) TC_MARK(32 + u);
        }
        // H H' of this CTA's tiles (complete: the last unit's commit covers the Gram MMAs issued before it) -> fp64 sums.
        // Lane r of quarter 0 holds row r of hi hi', of quarter 2 row r of lo hi'  (G = hi hi' + lo hi' + (lo hi')').
        if (quarter == 0 || quarter == 2) {
            tc_fence_after();
            float g[32];
            tc_ld32(tmem + TC_GRAM_COL + ((uint32_t)(quarter * 32) << 16), g);      // .sync.aligned: the whole warp
            const float gd = exp2f(-2.f * (float)h_exp);
            double *G = p.G64h + (size_t)p.par * TC_GS;
#pragma unroll
            for (int c = 0; c < TC_KP; ++c) {
                if (c < p.k && lane < p.k) {
                    const double v = (double)(g[c] * gd);
                    atomicAdd(G + lane * TC_KP + c, v);
                    if (quarter == 2) atomicAdd(G + c * TC_KP + lane, v);
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_SERVICE));
        if (warp == 0) {
            // =========================== TMA producer ===========================
            // All lanes walk the schedule and wait for the slot; one elected lane arms the barrier and issues the boxes
            // of the stage back to back (under elect.sync the UTMALDG are straight-line code)
            int stage = 0, phase = 0;
            if (p.mode == 0) {
                for (int t = 0; t < n_tiles; ++t) {
                    const int cell0 = cell_begin + t * TC_TILE;
                    const int nbox = (min(TC_TILE, cell_end - cell0) + TC_GROUP - 1) / TC_GROUP;
                    for (int ch = 0; ch < p.n_chunks; ++ch) {
                        mbar_wait(EMPTY(stage), phase ^ 1);
                        if (elect_one_sync()) {
                            mbar_expect_tx(FULL(stage), nbox * 2 * TC_BOX_BYTES + TC_W_STAGE);
                            for (int b = 0; b < nbox; ++b) {
                                tma_load_2d(sbase + TcSmem::a_hi + stage * TC_A_STAGE + b * TC_BOX_BYTES, &map_ahi, ch * TC_CHUNK,
                                            cell0 + b * TC_GROUP, FULL(stage));
                                tma_load_2d(sbase + TcSmem::a_lo + stage * TC_A_STAGE + b * TC_BOX_BYTES, &map_alo, ch * TC_CHUNK,
                                            cell0 + b * TC_GROUP, FULL(stage));
                            }
                            tma_load_2d(sbase + TcSmem::w + stage * TC_W_STAGE, &map_w, ch * TC_CHUNK, 0, FULL(stage));
                        }
                        __syncwarp();
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            if (lane == 0) TC_MARK(1);
            // phase 3: a stage = 64 cells x 128 genes (two 64-gene boxes side by side: the MN-major M = 128 operand)
            for (int u = 0; u < n_waves * n_pass; ++u) {
                const int ps = u % n_pass, wv = u / n_pass;
                const int mt0 = ps * TC_MT_PER_PASS, mt1 = min(p.n_mtiles, mt0 + TC_MT_PER_PASS);
                for (int t = wave_lo(wv); t < wave_lo(wv + 1); ++t) {
                    const int cell0 = cell_begin + t * TC_TILE;
                    const int nbox = (min(TC_TILE, cell_end - cell0) + TC_GROUP - 1) / TC_GROUP;
                    for (int mt = mt0; mt < mt1; ++mt) {
                        for (int hc = 0; hc < 2; ++hc) {
                            const int nb = min(2, nbox - 2 * hc);
                            if (nb <= 0) break;
                            mbar_wait(EMPTY(stage), phase ^ 1);
                            if (elect_one_sync()) {
                                mbar_expect_tx(FULL(stage), nb * 4 * TC_BOX_BYTES);
                                for (int cg = 0; cg < nb; ++cg)
#pragma unroll
                                    for (int gh = 0; gh < 2; ++gh) {
                                        const int off = stage * TC_A_STAGE + gh * 2 * TC_BOX_BYTES + cg * TC_BOX_BYTES;
                                        tma_load_2d(sbase + TcSmem::a_hi + off, &map_ahi, (2 * mt + gh) * TC_CHUNK,
                                                    cell0 + (2 * hc + cg) * TC_GROUP, FULL(stage));
                                        tma_load_2d(sbase + TcSmem::a_lo + off, &map_alo, (2 * mt + gh) * TC_CHUNK,
                                                    cell0 + (2 * hc + cg) * TC_GROUP, FULL(stage));
                                    }
                            }
                            __syncwarp();
                            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        } else if (warp == 1) {
            // =========================== MMA issuer ===========================
            // all lanes follow the barriers; one elected lane issues the MMAs and the commits
            constexpr uint32_t ID_P1 = idesc_f16(128, 32, 0, 0, 0);
            constexpr uint32_t ID_P3 = idesc_f16(128, 32, 1, 0, 0);
            int stage = 0, phase = 0;
            if (p.mode == 0) {
                for (int t = 0; t < n_tiles; ++t) {
                    const int b = t % TC_CBUF;
                    mbar_wait(CEMPTY(b), ((t / TC_CBUF) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t dC = tmem_c + b * 32;
                    for (int ch = 0; ch < p.n_chunks; ++ch) {
                        mbar_wait(FULL(stage), phase);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint32_t ahi = sbase + TcSmem::a_hi + stage * TC_A_STAGE;
                            const uint32_t alo = sbase + TcSmem::a_lo + stage * TC_A_STAGE;
                            const uint32_t ws = sbase + TcSmem::w + stage * TC_W_STAGE;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint64_t da_hi = desc_sw128(ahi + ks * 32, 16, 1024);
                                const uint64_t dw_hi = desc_sw128(ws + ks * 32, 16, 1024);
                                tc_mma_f16(dC, da_hi, dw_hi, ID_P1, (ch | ks) != 0);
                                tc_mma_f16(dC, da_hi, desc_sw128(ws + 32 * 128 + ks * 32, 16, 1024), ID_P1, 1);
                                tc_mma_f16(dC, desc_sw128(alo + ks * 32, 16, 1024), dw_hi, ID_P1, 1);
                            }
                            tc_commit(EMPTY(stage));
                            if (ch == p.n_chunks - 1) tc_commit(CFULL(b));
                        }
                        __syncwarp();
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            // the W stages are free once every phase-1 MMA has completed: the H tile loader may overwrite them
            if (elect_one_sync()) tc_commit(P1DONE);
            __syncwarp();
            if (lane == 0) TC_MARK(2);
            int use = 0;  // H tile buffer uses
            for (int u = 0; u < n_waves * n_pass; ++u) {
                const int ps = u % n_pass, wv = u / n_pass;
                const int mt0 = ps * TC_MT_PER_PASS, mt1 = min(p.n_mtiles, mt0 + TC_MT_PER_PASS);
                const int r = u & 1;
                const int t_lo = wave_lo(wv), t_hi = wave_lo(wv + 1);
                mbar_wait(DEMPTY(r), ((u >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int t = t_lo; t < t_hi; ++t, ++use) {
                    const int hb = use & 1;
                    const int nbox = (min(TC_TILE, cell_end - (cell_begin + t * TC_TILE)) + TC_GROUP - 1) / TC_GROUP;
                    mbar_wait(HFULL(hb), (use >> 1) & 1);
                    tc_fence_after();
                    const uint32_t hs = sbase + TcSmem::h + hb * TC_H_TILE;
                    if (ps == 0) {
                        // H H' of this tile from its image: A = the whole image as 128 K-major rows [hi c0; hi c1; lo c0; lo c1]
                        // (c0 / c1 = cells 0..63 / 64..127), B = hi of one cell half.  With B = hi c0 rows 0..31 of the
                        // accumulator get hi hi' and rows 64..95 lo hi' of that half; the image shifted by one sub-tile puts
                        // the c1 half on the same rows.  Rows 32..63 and 96..127 mix cells and are never read.
                        if (elect_one_sync()) {
                            const uint32_t dG = tmem + TC_GRAM_COL;
#pragma unroll
                            for (int hc = 0; hc < 2; ++hc)
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    tc_mma_f16(dG, desc_sw128(hs + hc * TC_H_SUB + ks * 32, 16, 1024),
                                               desc_sw128(hs + hc * TC_H_SUB + ks * 32, 16, 1024), ID_P1, (use | hc | ks) != 0);
                        }
                        __syncwarp();
                    }
                    for (int mt = mt0; mt < mt1; ++mt) {
                        const uint32_t dD = tmem_d(u) + (mt - mt0) * 32;
                        for (int hc = 0; hc < 2; ++hc) {
                            if (nbox - 2 * hc <= 0) break;
                            mbar_wait(FULL(stage), phase);
                            tc_fence_after();
                            if (elect_one_sync()) {
                                const uint32_t ahi = sbase + TcSmem::a_hi + stage * TC_A_STAGE;
                                const uint32_t alo = sbase + TcSmem::a_lo + stage * TC_A_STAGE;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    // A operand, MN-major: 64 genes contiguous (128 B), 8 cells per swizzle atom (SBO 1024),
                                    // the second 64-gene atom 8 KB further (LBO)
                                    const uint64_t da_hi = desc_sw128(ahi + ks * 2048, 2 * TC_BOX_BYTES, 1024);
                                    const uint64_t da_lo = desc_sw128(alo + ks * 2048, 2 * TC_BOX_BYTES, 1024);
                                    const uint32_t hoff = hc * TC_H_SUB + ks * 32;
                                    const uint64_t dh_hi = desc_sw128(hs + hoff, 16, 1024);
                                    const uint64_t dh_lo = desc_sw128(hs + 2 * TC_H_SUB + hoff, 16, 1024);
                                    tc_mma_f16(dD, da_hi, dh_hi, ID_P3, ((t - t_lo) | hc | ks) != 0);
                                    tc_mma_f16(dD, da_hi, dh_lo, ID_P3, 1);
                                    tc_mma_f16(dD, da_lo, dh_hi, ID_P3, 1);
                                }
                                tc_commit(EMPTY(stage));
                            }
                            __syncwarp();
                            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                    if (elect_one_sync()) {
                        tc_commit(HEMPTY(hb));
                        if (t == t_hi - 1) tc_commit(DFULL(r));
                    }
                    __syncwarp();
                    if (lane == 0 && ps == 0 && t < 8) TC_MARK(24 + t);
                }
                if (lane == 0 && u == n_waves * n_pass - 1) TC_MARK(3);
            }
        } else if (warp == 2) {
            // =========================== H tile loader ===========================
            mbar_wait(P1DONE, 0);
            int use = 0;
            for (int u = 0; u < n_waves * n_pass; ++u) {
                const int wv = u / n_pass;
                for (int t = wave_lo(wv); t < wave_lo(wv + 1); ++t, ++use) {
                    const int hb = use & 1;
                    mbar_wait(TDONE(t), 0);
                    mbar_wait(HEMPTY(hb), ((use >> 1) & 1) ^ 1);
                    // one 16 KB bulk copy (async proxy); the solver threads fenced their generic-proxy writes of the image
                    // with fence.proxy.async before arriving on TDONE
                    if (elect_one_sync()) {
                        asm volatile("fence.proxy.async;" ::: "memory");
                        mbar_expect_tx(HFULL(hb), TC_H_TILE);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                         sbase + TcSmem::h + hb * TC_H_TILE),
                                     "l"(my_img + (size_t)t * TC_H_TILE), "r"(TC_H_TILE), "r"(HFULL(hb))
                                     : "memory");
                    }
                    __syncwarp();
                }
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
    if (threadIdx.x < p.k) atomicAdd(p.G64h + (size_t)p.par * TC_GS + TC_KP * TC_KP + threadIdx.x, (double)s_hsum[threadIdx.x]);
    if (tc_last_cta(p.state + TC_ST_CNT_P, s_flag)) {
        // what the next W-side kernel needs: the GramPack of H H' (single GPU; with several ranks the host all-reduces the
        // sums first), zeroed accumulators for the next iteration
        if (p.finalize) tc_finalize_gram(p.G64h + (size_t)p.par * TC_GS, p.pack_next, p.k, p.pen_next0, p.pen_next1, s_hsum);
        double *Gz = p.G64h + (size_t)(p.par ^ 1) * TC_GS;
        for (int t = threadIdx.x; t < TC_GS; t += blockDim.x) Gz[t] = 0.0;
        if (threadIdx.x == 0) p.state[TC_ST_WMAX] = 0;       // the next W-side kernel takes the maximum of its W afresh
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
struct TcMatrix {
    bool ready = false;
    int n = 0, m = 0, n_pad = 0, a_exp = 0, sm_count = 148;
    __half *Ahi = nullptr, *Alo = nullptr, *Wst = nullptr;
    int *state = nullptr;      // TC_ST_* ints of the running fit
    double *G64w = nullptr, *G64h = nullptr;   // [2][TC_GS] fp64 sums, double buffered by iteration parity
    float *w_partial = nullptr;                // [ceil(n/128)][TC_GS] per-CTA partial sums of the W-side kernel
    float *packW = nullptr, *packH = nullptr;  // GramPacks: W'W (+beta) for the H update, H H' (+alpha) for the W update
    const char *env_timeline = nullptr, *env_clocks = nullptr;   // SWNE_TC_TIMELINE / SWNE_TC_CLOCKS, read once per matrix
    unsigned long long *dbg = nullptr;   // timeline buffer when SWNE_TC_TIMELINE is set
    unsigned long long *dbg_w = nullptr; // the same for the W-side kernel
    long long *clk = nullptr;            // solver clock ring when SWNE_TC_CLOCKS is set
    uint8_t *Himg = nullptr;   // split 16-bit H tile images, [grid][tiles_per_cta][TC_H_TILE]
    int grid = 0, tiles_per_cta = 0;
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
    return k >= 1 && k <= TC_KP && n >= 64 && m >= 128 && (long long)m <= 148ll * TC_MAX_TILES * TC_TILE / 2;
}

static inline void tc_release(TcMatrix &tc)
{
    if (tc.Ahi) cudaFree(tc.Ahi);
    if (tc.Alo) cudaFree(tc.Alo);
    if (tc.Wst) cudaFree(tc.Wst);
    if (tc.state) cudaFree(tc.state);
    if (tc.G64w) cudaFree(tc.G64w);
    if (tc.G64h) cudaFree(tc.G64h);
    if (tc.w_partial) cudaFree(tc.w_partial);
    if (tc.packW) cudaFree(tc.packW);
    if (tc.packH) cudaFree(tc.packH);
    if (tc.Himg) cudaFree(tc.Himg);
    if (tc.dbg) cudaFree(tc.dbg);
    if (tc.dbg_w) cudaFree(tc.dbg_w);
    if (tc.clk) cudaFree(tc.clk);
    tc = TcMatrix();
}

// builds the split copy of A and the tensor maps (once per resident matrix)
static inline void tc_prepare(TcMatrix &tc, const float *dA, size_t ldA, int n, int m, double a_max, int sm_count, cudaStream_t st)
{
    if (tc.ready && tc.src == dA && tc.n == n && tc.m == m) return;
    // a new matrix of the same shape (repeated RunNMF calls, FindNumFactors on resampled data) reuses the
    // buffers: cudaFree/cudaMalloc of the 4-byte-per-entry split copy cost more than the split itself
    const bool reuse = tc.Ahi && tc.n == n && tc.m == m && tc.sm_count == sm_count;
    if (!reuse) {
        tc_release(tc);
        tc.n = n; tc.m = m; tc.sm_count = sm_count;
        tc.n_pad = round_up(n, 64);
        const size_t cnt = (size_t)m * tc.n_pad;
        CUDA_CHECK(cudaMalloc(&tc.Ahi, cnt * sizeof(__half)));
        CUDA_CHECK(cudaMalloc(&tc.Alo, cnt * sizeof(__half)));
        CUDA_CHECK(cudaMalloc(&tc.Wst, (size_t)64 * tc.n_pad * sizeof(__half)));
        CUDA_CHECK(cudaMalloc(&tc.state, TC_ST_INTS * sizeof(int)));
        CUDA_CHECK(cudaMalloc(&tc.G64w, 2 * TC_GS * sizeof(double)));
        CUDA_CHECK(cudaMalloc(&tc.G64h, 2 * TC_GS * sizeof(double)));
        CUDA_CHECK(cudaMalloc(&tc.w_partial, (size_t)((n + 127) / 128) * TC_GS * sizeof(float)));
        CUDA_CHECK(cudaMalloc(&tc.packW, gram_pack_floats(TC_KP) * sizeof(float)));
        CUDA_CHECK(cudaMalloc(&tc.packH, gram_pack_floats(TC_KP) * sizeof(float)));
        const int groups = (m + TC_GROUP - 1) / TC_GROUP;
        tc.grid = std::min(sm_count, groups);
        const int max_groups = (groups + tc.grid - 1) / tc.grid;
        tc.tiles_per_cta = (max_groups * TC_GROUP + TC_TILE - 1) / TC_TILE;
        CUDA_CHECK(cudaMalloc(&tc.Himg, (size_t)tc.grid * tc.tiles_per_cta * TC_H_TILE));
    }
    tc.src = dA;
    CUDA_CHECK(cudaMemsetAsync(tc.state, 0, TC_ST_INTS * sizeof(int), st));
    tc.env_timeline = getenv("SWNE_TC_TIMELINE");
    tc.env_clocks = getenv("SWNE_TC_CLOCKS");
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

// ---- per-fit device state of the fused engine (owned by TcMatrix, reset by tc_fit_begin)
//   state ints (TC_ST_*), G64w / G64h: [2][TC_GS] fp64 sums (double buffered by iteration parity), packW / packH: the two
//   GramPacks (W'W + beta penalties for the H update, H H' + alpha penalties for the W update)

// Start of a fit: W0 / H0 are on the device.  Zeroes the state and records max(H0) as the first scale reference of the
// H images (the pre-pass needs no W operand).  Once per fit.
static inline void tc_fit_begin(TcMatrix &tc, const float *Wt, const float *H, cudaStream_t st)
{
    CUDA_CHECK(cudaMemsetAsync(tc.state, 0, TC_ST_INTS * sizeof(int), st));
    CUDA_CHECK(cudaMemsetAsync(tc.G64w, 0, 2 * TC_GS * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(tc.G64h, 0, 2 * TC_GS * sizeof(double), st));
    // the pre-pass (parity 1) reads max(H0) from half 0 and writes half 1; iteration 0 (parity 0) then reads half 1
    (void)Wt;
    tc_absmax_kernel<<<148, 256, 0, st>>>(H, (size_t)tc.m * TC_KP, tc.state + TC_ST_HMAX + 0);
    CUDA_CHECK(cudaGetLastError());
}

// one pass.  mode 0: H update (phase 1 + solve) followed by B = A H' and H H' of the new H; mode 1: B = A H', H H' only.
// par: iteration parity (the pre-pass before iteration 0 runs with par = 1).  finalize: the last CTA builds packH from
// H H' and (alpha0, alpha1) -- single GPU; with several ranks the caller all-reduces G64h[par] and finalises itself.
// launches: 1 kernel (+ 1 constant-bank copy of the triangular coefficients in mode 0)
static inline void tc_pass(TcMatrix &tc, int mode, int par, bool finalize, float *H, float *Cscr, float l1, float pen_diag, float pen_all,
                           double alpha0, double alpha1, int k, int max_iter, float rel_tol, float *B, FitScalars *scal, cudaStream_t st)
{
    TcPassParams p;
    p.n = tc.n; p.m = tc.m; p.k = k;
    p.n_mtiles = (tc.n + 127) / 128;
    p.n_chunks = 2 * p.n_mtiles;
    p.mode = mode; p.par = par; p.finalize = finalize ? 1 : 0;
    p.pack = tc.packW; p.pack_next = tc.packH; p.pen_next0 = alpha0; p.pen_next1 = alpha1;
    p.l1 = l1; p.pen_diag = pen_diag; p.pen_all = pen_all; p.max_iter = max_iter; p.rel_tol = rel_tol;
    p.H = H; p.Cscr = Cscr; p.B = B; p.scal = scal;
    p.state = tc.state; p.G64h = tc.G64h;
    p.a_exp = tc.a_exp;
    p.groups_total = (tc.m + TC_GROUP - 1) / TC_GROUP;
    p.Himg = tc.Himg;
    p.tiles_per_cta = tc.tiles_per_cta;
    const char *tl = tc.env_timeline;
    if (tl && !tc.dbg) CUDA_CHECK(cudaMalloc(&tc.dbg, sizeof(unsigned long long) * 64 * tc.grid));
    p.dbg = tl ? tc.dbg : nullptr;
    if (p.dbg) CUDA_CHECK(cudaMemsetAsync(tc.dbg, 0, sizeof(unsigned long long) * 64 * tc.grid, st));
    // SWNE_TC_CLOCKS=<file> (builds with -DTC_SOLVER_CLOCKS only): clock64 stamps of every block update of every solver group
    const char *ck = tc.env_clocks;
    const size_t clk_count = (size_t)tc.grid * TC_TGROUPS * TC_CLK_RECS * 8;
    if (ck && !tc.clk) CUDA_CHECK(cudaMalloc(&tc.clk, clk_count * sizeof(long long)));
    p.clk = (ck && mode == 0) ? tc.clk : nullptr;
    if (p.clk) CUDA_CHECK(cudaMemsetAsync(tc.clk, 0, clk_count * sizeof(long long), st));
    const int grid = tc.grid;
    if (tc.tiles_per_cta > TC_MAX_TILES) {
        set_error("tcgen05 engine: too many cells per SM");
        throw CudaError{cudaErrorInvalidValue, __FILE__, __LINE__};
    }
    if (mode == 0)
        CUDA_CHECK(cudaMemcpyToSymbolAsync(c_gram, tc.packW, sizeof(float) * TC_KP * TC_KP, 0, cudaMemcpyDeviceToDevice, st));
    tc_pass_kernel<<<grid, 768, TcSmem::total + 1024, st>>>(tc.map_ahi, tc.map_alo, tc.map_w, p);
    CUDA_CHECK(cudaGetLastError());
    if (p.clk) {
        std::vector<long long> hbuf(clk_count);
        CUDA_CHECK(cudaStreamSynchronize(st));
        CUDA_CHECK(cudaMemcpy(hbuf.data(), tc.clk, clk_count * sizeof(long long), cudaMemcpyDeviceToHost));
        FILE *f = fopen(ck, "wb");
        if (f) {
            const int hdr[4] = {tc.grid, TC_TGROUPS, TC_CLK_RECS, 8};
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(hbuf.data(), sizeof(long long), clk_count, f);
            fclose(f);
        }
    }
    if (p.dbg && mode == 0) {
        // debug only: dump the per-CTA timeline of this launch (ns since the earliest CTA start)
        std::vector<unsigned long long> hbuf(64 * (size_t)grid);
        CUDA_CHECK(cudaStreamSynchronize(st));
        CUDA_CHECK(cudaMemcpy(hbuf.data(), tc.dbg, hbuf.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (int c = 0; c < grid; ++c) if (hbuf[c * 64]) t0 = std::min(t0, hbuf[c * 64]);
        FILE *f = fopen(tl, "w");
        if (f) {
            for (int c = 0; c < grid; ++c) {
                fprintf(f, "cta %d", c);
                for (int s = 0; s < 40; ++s) fprintf(f, " %lld", hbuf[c * 64 + s] ? (long long)(hbuf[c * 64 + s] - t0) : -1ll);
                for (int s = 40; s < 46; ++s) fprintf(f, " %lld", (long long)hbuf[c * 64 + s]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
}
static inline int tc_pass_launches(int mode) { return mode == 0 ? 2 : 1; }

// the W side of iteration `par`: packH (H H' + alpha) -> solve -> W, its split image, W'W -> packW (beta0, beta1).
// B holds A H' (all-reduced) and is zeroed as it is consumed.  launches: 1 kernel + 1 constant-bank copy
static inline void tc_wside(TcMatrix &tc, int par, float *Wt, float *B, float l1, double beta0, double beta1, int k, int max_iter,
                            float rel_tol, FitScalars *scal, cudaStream_t st)
{
    TcWsideParams p;
    p.X = Wt; p.C = B; p.pack = tc.packH; p.pack_next = tc.packW; p.pen_next0 = beta0; p.pen_next1 = beta1;
    p.l1 = l1; p.k = k; p.cols = tc.n; p.n_pad = tc.n_pad; p.max_iter = max_iter; p.rel_tol = rel_tol;
    p.Wst = tc.Wst; p.state = tc.state; p.par = par; p.G64w = tc.G64w; p.partial = tc.w_partial; p.scal = scal;
    const int grid = (tc.n + 127) / 128;
    const char *tl = tc.env_timeline;
    if (tl && !tc.dbg_w) CUDA_CHECK(cudaMalloc(&tc.dbg_w, sizeof(unsigned long long) * 16 * grid));
    p.dbg = tl ? tc.dbg_w : nullptr;
    if (p.dbg) CUDA_CHECK(cudaMemsetAsync(tc.dbg_w, 0, sizeof(unsigned long long) * 16 * grid, st));
    CUDA_CHECK(cudaMemcpyToSymbolAsync(c_gram, tc.packH, sizeof(float) * TC_KP * TC_KP, 0, cudaMemcpyDeviceToDevice, st));
    tc_wside_kernel<<<grid, TC_WSIDE_THREADS, 0, st>>>(p);
    tc_wfinish_kernel<<<(tc.n_pad + 127) / 128, 1024, 0, st>>>(p, grid);
    CUDA_CHECK(cudaGetLastError());
    if (p.dbg) {
        // debug only: per-CTA marks of this launch (ns since the earliest CTA start), next to the pass timeline
        std::vector<unsigned long long> hbuf(16 * (size_t)grid);
        CUDA_CHECK(cudaStreamSynchronize(st));
        CUDA_CHECK(cudaMemcpy(hbuf.data(), tc.dbg_w, hbuf.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (int c = 0; c < grid; ++c) if (hbuf[c * 16]) t0 = std::min(t0, hbuf[c * 16]);
        FILE *f = fopen((std::string(tl) + ".wside").c_str(), "w");
        if (f) {
            for (int c = 0; c < grid; ++c) {
                fprintf(f, "cta %d", c);
                for (int q = 0; q < 12; ++q) fprintf(f, " %lld", hbuf[c * 16 + q] ? (long long)(hbuf[c * 16 + q] - t0) : -1ll);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
}
static inline int tc_wside_launches() { return 3; }

}  // namespace swne
