// kernels_kp.cu -- instantiates the KP-templated kernels for one KP (-DKP_VALUE=8|16|...|64).
#include <stdlib.h>

#include "kernels_simt.cuh"
#include "kp_ops.h"

#ifndef KP_VALUE
#error "compile with -DKP_VALUE=<multiple of 8 up to 64>"
#endif

namespace swne {
namespace {
constexpr int KP = KP_VALUE;

void gram_l(const float *X, int rows, double *G64, double *colsum64, cudaStream_t st)
{
    const int grid = rows < 296 * 64 ? (rows + 63) / 64 : 296;
    gram_kernel<KP><<<grid, 256, 0, st>>>(X, rows, G64, colsum64);
}
void contract_cells_l(const float *A, size_t ldA, int n, int m, const float *Wt, float *C, cudaStream_t st)
{
    contract_cells_kernel<KP><<<(m + 63) / 64, 128, 0, st>>>(A, ldA, n, m, Wt, C);
}
void contract_genes_l(const float *A, size_t ldA, int n, int m, const float *H, float *B, int cpc, int chunks, cudaStream_t st)
{
    dim3 grid((n + 63) / 64, chunks);
    contract_genes_kernel<KP><<<grid, 128, 0, st>>>(A, ldA, n, m, H, B, cpc);
}
void solve_l(float *X, const float *C, const float *pack, float l1, int k, int cols, int max_iter, float rel_tol,
             unsigned long long *sweeps, double *loss_part, cudaStream_t st)
{
    cudaMemcpyToSymbolAsync(c_gram, pack, sizeof(float) * 2 * KP * KP, 0, cudaMemcpyDeviceToDevice, st);
    // few columns (the W side): one warp per block spreads the latency-bound solves over all SMs
    const int bs = cols <= 148 * 64 ? 32 : 128;
    scd_ls_solve_kernel<KP><<<(cols + bs - 1) / bs, bs, 0, st>>>(X, C, pack, l1, k, cols, max_iter, rel_tol, sweeps, loss_part);
}
// blocks of scd_kl_kernel<KP, NT> that share an SM when a column of up to `cap` nonzeros is staged: (KP + 2) floats each
static int kl_per_sm(int cap, int NT)
{
    const size_t smem = (size_t)cap * (KP + 2) * sizeof(float);
    int per_sm = (int)((220 * 1024) / (smem + 2048));
    if (per_sm > 2048 / NT) per_sm = 2048 / NT;
    return per_sm < 1 ? 1 : per_sm;
}
template <int NT>
static void kl_launch_nt(const int *ptr, const int *idx, const float *val, float *ajt, float *X, const float *Y,
                         const float *sumY, float b0, float b1, float b2, int k, int cols, int max_iter, float rel_tol,
                         unsigned long long *sweeps, double *kl_part, double *sq_part, int err, int max_nnz, const int *order,
                         const int *sorted_cnt, cudaStream_t st, int first0 = 0)
{
    // the launch covers columns order[first0 .. first0 + cols) (first0 > 0: the warp kernel took the shorter ones)
    if (sorted_cnt) sorted_cnt += first0;
    // shared-memory staging of one column: (KP + 2) floats per nonzero; columns longer than cap run on global arrays
    // set on every launch: the attribute is per device, and one process may drive several GPUs (one handle each)
    const int smem_budget = 200 * 1024;
    cudaFuncSetAttribute(scd_kl_kernel<KP, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_budget);
    const int cap_max = (smem_budget / (int)((KP + 2) * sizeof(float))) & ~3;
    int cap = (max_nnz + 3) & ~3;
    if (cap < 4) cap = 4;
    if (cap > cap_max) cap = cap_max;
    // debug knob (read at launch, so a test can flip it in-process): an upper bound on the staged column length -- shorter
    // staging buffers let more blocks share an SM, columns above the bound run on the global arrays
    const char *cap_e = getenv("SWNE_KL_CAP");
    const int cap_env = cap_e ? atoi(cap_e) : 0;
    if (cap_env >= 4 && cap > (cap_env & ~3)) cap = cap_env & ~3;
    // The sweeps are latency-bound chains, so what counts is how many columns an SM works on at once, and that is set by the
    // staging buffer of the LONGEST column of the launch (cfg2's gene side: longest 1804 of mean 684 nonzeros -> one block
    // per SM).  With the columns sorted by length the launch is cut in two: the long columns with the big buffer, the rest
    // with a buffer sized for P blocks per SM; P and the cut minimise the number of block rounds (one extra round is charged
    // for the second launch's tail).  SWNE_KL_SPLIT=0 keeps the single launch.
    int n_long = 0, cap_short = cap;
    const char *split_e = getenv("SWNE_KL_SPLIT");
    if (order && sorted_cnt && cols > 148 && !(split_e && atoi(split_e) == 0) && cap_env < 4) {
        const int p_long = kl_per_sm(cap, NT);
        double best = (double)cols / (148.0 * p_long);
        for (int P = p_long + 1; P <= 2048 / NT; ++P) {
            int cap_p = (int)(((220 * 1024) / P - 2048) / ((KP + 2) * sizeof(float))) & ~3;
            if (cap_p < 4) break;
            // columns longer than cap_p: sorted_cnt is descending
            int lo = 0, hi = cols;
            while (lo < hi) { const int mid = (lo + hi) / 2; if (sorted_cnt[mid] > cap_p) lo = mid + 1; else hi = mid; }
            const double rounds = (double)lo / (148.0 * p_long) + (double)(cols - lo) / (148.0 * P) + (lo ? 1.0 : 0.0);
            if (rounds < best) { best = rounds; n_long = lo; cap_short = lo < cols ? ((sorted_cnt[lo] + 3) & ~3) : 4; }
        }
        if (cap_short < 4) cap_short = 4;
    }
    for (int part = 0; part < 2; ++part) {
        const int first = first0 + (part == 0 ? 0 : n_long);
        const int count = part == 0 ? (cap_short == cap ? cols : n_long) : cols - n_long;
        if (part == 1 && cap_short == cap) break;
        if (count <= 0) continue;
        const int cap_l = part == 0 ? cap : cap_short;
        const size_t smem = (size_t)cap_l * (KP + 2) * sizeof(float);
        const int per_sm = kl_per_sm(cap_l, NT);
        const int grid = count < 148 * per_sm ? count : 148 * per_sm;
        scd_kl_kernel<KP, NT><<<grid, NT, smem, st>>>(ptr, idx, val, ajt, X, Y, sumY, b0, b1, b2, k, count, max_iter, rel_tol,
                                                      sweeps, kl_part, sq_part, err, cap_l, order, first);
    }
}
// ---- scd_kl_warp_kernel: one warp per column, 12 bytes of shared memory per nonzero
constexpr int KLW_CAP_MAX = 2048;     // longer columns stay with the block kernel
static int klw_blocks_per_sm(int cap)
{
    const size_t smem = (size_t)(KLW_THREADS / 32) * (3 * (size_t)cap + KP) * sizeof(float);
    int b = (int)((216 * 1024) / (smem + 1024));
    return b < 1 ? 1 : (b > 8 ? 8 : b);           // 8 blocks of 4 warps: beyond 32 warps per SM the kernel is issue-bound anyway
}
static void kl_warp_launch(const int *ptr, const int *idx, const float *val, float *X, const float *Y, const float *Yt, size_t ld,
                           const float *sumY, float b0, float b1, float b2, int k, int max_iter, float rel_tol,
                           unsigned long long *sweeps, double *kl_part, double *sq_part, int err, const int *order,
                           const int *sorted_cnt, int first0, int cols, cudaStream_t st)
{
    if (cols <= 0) return;
    constexpr int WPB = KLW_THREADS / 32;
    cudaFuncSetAttribute(scd_kl_warp_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int *cnt = sorted_cnt + first0;          // descending
    int cap_all = (cnt[0] + 3) & ~3;
    if (cap_all < 4) cap_all = 4;
    // two classes as in kl_launch_nt: the longer columns with the buffer of the longest, the rest with a buffer that lets
    // more blocks share an SM; the cut minimises the number of block rounds
    const int p_all = klw_blocks_per_sm(cap_all);
    int n_long = 0, cap_short = cap_all;
    double best = (double)cols / (148.0 * p_all);
    for (int P = p_all + 1; P <= 8 && cols > 148 * WPB; ++P) {
        const long per_warp_floats = ((216 * 1024) / P - 1024) / WPB / (long)sizeof(float);
        int cap_p = (int)((per_warp_floats - KP) / 3) & ~3;
        if (cap_p < 4) break;
        int lo = 0, hi = cols;
        while (lo < hi) { const int mid = (lo + hi) / 2; if (cnt[mid] > cap_p) lo = mid + 1; else hi = mid; }
        const double rounds = (double)lo / (148.0 * p_all) + (double)(cols - lo) / (148.0 * P);
        if (lo < cols && rounds < best) { best = rounds; n_long = lo; cap_short = (cnt[lo] + 3) & ~3; }
    }
    if (cap_short < 4) cap_short = 4;
    for (int part = 0; part < 2; ++part) {
        const bool split = cap_short != cap_all;
        if (part == 1 && !split) break;
        const int first = first0 + (part == 0 ? 0 : n_long);
        const int count = part == 0 ? (split ? n_long : cols) : cols - n_long;
        if (count <= 0) continue;
        const int cap_l = part == 0 ? cap_all : cap_short;
        const size_t smem = (size_t)WPB * (3 * (size_t)cap_l + KP) * sizeof(float);
        const int bps = klw_blocks_per_sm(cap_l);
        int grid = (count + WPB - 1) / WPB;
        if (grid > 148 * bps) grid = 148 * bps;
        scd_kl_warp_kernel<KP><<<grid, KLW_THREADS, smem, st>>>(ptr, idx, val, X, Y, Yt, ld, sumY, b0, b1, b2, k, count, max_iter,
                                                                rel_tol, sweeps, kl_part, sq_part, err, cap_l, order, first);
    }
}
void kl_l(const int *ptr, const int *idx, const float *val, float *ajt, float *X, const float *Y, const float *sumY, float b0,
          float b1, float b2, int k, int cols, int max_iter, float rel_tol, unsigned long long *sweeps, double *kl_part,
          double *sq_part, int err, int max_nnz, int avg_nnz, const int *order, const int *sorted_cnt, const float *Yt, size_t ld,
          cudaStream_t st)
{
    // Columns of up to KLW_CAP_MAX nonzeros: one warp per column (scd_kl_warp_kernel); the longer ones: one block per column.
    // A warp walks its column alone (14-20 nonzeros per lane and coordinate at cfg2), so the warp kernel needs many more columns
    // than warp slots to win: with 3000 columns (cfg2) it has 20 warps per SM running one column each and takes 440 us per
    // iteration against the block kernel's 234; it is used from 2 x 32 warps per SM upward.  SWNE_KL_WARP=0 / 1 (debug
    // knob, read at launch) forces the block / warp kernel.
    const char *warp_e = getenv("SWNE_KL_WARP");
    const bool warp_want = warp_e ? atoi(warp_e) != 0 : cols >= 148 * 64;
    const bool warp_on = Yt && order && sorted_cnt && cols > 0 && warp_want;
    int n_big = 0;      // columns [0, n_big) of the sorted order are longer than KLW_CAP_MAX
    if (warp_on) {
        int lo = 0, hi = cols;
        while (lo < hi) { const int mid = (lo + hi) / 2; if (sorted_cnt[mid] > KLW_CAP_MAX) lo = mid + 1; else hi = mid; }
        n_big = lo;
        kl_warp_launch(ptr, idx, val, X, Y, Yt, ld, sumY, b0, b1, b2, k, max_iter, rel_tol, sweeps, kl_part, sq_part, err, order,
                       sorted_cnt, n_big, cols - n_big, st);
        if (n_big == 0) return;
        cols = n_big;
    }
    // Threads per column.  Every warp of the block pays the per-coordinate reduction (two butterflies, a barrier, the Newton
    // step) whatever the column length, so a thread should own several nonzeros: at cfg2 (456 / 684 nonzeros per column,
    // k = 16) 512 threads per column executed ~150 warp instructions per 32 nonzero-coordinates and ran issue-bound (65 %
    // issue slots, ncu); measured per iteration, both sides: NT 512 / 256 / 128 / 64 = 307 / 239 / 232 / 308 us.
    int nt = avg_nnz > 2048 ? 512 : (avg_nnz > 1024 ? 256 : (avg_nnz > 96 ? 128 : 64));
    if (warp_on) nt = 512;                        // only the long columns are left
    const char *nt_e = getenv("SWNE_KL_NT");     // debug knob, read at launch: threads per column
    if (nt_e) nt = atoi(nt_e);
#define KL_NT_CASE(NTV)                                                                                                         \
    case NTV:                                                                                                                   \
        kl_launch_nt<NTV>(ptr, idx, val, ajt, X, Y, sumY, b0, b1, b2, k, cols, max_iter, rel_tol, sweeps, kl_part, sq_part, err, \
                          max_nnz, order, sorted_cnt, st);                                                                      \
        break;
    switch (nt) {
        KL_NT_CASE(64)
        KL_NT_CASE(128)
        KL_NT_CASE(256)
        default:
        KL_NT_CASE(512)
    }
#undef KL_NT_CASE
}
void kl_shard_init_l(const int *ptr, const int *idx, float *ajt, const float *X, const float *Y, int k, int cols, float *dprev,
                     float *sumX, int *active, cudaStream_t st)
{
    kl_shard_init_kernel<KP><<<cols < 148 * 8 ? cols : 148 * 8, 256, 0, st>>>(ptr, idx, ajt, X, Y, k, cols, dprev, sumX, active);
}
void kl_shard_transpose_l(const float *Y, int rows, float *Yt, size_t ld, cudaStream_t st)
{
    if (rows > 0) kl_shard_transpose_kernel<KP><<<(rows + 31) / 32, 256, 0, st>>>(Y, rows, Yt, ld);
}
void kl_shard_accum_l(const int *ptr, const int *idx, const float *val, float *ajt, const float *Yt, size_t ld, const float *dprev,
                      const int *active, int kk, int kprev, int cols, float *part, cudaStream_t st)
{
    kl_shard_accum_kernel<KP><<<cols < 148 * 8 ? cols : 148 * 8, 256, 0, st>>>(ptr, idx, val, ajt, Yt, ld, dprev, active, kk, kprev,
                                                                               cols, reinterpret_cast<float2 *>(part));
}
void kl_shard_update_l(const float *part, float *X, const float *sumY, float *sumX, float *dprev, int *active, float b0, float b1,
                       float b2, int kk, int cols, float rel_tol, int flags, cudaStream_t st)
{
    kl_shard_update_kernel<KP><<<(cols + 127) / 128, 128, 0, st>>>(reinterpret_cast<const float2 *>(part), X, sumY, sumX, dprev,
                                                                   active, b0, b1, b2, kk, cols, rel_tol, flags);
}
void kl_shard_sweep_end_l(int *active, int cols, unsigned long long *sweeps, unsigned *n_active, cudaStream_t st)
{
    kl_shard_sweep_end_kernel<KP><<<(cols + 127) / 128, 128, 0, st>>>(active, cols, sweeps, n_active);
}
void kl_trace_l(const float *A, size_t ldA, int n, int m, const float *H, const float *Wt, int k, double *kl_part,
                cudaStream_t st)
{
    // grid.y is capped at 65535 blocks of KL_CELLS cells: longer matrices go in slabs
    const int slab = 65535 * KL_CELLS;
    for (int c0 = 0; c0 < m; c0 += slab) {
        const int mc = (m - c0 < slab) ? m - c0 : slab;
        dim3 grid((n + 255) / 256, (mc + KL_CELLS - 1) / KL_CELLS);
        kl_trace_kernel<KP><<<grid, 256, 0, st>>>(A + (size_t)c0 * ldA, ldA, n, mc, H + (size_t)c0 * KP, Wt, k, kl_part);
    }
}
void recon_error_l(const float *A, size_t ldA, int n, int m, const float *Wt, const float *H, int k, double *out2, int grid,
                   cudaStream_t st)
{
    recon_error_kernel<KP><<<grid, 256, 0, st>>>(A, ldA, n, m, Wt, H, k, out2);
}
void right_multiply_l(float *X, size_t rows, const float *T, int l, cudaStream_t st)
{
    right_multiply_kernel<KP><<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(X, rows, T, l);
}
void ica_step_l(const float *Y, int rows, int nc, const float *R, double *out, int grid, cudaStream_t st)
{
    ica_step_kernel<KP><<<grid, 256, 0, st>>>(Y, rows, nc, R, out);
}
}  // namespace

#define KP_OPS_NAME2(v) kp_ops_table_##v
#define KP_OPS_NAME(v) KP_OPS_NAME2(v)
extern const KpOps KP_OPS_NAME(KP_VALUE) = {gram_l, contract_cells_l, contract_genes_l, solve_l, kl_l, kl_shard_init_l, kl_shard_transpose_l, kl_shard_accum_l,
                                            kl_shard_update_l, kl_shard_sweep_end_l, kl_trace_l,
                                            recon_error_l, right_multiply_l, ica_step_l};
}  // namespace swne
