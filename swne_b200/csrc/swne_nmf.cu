// swne_nmf.cu -- C-ABI (include/swne_nmf.h) and host orchestration of the B200-native NMF hot path.
//
// Host-side restatement of NNLM's c_nnmf outer loop (reference call sites
// /root/reference/R/nmf.R:117,129,167-168) driving the CUDA kernels in kernels_simt.cuh (fp32 CUDA-core
// engine) and pass_tc.cuh (fused tcgen05 engine).  No CPU fallback: every numerical step runs on the GPU.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>  // types only; the functions are resolved with dlopen so a 1-GPU user needs no NCCL
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <chrono>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_misc.cuh"
#include "kp_ops.h"
#include "linalg_small.h"
#include "pass_tc.cuh"
#include "contract_tc.cuh"
#include "seed.cuh"
#include "peer_xchg.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace swne {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
struct ArgError { int code; };
#define FAIL(code, ...)               \
    do {                              \
        swne::set_error(__VA_ARGS__); \
        throw swne::ArgError{code};   \
    } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static void load_nccl()
{
    if (g_nccl.lib) return;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *nm : names) {
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) FAIL(SWNE_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LOADSYM(field, name)                                             \
    g_nccl.field = (decltype(g_nccl.field))dlsym(lib, name);             \
    if (!g_nccl.field) FAIL(SWNE_ERR_NCCL, "NCCL symbol %s missing", name)
    LOADSYM(GetUniqueId, "ncclGetUniqueId");
    LOADSYM(CommInitRank, "ncclCommInitRank");
    LOADSYM(CommDestroy, "ncclCommDestroy");
    LOADSYM(AllReduce, "ncclAllReduce");
    LOADSYM(AllGather, "ncclAllGather");
    LOADSYM(GroupStart, "ncclGroupStart");
    LOADSYM(GroupEnd, "ncclGroupEnd");
    LOADSYM(GetErrorString, "ncclGetErrorString");
#undef LOADSYM
    g_nccl.lib = lib;
}
#define NCCL_CHECK(expr)                                                                              \
    do {                                                                                              \
        ncclResult_t _r = (expr);                                                                     \
        if (_r != ncclSuccess) FAIL(SWNE_ERR_NCCL, "NCCL error at %s:%d: %s", __FILE__, __LINE__,     \
                                    g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");         \
    } while (0)

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    void reserve(size_t count)
    {
        if (count <= cap) return;
        release();
        CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
        cap = count;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
    }
};

struct SparseView {
    bool ready = false;
    DevBuf<int> cptr, cidx, gptr, gidx;
    DevBuf<float> cval, gval, ajt;
    DevBuf<int> corder, gorder;                 // columns of each view by descending length (the KL solver's launch classes)
    std::vector<int> ccnt_sorted, gcnt_sorted;  // their lengths, host side
    size_t nnz = 0;
    int cmax = 0, gmax = 0;   // longest column of each view (sizes the shared-memory staging of the KL solver)
};

}  // namespace swne

struct swne_nmf_handle_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    // resident matrix (this rank's cells)
    int n = 0, m = 0;
    size_t ldA = 0;
    float *dA = nullptr;
    bool own_A = false;
    swne::DevBuf<float> A_store;
    swne::MatStats stats{};          // local
    swne::SparseView sp;
    swne::TcMatrix tc;               // split-fp16 copy + TMA descriptors for the tcgen05 engine
    swne::TcxState tcx;              // operands of the un-fused tcgen05 contractions (32 < k <= 64)
    // communicator
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // per fit workspaces (grown on demand)
    swne::DevBuf<float> Wt, H, C, B, G, pack;   // pack: GramPack (kernels_simt.cuh); G: small scratch of the rSVD
    swne::DevBuf<double> G64w, G64h, stage;
    swne::DevBuf<swne::MatStats> dstats;
    swne::DevBuf<swne::FitScalars> dscal;
    swne::DevBuf<double> red;        // packed all-reduce buffer of doubles
    swne::DevBuf<unsigned> flags;    // device-side validation flags of the init
    void *pinned = nullptr;          // host mirror of scalars + Grams
    size_t pinned_bytes = 0;
    swne_nmf_progress_fn progress = nullptr;
    void *progress_user = nullptr;
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // slab pipeline of the wide engine: [0] C = A W, [1] solves (low priority), [2] A H'
    cudaEvent_t ev_ready = nullptr, ev_done[3] = {nullptr, nullptr, nullptr};
    swne::PeerXchg xchg;                               // NVLink peer-memory exchange of the fused engine's per-iteration sums
    std::vector<cudaEvent_t> ev_slab;           // [2 s] contraction of slab s done, [2 s + 1] its solve done
    swne::DevBuf<float> kls_f;       // sharded KL gene side: Newton partials (2 n), pending delta (n), row sums of W (n)
    swne::DevBuf<float> kls_yt;      // ... the fixed factor in factor-major order [KP][rows padded to 32]
    swne::DevBuf<int> kls_i;         // ... active / again flags (2 n) and the count of genes still sweeping
    int kl_flags = 0;                // KL_L2_LATE when the fit asks for the alternative reading of NNLM's KL step
    int last_KPv = 0;                // row stride of Wt / H as the last fit left them on the device
    int last_k = 0;                  // k of the fit whose H is still in h->H (0: none -- another call has used the buffer since)
};

namespace swne {

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        CUDA_CHECK(cudaSetDevice(dev));
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

extern const KpOps kp_ops_table_8, kp_ops_table_16, kp_ops_table_24, kp_ops_table_32, kp_ops_table_40, kp_ops_table_48,
    kp_ops_table_56, kp_ops_table_64;
const KpOps *kp_ops(int KP)
{
    switch (KP) {
        case 8: return &kp_ops_table_8;
        case 16: return &kp_ops_table_16;
        case 24: return &kp_ops_table_24;
        case 32: return &kp_ops_table_32;
        case 40: return &kp_ops_table_40;
        case 48: return &kp_ops_table_48;
        case 56: return &kp_ops_table_56;
        case 64: return &kp_ops_table_64;
        default: return nullptr;
    }
}

static int kp_of(int k) { return std::max(8, round_up(k, 8)); }

static const KpOps &ops_of(int KPv)
{
    const KpOps *o = kp_ops(KPv);
    if (!o) FAIL(SWNE_ERR_BAD_ARG, "k out of range");
    return *o;
}

// ------------------------------------------------------------------------------------------------
// matrix ingest
// ------------------------------------------------------------------------------------------------
static void check_stats(swne_nmf_handle_s *h)
{
    CUDA_CHECK(cudaMemcpyAsync(&h->stats, h->dstats.p, sizeof(MatStats), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (h->stats.bad & 4u) { h->n = h->m = 0; FAIL(SWNE_ERR_BAD_ARG, "CSC row index out of range"); }
    if (h->stats.bad & 1u) { h->n = h->m = 0; FAIL(SWNE_ERR_BAD_ARG, "The input matrix contains negative elements !"); }
    if (h->stats.bad & 2u) { h->n = h->m = 0; FAIL(SWNE_ERR_BAD_ARG, "The input matrix contains non-finite elements (NA/NaN/Inf); masks and missing values are not supported"); }
}

static void begin_matrix(swne_nmf_handle_s *h, int n, int m, bool alloc)
{
    if (n < 1 || m < 1) FAIL(SWNE_ERR_BAD_ARG, "matrix must have at least one row and one column");
    h->sp.ready = false;
    h->tc.ready = false;
    h->last_k = 0;
    h->n = n; h->m = m;
    h->ldA = round_up_sz((size_t)n, 4);
    if (alloc) {
        h->A_store.reserve((size_t)m * h->ldA);
        h->dA = h->A_store.p;
        h->own_A = true;
    }
    h->dstats.reserve(1);
    CUDA_CHECK(cudaMemsetAsync(h->dstats.p, 0, sizeof(MatStats), h->stream));
}

template <typename T>
static void set_dense(swne_nmf_handle_s *h, const T *A, int n, int m)
{
    if (!A) FAIL(SWNE_ERR_BAD_ARG, "A is NULL");
    begin_matrix(h, n, m, true);
    if (std::is_same<T, float>::value && h->ldA == (size_t)n) {
        // fp32 with no row padding: the host buffer is already the device layout -- one copy straight into place,
        // then a read-only pass for the statistics and the validation
        CUDA_CHECK(cudaMemcpyAsync(h->dA, A, (size_t)m * n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        const int grid = (int)std::min<size_t>((size_t)m, (size_t)h->sm_count * 8);
        ingest_dense_kernel<float><<<grid, 256, 0, h->stream>>>(h->dA, (size_t)n, nullptr, h->ldA, n, m, h->dstats.p);
        CUDA_CHECK(cudaGetLastError());
        check_stats(h);
        return;
    }
    const size_t rows_per_chunk = std::max<size_t>(1, ((size_t)256 << 20) / ((size_t)n * sizeof(T)));
    DevBuf<T> stagebuf;
    stagebuf.reserve(std::min<size_t>(rows_per_chunk, m) * (size_t)n);
    for (size_t r0 = 0; r0 < (size_t)m; r0 += rows_per_chunk) {
        const size_t nr = std::min(rows_per_chunk, (size_t)m - r0);
        CUDA_CHECK(cudaMemcpyAsync(stagebuf.p, A + r0 * n, nr * n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        const int grid = (int)std::min<size_t>(nr, (size_t)h->sm_count * 8);
        ingest_dense_kernel<T><<<grid, 256, 0, h->stream>>>(stagebuf.p, (size_t)n, h->dA + r0 * h->ldA, h->ldA, n, (int)nr,
                                                            h->dstats.p);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    stagebuf.release();
    check_stats(h);
}

template <typename T>
static void set_csc(swne_nmf_handle_s *h, const int *p, const int *idx, const T *x, int n, int m, const double *cell_scale = nullptr,
                    const double *gene_scale = nullptr, int transform = 0)
{
    if (transform < 0 || transform > 2) FAIL(SWNE_ERR_BAD_ARG, "transform must be 0 (none), 1 (log) or 2 (ft)");
    if (!p || (!idx && p[m] > 0) || (!x && p[m] > 0)) FAIL(SWNE_ERR_BAD_ARG, "CSC slot pointer is NULL");
    if (p[0] != 0) FAIL(SWNE_ERR_BAD_ARG, "CSC column pointer must start at 0 (dgCMatrix @p)");
    for (int j = 0; j < m; ++j)
        if (p[j + 1] < p[j]) FAIL(SWNE_ERR_BAD_ARG, "CSC column pointer is not non-decreasing");
    begin_matrix(h, n, m, true);
    CUDA_CHECK(cudaMemsetAsync(h->dA, 0, (size_t)m * h->ldA * sizeof(float), h->stream));
    DevBuf<int> dp, di;
    DevBuf<T> dx;
    DevBuf<double> dcs, dgs;
    dp.reserve((size_t)m + 1);
    CUDA_CHECK(cudaMemcpyAsync(dp.p, p, ((size_t)m + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (cell_scale) {
        dcs.reserve((size_t)m);
        CUDA_CHECK(cudaMemcpyAsync(dcs.p, cell_scale, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    if (gene_scale) {
        dgs.reserve((size_t)n);
        CUDA_CHECK(cudaMemcpyAsync(dgs.p, gene_scale, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    // upload the nonzeros in column chunks of <= 64M entries
    const long long chunk_nnz = 64ll << 20;
    int c0 = 0;
    while (c0 < m) {
        int c1 = c0 + 1;
        while (c1 < m && (long long)p[c1 + 1] - p[c0] <= chunk_nnz) ++c1;
        const size_t cnt = (size_t)(p[c1] - p[c0]);
        if (cnt) {
            di.reserve(cnt); dx.reserve(cnt);
            CUDA_CHECK(cudaMemcpyAsync(di.p, idx + p[c0], cnt * sizeof(int), cudaMemcpyHostToDevice, h->stream));
            CUDA_CHECK(cudaMemcpyAsync(dx.p, x + p[c0], cnt * sizeof(T), cudaMemcpyHostToDevice, h->stream));
            const int cols = c1 - c0;
            const int grid = std::min((cols + 7) / 8, h->sm_count * 8);
            ingest_csc_kernel<T><<<grid, 256, 0, h->stream>>>(dp.p, di.p, dx.p, h->dA + (size_t)c0 * h->ldA, h->ldA, n, c0,
                                                              cols, p[c0], h->dstats.p, cell_scale ? dcs.p : nullptr,
                                                              gene_scale ? dgs.p : nullptr, transform);
            CUDA_CHECK(cudaGetLastError());
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
        }
        c0 = c1;
    }
    // the implicit zeros contribute (0+tiny) log(0+tiny) each to the KL constant
    check_stats(h);
    const double zeros = (double)n * (double)m - (double)h->stats.nnz;
    h->stats.klc += zeros * (TINY_NUM_D * log(TINY_NUM_D));
    dp.release(); di.release(); dx.release(); dcs.release(); dgs.release();
}

static void ensure_sparse(swne_nmf_handle_s *h)
{
    if (h->sp.ready) return;
    // the sparse views index their nonzeros with 32-bit offsets
    if (h->stats.nnz > 0x7fffffffull)
        FAIL(SWNE_ERR_UNSUPPORTED, "loss = mkl needs the sparse views of A, which hold at most 2^31 - 1 nonzeros (this matrix has %llu)",
             (unsigned long long)h->stats.nnz);
    SparseView &sp = h->sp;
    const int n = h->n, m = h->m;
    cudaStream_t st = h->stream;
    DevBuf<int> cnt;
    // by cell
    cnt.reserve((size_t)m);
    sp.cptr.reserve((size_t)m + 1);
    count_nz_by_cell_kernel<<<std::min((m + 7) / 8, h->sm_count * 16), 256, 0, st>>>(h->dA, h->ldA, n, m, cnt.p);
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(cnt.p, sp.cptr.p, m);
    int nnz = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nnz, sp.cptr.p + m, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    sp.nnz = (size_t)nnz;
    sp.cidx.reserve((size_t)nnz); sp.cval.reserve((size_t)nnz);
    sp.gidx.reserve((size_t)nnz); sp.gval.reserve((size_t)nnz); sp.ajt.reserve((size_t)nnz);
    fill_by_cell_kernel<<<std::min((m + 7) / 8, h->sm_count * 16), 256, 0, st>>>(h->dA, h->ldA, n, m, sp.cptr.p, sp.cidx.p,
                                                                                sp.cval.p);
    // by gene
    const int chunk = 1024, nchunks = (m + chunk - 1) / chunk;
    DevBuf<int> cnt2, tot;
    cnt2.reserve((size_t)nchunks * n); tot.reserve((size_t)n);
    sp.gptr.reserve((size_t)n + 1);
    dim3 grid((n + 127) / 128, nchunks);
    count_nz_by_gene_kernel<<<grid, 128, 0, st>>>(h->dA, h->ldA, n, m, chunk, cnt2.p);
    chunk_offsets_kernel<<<(n + 127) / 128, 128, 0, st>>>(cnt2.p, nchunks, n, tot.p);
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(tot.p, sp.gptr.p, n);
    fill_by_gene_kernel<<<grid, 128, 0, st>>>(h->dA, h->ldA, n, m, chunk, cnt2.p, sp.gptr.p, sp.gidx.p, sp.gval.p);
    CUDA_CHECK(cudaGetLastError());
    {
        std::vector<int> hc((size_t)m), hg((size_t)n);
        CUDA_CHECK(cudaMemcpyAsync(hc.data(), cnt.p, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(hg.data(), tot.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        sp.cmax = m ? *std::max_element(hc.begin(), hc.end()) : 0;
        sp.gmax = n ? *std::max_element(hg.begin(), hg.end()) : 0;
        // the columns of each view by descending length: the KL solver sizes its staging buffers per launch, and runs the
        // few long columns apart from the many short ones (kernels_kp.cu, kl_launch_nt)
        auto sorted_order = [&](const std::vector<int> &cntv, DevBuf<int> &order, std::vector<int> &sorted) {
            std::vector<int> ord(cntv.size());
            for (size_t q = 0; q < ord.size(); ++q) ord[q] = (int)q;
            std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return cntv[a] > cntv[b]; });
            sorted.resize(ord.size());
            for (size_t q = 0; q < ord.size(); ++q) sorted[q] = cntv[ord[q]];
            order.reserve(ord.size());
            if (!ord.empty())
                CUDA_CHECK(cudaMemcpyAsync(order.p, ord.data(), sizeof(int) * ord.size(), cudaMemcpyHostToDevice, st));
            CUDA_CHECK(cudaStreamSynchronize(st));   // ord is a local
        };
        sorted_order(hc, sp.corder, sp.ccnt_sorted);
        sorted_order(hg, sp.gorder, sp.gcnt_sorted);
    }
    cnt.release(); cnt2.release(); tot.release();
    sp.ready = true;
}

// ------------------------------------------------------------------------------------------------
// profiling helper: event pairs per kernel family
// ------------------------------------------------------------------------------------------------
struct Prof {
    bool on = false;
    cudaStream_t st = nullptr;
    struct Rec { int slot; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    long long launches[SWNE_PROF_SLOTS] = {0};
    long long total_launches = 0;
    std::vector<cudaEvent_t> pool;
    size_t pool_next = 0;
    cudaEvent_t get()
    {
        if (pool_next == pool.size()) {
            for (int i = 0; i < 256; ++i) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        }
        return pool[pool_next++];
    }
    void begin(int slot)
    {
        if (!on) return;
        Rec r{slot, get(), get()};
        cudaEventRecord(r.a, st);
        recs.push_back(r);
    }
    void end(int slot, int n_launch)
    {
        launches[slot] += n_launch;
        total_launches += n_launch;
        if (!on) return;
        cudaEventRecord(recs.back().b, st);
    }
    void finish(swne_nmf_result_t *res)
    {
        for (int s = 0; s < SWNE_PROF_SLOTS; ++s) { res->prof_ms[s] = 0; res->prof_launches[s] = launches[s]; }
        res->gpu_launches = total_launches;
        for (auto &r : recs) {
            float ms = 0.f;
            cudaEventSynchronize(r.b);
            cudaEventElapsedTime(&ms, r.a, r.b);
            res->prof_ms[r.slot] += ms;
        }
        recs.clear();
        for (auto e : pool) cudaEventDestroy(e);
        pool.clear(); pool_next = 0;
    }
    ~Prof()
    {
        for (auto e : pool) cudaEventDestroy(e);   // exit through an exception: finish() never ran
    }
};

// ------------------------------------------------------------------------------------------------
// building blocks
// ------------------------------------------------------------------------------------------------
// zero `bytes` (multiple of 4, pointer 16-byte aligned) with a kernel
static void dev_zero(void *ptr, size_t bytes, cudaStream_t st)
{
    const size_t n16 = bytes / 16;
    const int ntail = (int)((bytes % 16) / 4);
    const int grid = (int)std::min<size_t>(std::max<size_t>(1, (n16 + 255) / 256), 296);
    zero_fill_kernel<<<grid, 256, 0, st>>>((uint4 *)ptr, n16, (uint32_t *)((char *)ptr + n16 * 16), ntail);
    CUDA_CHECK(cudaGetLastError());
}
static void launch_gram(swne_nmf_handle_s *h, int KPv, const float *X, int rows, double *G64 /* KP*KP + KP */)
{
    dev_zero(G64, sizeof(double) * (size_t)(KPv * KPv + KPv), h->stream);
    ops_of(KPv).gram(X, rows, G64, G64 + KPv * KPv, h->stream);
    CUDA_CHECK(cudaGetLastError());
}
static void launch_finalize_gram(swne_nmf_handle_s *h, int k, int KPv, const double *G64, float *pack, double p0, double p1)
{
    finalize_gram_kernel<<<1, 256, 0, h->stream>>>(G64, pack, k, KPv, p0, p1);
    CUDA_CHECK(cudaGetLastError());
}
static float *pack_colsum(float *pack, int KPv) { return pack + 2 * KPv * KPv + 2 * KPv; }
static void launch_contract_cells(swne_nmf_handle_s *h, int KPv, const float *Wt, float *C)
{
    ops_of(KPv).contract_cells(h->dA, h->ldA, h->n, h->m, Wt, C, h->stream);
    CUDA_CHECK(cudaGetLastError());
}
static void launch_contract_genes(swne_nmf_handle_s *h, int KPv, const float *H, float *B)
{
    CUDA_CHECK(cudaMemsetAsync(B, 0, sizeof(float) * (size_t)h->n * KPv, h->stream));
    const int gx = (h->n + 63) / 64;
    // enough chunks to fill the machine ~4x over, each at least 256 cells
    int chunks = std::max(1, std::min((h->m + 255) / 256, (h->sm_count * 8 + gx - 1) / gx));
    int cpc = round_up((h->m + chunks - 1) / chunks, 32);
    chunks = (h->m + cpc - 1) / cpc;
    ops_of(KPv).contract_genes(h->dA, h->ldA, h->n, h->m, H, B, cpc, chunks, h->stream);
    CUDA_CHECK(cudaGetLastError());
}
static void launch_solve(swne_nmf_handle_s *h, int k, int KPv, float *X, const float *C, const float *pack, float l1, int cols,
                         int max_iter, float rel_tol, unsigned long long *sweeps, double *loss_part)
{
    ops_of(KPv).solve(X, C, pack, l1, k, cols, max_iter, rel_tol, sweeps, loss_part, h->stream);
    CUDA_CHECK(cudaGetLastError());
}
static void launch_kl(swne_nmf_handle_s *h, int k, int KPv, const int *ptr, const int *idx, const float *val, float *X,
                      const float *Y, const float *sumY, const double *pen, int cols, int max_iter, float rel_tol,
                      unsigned long long *sweeps, FitScalars *sc, int err)
{
    const bool by_cell = ptr == h->sp.cptr.p;
    const int max_nnz = by_cell ? h->sp.cmax : h->sp.gmax;
    const int avg_nnz = (int)(h->sp.nnz / (size_t)std::max(1, by_cell ? h->m : h->n));
    // the fixed factor in factor-major order for the warp-per-column kernel (its coordinate steps read one column of it); the
    // kernel is chosen by column count (kernels_kp.cu, kl_l: the same rule), so the copy is only made when it will be read
    const int rows = by_cell ? h->n : h->m;
    const size_t ld = (size_t)round_up(std::max(rows, 1), 32);
    const char *warp_e = getenv("SWNE_KL_WARP");
    const bool warp_want = warp_e ? atoi(warp_e) != 0 : cols >= 148 * 64;
    if (warp_want) {
        h->kls_yt.reserve((size_t)KPv * ld);
        ops_of(KPv).kl_shard_transpose(Y, rows, h->kls_yt.p, ld, h->stream);
    }
    ops_of(KPv).kl(ptr, idx, val, h->sp.ajt.p, X, Y, sumY, (float)pen[0], (float)pen[1], (float)pen[2], k, cols, max_iter,
                   rel_tol, sweeps, &sc->kl_part, &sc->sq_part, err | h->kl_flags, max_nnz, avg_nnz,
                   by_cell ? h->sp.corder.p : h->sp.gorder.p, by_cell ? h->sp.ccnt_sorted.data() : h->sp.gcnt_sorted.data(),
                   warp_want ? h->kls_yt.p : nullptr, ld, h->stream);
    CUDA_CHECK(cudaGetLastError());
}

static void allreduce(swne_nmf_handle_s *h, float *f, size_t nf, double *d, size_t nd, double *d2 = nullptr, size_t nd2 = 0);
// Gene side of the KL update when the cells are sharded over ranks: a gene's Newton sums run over all cells, so every
// coordinate step is   partial sums on the rank's nonzeros -> all-reduce of 2 n floats -> the step, replicated
// (k all-reduces per sweep; NNLM runs one sweep per outer iteration for mkl).  Returns the number of kernel launches.
static int launch_kl_sharded(swne_nmf_handle_s *h, int k, int KPv, float *X, const float *Y, int rows, const float *sumY,
                             const double *pen, int cols, int max_iter, float rel_tol, unsigned long long *sweeps)
{
    const KpOps &o = ops_of(KPv);
    cudaStream_t st = h->stream;
    SparseView &sp = h->sp;
    h->kls_f.reserve(4 * (size_t)cols);
    h->kls_i.reserve(2 * (size_t)cols + 4);
    float *part = h->kls_f.p, *dprev = part + 2 * (size_t)cols, *sumX = dprev + cols;
    int *active = h->kls_i.p;
    unsigned *n_active = reinterpret_cast<unsigned *>(active + 2 * (size_t)cols);
    const size_t ld = (size_t)round_up(std::max(rows, 1), 32);
    h->kls_yt.reserve((size_t)KPv * ld);
    int launches = 2;
    o.kl_shard_transpose(Y, rows, h->kls_yt.p, ld, st);
    o.kl_shard_init(sp.gptr.p, sp.gidx.p, sp.ajt.p, X, Y, k, cols, dprev, sumX, active, st);
    for (int it = 0; it < max_iter; ++it) {
        for (int kk = 0; kk < k; ++kk) {
            o.kl_shard_accum(sp.gptr.p, sp.gidx.p, sp.gval.p, sp.ajt.p, h->kls_yt.p, ld, dprev, active, kk, (kk + k - 1) % k, cols, part,
                             st);
            allreduce(h, part, 2 * (size_t)cols, nullptr, 0, nullptr, 0);
            o.kl_shard_update(part, X, sumY, sumX, dprev, active, (float)pen[0], (float)pen[1], (float)pen[2], kk, cols, rel_tol,
                              h->kl_flags, st);
        }
        CUDA_CHECK(cudaMemsetAsync(n_active, 0, sizeof(unsigned), st));
        o.kl_shard_sweep_end(active, cols, sweeps, n_active, st);
        CUDA_CHECK(cudaGetLastError());
        launches += 2 * k + 1;
        if (it + 1 < max_iter) {
            // W is replicated, so every rank reads the same count and leaves the loop at the same sweep
            unsigned left = 0;
            CUDA_CHECK(cudaMemcpyAsync(&left, n_active, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
            if (!left) break;
        }
    }
    return launches;
}

static bool mse_loss_of(const swne_nmf_params_t &p) { return p.loss == SWNE_LOSS_MSE; }
// mean gene row of the by-gene view longer than scd_kl_kernel's shared-memory staging (200 KB / (KP + 2) floats per nonzero)
static bool kl_rows_too_long(const swne_nmf_handle_s *h, int KPv)
{
    const size_t cap_max = (200 * 1024) / ((size_t)(KPv + 2) * sizeof(float));
    return h->sp.ready && h->n > 0 && h->sp.nnz / (size_t)h->n > cap_max;
}

static void upload_factor_cellmajor(swne_nmf_handle_s *h, const double *src, float *dst, int k, int KPv, size_t cols,
                                    unsigned *bad = nullptr, unsigned bit = 0)
{
    h->stage.reserve(cols * (size_t)k);
    CUDA_CHECK(cudaMemcpyAsync(h->stage.p, src, cols * k * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const size_t tot = cols * (size_t)KPv;
    factor_in_cellmajor_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->stage.p, dst, k, KPv, cols, bad, bit);
    CUDA_CHECK(cudaGetLastError());
}
static void download_factor_cellmajor(swne_nmf_handle_s *h, const float *src, double *dst, int k, int KPv, size_t cols)
{
    h->stage.reserve(cols * (size_t)k);
    const size_t tot = cols * (size_t)k;
    factor_out_cellmajor_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(src, h->stage.p, k, KPv, cols);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(dst, h->stage.p, tot * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
}
static void upload_factor_colmajor(swne_nmf_handle_s *h, const double *src, float *dst, int k, int KPv, int n,
                                   unsigned *bad = nullptr, unsigned bit = 0)
{
    h->stage.reserve((size_t)n * k);
    CUDA_CHECK(cudaMemcpyAsync(h->stage.p, src, (size_t)n * k * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const size_t tot = (size_t)n * KPv;
    factor_in_colmajor_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->stage.p, dst, k, KPv, n, bad, bit);
    CUDA_CHECK(cudaGetLastError());
}
static void download_factor_colmajor(swne_nmf_handle_s *h, const float *src, double *dst, int k, int KPv, int n)
{
    h->stage.reserve((size_t)n * k);
    const size_t tot = (size_t)n * k;
    factor_out_colmajor_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(src, h->stage.p, k, KPv, n);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(dst, h->stage.p, tot * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

// one NCCL group (one launch) for up to three buffers of the per-iteration exchange
static void allreduce(swne_nmf_handle_s *h, float *f, size_t nf, double *d, size_t nd, double *d2, size_t nd2)
{
    if (h->nranks <= 1) return;
    NCCL_CHECK(g_nccl.GroupStart());
    if (nf) NCCL_CHECK(g_nccl.AllReduce(f, f, nf, ncclFloat32, ncclSum, h->comm, h->stream));
    if (nd) NCCL_CHECK(g_nccl.AllReduce(d, d, nd, ncclFloat64, ncclSum, h->comm, h->stream));
    if (nd2) NCCL_CHECK(g_nccl.AllReduce(d2, d2, nd2, ncclFloat64, ncclSum, h->comm, h->stream));
    NCCL_CHECK(g_nccl.GroupEnd());
}

// ---- peer-memory exchange (peer_xchg.cuh): set-up is a collective over the handle's communicator
static void xchg_release(swne_nmf_handle_s *h, bool collective)
{
    PeerXchg &x = h->xchg;
    if (!x.region && x.opened.empty()) { x.ready = false; return; }
    cudaStreamSynchronize(h->stream);
    for (void *q : x.opened) cudaIpcCloseMemHandle(q);
    x.opened.clear();
    if (collective && h->comm && h->nranks > 1) {
        // nobody may free a region a peer still has mapped: barrier between closing and freeing
        h->red.reserve(8);
        cudaMemsetAsync(h->red.p, 0, sizeof(double), h->stream);
        g_nccl.AllReduce(h->red.p, h->red.p, 1, ncclFloat64, ncclSum, h->comm, h->stream);
        cudaStreamSynchronize(h->stream);
    }
    if (x.region) cudaFree(x.region);
    if (x.peers_dev) cudaFree(x.peers_dev);
    x.region = nullptr; x.peers_dev = nullptr; x.ready = false; x.payload = 0; x.epoch = 0;
}

// returns true when every rank mapped every other rank's region (all ranks return the same answer)
static bool xchg_setup(swne_nmf_handle_s *h, size_t payload)
{
    PeerXchg &x = h->xchg;
    if (x.tried && x.payload == payload) return x.ready;
    xchg_release(h, true);
    x.tried = true; x.payload = payload; x.nranks = h->nranks; x.rank = h->rank;
    cudaStream_t st = h->stream;
    const int R = h->nranks;
    x.flags_off = 2 * (size_t)R * payload;
    x.total = x.flags_off + 2 * (size_t)R * 8 + 64;
    int ok = 1;
    if (getenv("SWNE_PEER_XCHG") && atoi(getenv("SWNE_PEER_XCHG")) == 0) ok = 0;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok) {
        if (cudaMalloc(&x.region, x.total) != cudaSuccess) { cudaGetLastError(); x.region = nullptr; ok = 0; }
        else {
            CUDA_CHECK(cudaMemsetAsync(x.region, 0, x.total, st));
            if (cudaIpcGetMemHandle(&mine, x.region) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        }
    }
    // swap the handles (64 bytes each) and the "I could allocate" bits
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    DevBuf<unsigned char> hb;
    hb.reserve(rec * R);
    std::vector<unsigned char> host(rec * R, 0);
    memcpy(host.data() + rec * h->rank, &mine, sizeof(mine));
    host[rec * h->rank + sizeof(mine)] = (unsigned char)ok;
    CUDA_CHECK(cudaMemcpyAsync(hb.p + rec * h->rank, host.data() + rec * h->rank, rec, cudaMemcpyHostToDevice, st));
    NCCL_CHECK(g_nccl.AllGather(hb.p + rec * h->rank, hb.p, rec, ncclUint8, h->comm, st));
    CUDA_CHECK(cudaMemcpyAsync(host.data(), hb.p, rec * R, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    for (int r = 0; r < R; ++r) ok &= host[rec * r + sizeof(mine)];
    std::vector<uint8_t *> peers(R, nullptr);
    if (ok) {
        for (int r = 0; r < R && ok; ++r) {
            if (r == h->rank) { peers[r] = x.region; continue; }
            cudaIpcMemHandle_t mh;
            memcpy(&mh, host.data() + rec * r, sizeof(mh));
            void *q = nullptr;
            if (cudaIpcOpenMemHandle(&q, mh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            x.opened.push_back(q);
            peers[r] = (uint8_t *)q;
        }
    }
    // every rank must have mapped every region (this all-reduce is also the barrier behind which the zeroed flags are safe to use)
    double v = ok ? 1.0 : 0.0;
    h->red.reserve(8);
    CUDA_CHECK(cudaMemcpyAsync(h->red.p, &v, sizeof(double), cudaMemcpyHostToDevice, st));
    NCCL_CHECK(g_nccl.AllReduce(h->red.p, h->red.p, 1, ncclFloat64, ncclMin, h->comm, st));
    CUDA_CHECK(cudaMemcpyAsync(&v, h->red.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (v < 0.5) {
        const size_t keep = x.payload;
        xchg_release(h, true);
        x.payload = keep; x.tried = true; x.ready = false;
        return false;
    }
    CUDA_CHECK(cudaMalloc(&x.peers_dev, sizeof(uint8_t *) * R));
    CUDA_CHECK(cudaMemcpyAsync(x.peers_dev, peers.data(), sizeof(uint8_t *) * R, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    x.epoch = 0;
    x.ready = true;
    return true;
}

static void validate_init(const double *x, size_t len, const char *name)
{
    for (size_t i = 0; i < len; ++i)
        if (!(x[i] >= 0.0) || isinf(x[i])) FAIL(SWNE_ERR_BAD_ARG, "init %s has negative or non-finite entries", name);
}

// the five timing events of one fit; destroyed on every exit path, including exceptions from CUDA_CHECK / FAIL
struct FitEvents {
    cudaEvent_t e[5];
    FitEvents() { for (auto &x : e) cudaEventCreate(&x); }
    ~FitEvents() { for (auto &x : e) cudaEventDestroy(x); }
    FitEvents(const FitEvents &) = delete;
    FitEvents &operator=(const FitEvents &) = delete;
};

// host mirror read back at every error block
struct HostTrace {
    FitScalars sc;
    int tc_overflow;   // the tcgen05 engine's "a scaled H entry left the fp16 range" flag
    int xchg_err;      // peer-memory exchange: a rank's flag did not arrive in time
    double G64w[SWNE_MAX_K * SWNE_MAX_K + SWNE_MAX_K];  // W'W and column sums of W
    double G64h[SWNE_MAX_K * SWNE_MAX_K + SWNE_MAX_K];  // H H' and row sums of H (global)
};

// ------------------------------------------------------------------------------------------------
// the fit (c_nnmf)
// ------------------------------------------------------------------------------------------------
// device_init(KPv, mean_A, m_global): when given, the start is produced on the device -- it must fill h->Wt.p [n][KPv]
// and h->H.p [m][KPv] on the handle's stream (RunNMF's seeds: seed.cuh); W / H are then outputs only.
typedef std::function<void(int, double, double)> DeviceInitFn;
static int fit_impl(swne_nmf_handle_s *h, const swne_nmf_params_t *pp, double *W, double *H, double *mse_tr, double *mkl_tr,
                    double *terr_tr, double *epo_tr, int trace_len, swne_nmf_result_t *res, const DeviceInitFn *device_init = nullptr)
{
    h->last_k = 0;     // set again when this fit has left its H on the device
    if (!h->dA || h->n == 0) FAIL(SWNE_ERR_NO_MATRIX, "no matrix resident: call swne_nmf_set_matrix_* first");
    if (!pp || ((!W || !H) && !device_init)) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    swne_nmf_params_t p = *pp;
    const int k = p.k, n = h->n, m = h->m;
    if (k < 1 || k > SWNE_MAX_K) FAIL(SWNE_ERR_BAD_ARG, "k must be in 1..%d", SWNE_MAX_K);
    if (p.loss != SWNE_LOSS_MSE && p.loss != SWNE_LOSS_MKL) FAIL(SWNE_ERR_BAD_ARG, "loss must be mse or mkl");
    if (p.max_iter < 1) FAIL(SWNE_ERR_BAD_ARG, "max_iter must be >= 1");
    for (int q = 0; q < 3; ++q)
        if (!(p.alpha[q] >= 0.0) || !(p.beta[q] >= 0.0)) FAIL(SWNE_ERR_BAD_ARG, "penalties must be >= 0");
    if (p.inner_max_iter <= 0) p.inner_max_iter = (p.loss == SWNE_LOSS_MSE) ? 50 : 1;
    if (p.inner_rel_tol <= 0.0) p.inner_rel_tol = 1e-9;
    if (p.trace <= 0) p.trace = std::max(1, 100 / p.inner_max_iter);
    const int need_len = (p.max_iter + p.trace - 1) / p.trace + 1;
    if ((mse_tr || mkl_tr || terr_tr || epo_tr) && trace_len < need_len)
        FAIL(SWNE_ERR_BAD_ARG, "trace arrays need ceil(max_iter/trace)+1 = %d entries", need_len);
    h->kl_flags = (p.nnlm_variant & SWNE_VARIANT_KL_L2_LATE) ? KL_L2_LATE : 0;

    cudaStream_t st = h->stream;
    memset(res, 0, sizeof(*res));

    int engine = p.engine;
    if (p.loss == SWNE_LOSS_MKL) engine = SWNE_ENGINE_SIMT;
    // k <= 32: the fused pass (pass_tc.cuh); 32 < k <= 64: tcgen05 contractions + CUDA-core solves (contract_tc.cuh)
    // (the un-fused kernels also take k <= 32 when the fused pass cannot: more than ~1.2 M cells on one GPU;
    // SWNE_TC_FORCE_WIDE selects them for any k, for tests and comparisons)
    bool fused_ok = k <= TC_KP && tc_supported(h->n, h->m, k) && !getenv("SWNE_TC_FORCE_WIDE");
    bool wide_ok = tcx_supported(h->n, h->m, k);

    // ---- global constants (all-reduced once).  The engine is part of the exchange: the engines differ in their
    // per-iteration collectives, so every rank must take the same one even when a shard boundary leaves one rank
    // with a shape another engine would refuse.
    double gl[7] = {h->stats.sum, h->stats.sumsq, h->stats.klc, (double)h->stats.nnz, (double)m,
                    fused_ok ? 0.0 : 1.0, wide_ok ? 0.0 : 1.0};
    if (h->nranks > 1) {
        h->red.reserve(8);
        CUDA_CHECK(cudaMemcpyAsync(h->red.p, gl, sizeof(gl), cudaMemcpyHostToDevice, st));
        allreduce(h, nullptr, 0, h->red.p, 7);
        CUDA_CHECK(cudaMemcpyAsync(gl, h->red.p, sizeof(gl), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        fused_ok = gl[5] == 0.0;
        wide_ok = gl[6] == 0.0;
    }
    const double sumsqA = gl[1], klcA = gl[2], m_glob = gl[4];
    const double N = (double)n * m_glob;

    const bool tc_ok = fused_ok || wide_ok;
    if (engine == SWNE_ENGINE_AUTO) engine = tc_ok ? SWNE_ENGINE_TC : SWNE_ENGINE_SIMT;
    if (engine == SWNE_ENGINE_TC && !tc_ok)
        FAIL(SWNE_ERR_UNSUPPORTED, "tcgen05 engine does not support this shape (n=%d, k=%d)", n, k);
    res->engine_used = engine;
    const bool tcx = engine == SWNE_ENGINE_TC && !fused_ok;   // wide variant: the loop below runs its fp32-engine branch
    const bool fused = engine == SWNE_ENGINE_TC && !tcx;
    const int KPv = fused ? TC_KP : kp_of(k);
    const int GS = KPv * KPv + KPv;  // Gram + column sums

    // ---- workspaces
    h->Wt.reserve((size_t)n * KPv); h->H.reserve((size_t)m * KPv);
    h->C.reserve((size_t)m * KPv); h->B.reserve((size_t)n * KPv);
    h->pack.reserve((size_t)gram_pack_floats(KPv));
    h->G64w.reserve(GS); h->G64h.reserve(GS);
    h->dscal.reserve(1); h->red.reserve(8);
    if (!h->pinned) {
        CUDA_CHECK(cudaMallocHost(&h->pinned, sizeof(HostTrace)));
        h->pinned_bytes = sizeof(HostTrace);
    }
    HostTrace *ht = (HostTrace *)h->pinned;
    if (p.loss == SWNE_LOSS_MKL) ensure_sparse(h);
    if (engine == SWNE_ENGINE_TC) {
        tc_prepare(h->tc, h->dA, h->ldA, n, m, (double)__builtin_bit_cast(float, h->stats.max_bits), h->sm_count, st);
        if (tcx) tcx_prepare(h->tcx, h->tc);
    }
    bool tc_overflow = false;


    // ---- upload the init
    FitEvents ev;
    cudaEvent_t e0 = ev.e[0], e1 = ev.e[1], e2 = ev.e[2], e3 = ev.e[3], e4 = ev.e[4];
    cudaEventRecord(e0, st);
    // the init is validated on the device while it is converted (NNLM: negative or non-finite init is an error)
    h->flags.reserve(1);
    CUDA_CHECK(cudaMemsetAsync(h->flags.p, 0, sizeof(unsigned), st));
    if (device_init) {
        (*device_init)(KPv, gl[0] / N, m_glob);
    } else {
        upload_factor_colmajor(h, W, h->Wt.p, k, KPv, n, h->flags.p, 1u);
        CUDA_CHECK(cudaStreamSynchronize(st));  // stage buffer is reused
        upload_factor_cellmajor(h, H, h->H.p, k, KPv, (size_t)m, h->flags.p, 2u);
    }
    CUDA_CHECK(cudaMemsetAsync(h->dscal.p, 0, sizeof(FitScalars), st));
    cudaEventRecord(e1, st);
    {
        unsigned bad = 0;
        CUDA_CHECK(cudaMemcpyAsync(&bad, h->flags.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (bad) {
            FAIL(SWNE_ERR_BAD_ARG, "init %s has negative or non-finite entries", (bad & 1u) ? "W" : "H");
        }
    }

    const bool tcx_solve_simt = getenv("SWNE_TCX_SOLVE_SIMT") != nullptr;   // debug knob, read once per fit
    // mkl with cells sharded over ranks: the gene side runs coordinate by coordinate around an all-reduce (launch_kl_sharded);
    // SWNE_KL_SHARD_STEP takes the same kernels on one GPU (tests)
    // ... and the gene side of one GPU takes them too when a mean gene row is longer than the one-block-per-gene kernel can
    // stage in shared memory (it would run on the global arrays, measured 1.4-1.5x slower at 3000 x 20 000 / 100 000 cells)
    const bool kl_sharded = h->nranks > 1 || getenv("SWNE_KL_SHARD_STEP") != nullptr ||
                            (!mse_loss_of(p) && kl_rows_too_long(h, KPv));
    const bool slab_timeline = getenv("SWNE_TCX_TIMELINE") != nullptr;
    const bool three_streams = getenv("SWNE_TCX_STREAMS2") == nullptr;     // debug knob: both contractions on one stream
    const int genes_regions = getenv("SWNE_TCX_REGIONS") ? atoi(getenv("SWNE_TCX_REGIONS")) : 1;   // debug knob: 2 = double-buffered accumulators
    // slabs of the wide engine's H update: enough tiles per slab to keep every SM's contraction CTA busy (>= 4 tiles per SM)
    int tcx_slabs = 1;
    if (tcx) {
        const char *e = getenv("SWNE_TCX_SLABS");
        // measured (round 2, three streams): 650 k cells 1 / 2 / 3 / 4 / 8 slabs = 4.92 / 4.39 / 4.27 / 4.31 / 4.51 ms per iteration:
        // slabs of about 10 tiles per contraction CTA
        const int tiles = h->tcx.tiles_total;
        // (a 162.5 k-cell shard: 1 / 2 / 3 slabs = 1.47 / 1.44 / 1.37 ms)
        tcx_slabs = e ? atoi(e) : (tiles >= h->sm_count * 8 ? std::max(3, std::min(8, (2 * tiles + h->sm_count * 21 / 2) / (h->sm_count * 21))) : 1);
        if (h->nranks > 1) {
            // every rank must take the same path (the collectives differ): the smallest count wins
            double v = -(double)tcx_slabs;
            CUDA_CHECK(cudaMemcpyAsync(h->red.p, &v, sizeof(double), cudaMemcpyHostToDevice, st));
            NCCL_CHECK(g_nccl.AllReduce(h->red.p, h->red.p, 1, ncclFloat64, ncclMax, h->comm, h->stream));
            CUDA_CHECK(cudaMemcpyAsync(&v, h->red.p, sizeof(double), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
            tcx_slabs = (int)(-v);
        }
    }
    Prof prof;
    prof.on = p.profile != 0;
    prof.st = st;
    if (prof.on) { prof.get(); prof.pool_next = 0; }
    const float inner_tol = (float)p.inner_rel_tol;
    const bool mse = p.loss == SWNE_LOSS_MSE;

    bool have_B = false;  // the tcgen05 pass leaves A H' of the *new* H behind for the next W update
    bool peer_x = false;  // fused engine, several ranks: the per-iteration sums go over NVLink peer memory
    // fused engine: fp64 sums and GramPacks live in the engine's double-buffered state (half = iteration parity)
    auto g64w_of = [&](int it) { return fused ? h->tc.G64w + (size_t)(it & 1) * GS : h->G64w.p; };
    auto g64h_of = [&](int it) { return fused ? h->tc.G64h + (size_t)(it & 1) * GS : h->G64h.p; };
    if (fused) {
        // pre-pass (parity 1 = "iteration -1"): B = A H0', H0 H0' and the GramPack of the first W update in one launch
        prof.begin(0);
        tc_fit_begin(h->tc, h->Wt.p, h->H.p, st);
        CUDA_CHECK(cudaMemsetAsync(h->B.p, 0, sizeof(float) * (size_t)n * KPv, st));
        tc_pass(h->tc, 1, 1, h->nranks == 1, h->H.p, h->C.p, 0.f, 0.f, 0.f, p.alpha[0], p.alpha[1], k, p.inner_max_iter, 0.f, h->B.p,
                h->dscal.p, st);
        prof.end(0, 1 + tc_pass_launches(1));
        if (h->nranks > 1) {
            // the sums of the ranks: over NVLink peer memory (peer_xchg.cuh) when every rank could map every other's region,
            // else three grouped NCCL all-reduces and the finalize kernel
            peer_x = xchg_setup(h, xchg_payload_bytes((size_t)n * SWNE_MAX_K, (size_t)SWNE_MAX_K * SWNE_MAX_K + SWNE_MAX_K));
            if (peer_x)
                xchg_launch(h->xchg, h->B.p, (size_t)n * KPv, g64h_of(1), (size_t)GS, nullptr, h->tc.packH, k, p.alpha[0], p.alpha[1],
                            h->sm_count, st);
            else {
                allreduce(h, h->B.p, (size_t)n * KPv, g64h_of(1), GS);
                launch_finalize_gram(h, k, KPv, g64h_of(1), h->tc.packH, p.alpha[0], p.alpha[1]);
            }
        }
        have_B = true;
    } else {
        // H0 H0' (W update of iteration 0 needs it)
        prof.begin(4);
        launch_gram(h, KPv, h->H.p, m, h->G64h.p);
        prof.end(4, 1);
        if (h->nranks > 1) allreduce(h, nullptr, 0, h->G64h.p, GS);
    }

    double rel_err = p.rel_tol + 1.0, terr_last = 1e99;
    int i = 0, i_e = 0;
    int rc = SWNE_OK;

    auto error_block = [&](bool final_block) {
        // W'W is current (computed before the H update); H H' and the loss partials are current
        const bool want_mkl = p.full_traces > 0 || (p.full_traces == 0 && final_block);
        if (!mse || !want_mkl) {
            // nothing extra
        } else {
            prof.begin(4);
            CUDA_CHECK(cudaMemsetAsync(&h->dscal.p->kl_part, 0, sizeof(double), st));
            ops_of(KPv).kl_trace(h->dA, h->ldA, n, m, h->H.p, h->Wt.p, k, &h->dscal.p->kl_part, st);
            CUDA_CHECK(cudaGetLastError());
            prof.end(4, 1);
            if (h->nranks > 1) allreduce(h, nullptr, 0, &h->dscal.p->kl_part, 1);
        }
        // an mkl fit's loss partials come out of the cell side's update: rank-local sums over the rank's cells
        if (!mse && h->nranks > 1) allreduce(h, nullptr, 0, &h->dscal.p->kl_part, 2);
        // average.epochs counts the sweeps of all cells: the H-side count is rank-local until here
        if (h->nranks > 1)
            NCCL_CHECK(g_nccl.AllReduce(&h->dscal.p->sweeps_h, &h->dscal.p->sweeps_h, 1, ncclUint64, ncclSum, h->comm, h->stream));
        CUDA_CHECK(cudaMemcpyAsync(&ht->sc, h->dscal.p, sizeof(FitScalars), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(ht->G64w, g64w_of(i), sizeof(double) * GS, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(ht->G64h, g64h_of(i), sizeof(double) * GS, cudaMemcpyDeviceToHost, st));
        ht->tc_overflow = 0;
        // every rank must take the same decision (the refit is a collective): the flag is max-reduced first
        if (fused && h->nranks > 1)
            NCCL_CHECK(g_nccl.AllReduce(h->tc.state + TC_ST_OVF, h->tc.state + TC_ST_OVF, 1, ncclInt32, ncclMax, h->comm, h->stream));
        if (fused)
            CUDA_CHECK(cudaMemcpyAsync(&ht->tc_overflow, h->tc.state + TC_ST_OVF, sizeof(int), cudaMemcpyDeviceToHost, st));
        ht->xchg_err = 0;
        if (h->xchg.ready)
            CUDA_CHECK(cudaMemcpyAsync(&ht->xchg_err, h->xchg.region + h->xchg.flags_off + 2 * (size_t)h->nranks * 8 + 8, sizeof(int),
                                       cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (ht->xchg_err) FAIL(SWNE_ERR_NCCL, "peer-memory exchange: a rank did not publish its partial sums within 5 s");
        if (ht->tc_overflow) tc_overflow = true;
        double w2 = 0, wg = 0, w1 = 0, h2 = 0, hg = 0, h1 = 0, wwhh = 0, sum_ahat = 0;
        for (int a = 0; a < k; ++a) {
            for (int b = 0; b < k; ++b) {
                const double gw = ht->G64w[a * KPv + b], gh = ht->G64h[a * KPv + b];
                wg += gw; hg += gh; wwhh += gw * gh;
                if (a == b) { w2 += gw; h2 += gh; }
            }
            const double cw = ht->G64w[KPv * KPv + a], ch = ht->G64h[KPv * KPv + a];
            w1 += cw; h1 += ch; sum_ahat += cw * ch;
        }
        double mse_v, mkl_v;
        const double nan_v = nan("");
        if (mse) {
            mse_v = (sumsqA + ht->sc.loss_part) / N;
            mkl_v = want_mkl ? (klcA - ht->sc.kl_part + sum_ahat) / N : nan_v;
        } else {
            mse_v = (ht->sc.sq_part + wwhh) / N;
            mkl_v = (klcA - ht->sc.kl_part + sum_ahat) / N;
        }
        double terr = mse ? 0.5 * mse_v : mkl_v;
        if (p.alpha[0] != p.alpha[1]) terr += 0.5 * (p.alpha[0] - p.alpha[1]) * w2 / N;
        if (p.beta[0] != p.beta[1]) terr += 0.5 * (p.beta[0] - p.beta[1]) * h2 / N;
        if (p.alpha[1] != 0.0) terr += 0.5 * p.alpha[1] * wg / N;
        if (p.beta[1] != 0.0) terr += 0.5 * p.beta[1] * hg / N;
        if (p.alpha[2] != 0.0) terr += p.alpha[2] * w1 / N;
        if (p.beta[2] != 0.0) terr += p.beta[2] * h1 / N;
        const double epochs = (double)(ht->sc.sweeps_w + ht->sc.sweeps_h) / ((double)n + m_glob);
        if (mse_tr) mse_tr[i_e] = mse_v;
        if (mkl_tr) mkl_tr[i_e] = mkl_v;
        if (terr_tr) terr_tr[i_e] = terr;
        if (epo_tr) epo_tr[i_e] = epochs;
        rel_err = 2.0 * (terr_last - terr) / (terr_last + terr + TINY_NUM_D);
        terr_last = terr;
        if (p.verbose >= 2) fprintf(stderr, "%6d | %12.6g %12.6g %12.6g | %10.3g | %8.2f\n", i + 1, mse_v, mkl_v, terr, rel_err, epochs);
        ++i_e;
        // sweeps restart after every block (total_raw_iter = 0)
        CUDA_CHECK(cudaMemsetAsync(&h->dscal.p->sweeps_w, 0, 2 * sizeof(unsigned long long), st));
        int stop = (h->progress && h->progress(h->progress_user, i + 1, mse_v, mkl_v, terr, rel_err)) ? 1 : 0;
        if (h->nranks > 1 && h->progress) {
            // an interrupt on one rank must stop all of them (the next iteration is a collective; every rank that set a
            // callback reaches this point at the same iteration)
            double flag = (double)stop;
            CUDA_CHECK(cudaMemcpyAsync(h->red.p, &flag, sizeof(double), cudaMemcpyHostToDevice, st));
            NCCL_CHECK(g_nccl.AllReduce(h->red.p, h->red.p, 1, ncclFloat64, ncclMax, h->comm, h->stream));
            CUDA_CHECK(cudaMemcpyAsync(&flag, h->red.p, sizeof(double), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
            stop = flag != 0.0;
        }
        if (stop) rc = SWNE_ERR_INTERRUPTED;
    };

    cudaEventRecord(e2, st);
    for (; i < p.max_iter && fabs(rel_err) > p.rel_tol && rc == SWNE_OK && !tc_overflow; ++i) {
        if (fused) {
            // ---------------- fused engine: two kernels per iteration (each preceded by one constant-bank copy)
            //   W side: packH -> solve -> W, split image, W'W -> packW      H side: the pass (C = A W, solve, A H', H H' -> packH)
            prof.begin(3);
            tc_wside(h->tc, i & 1, h->Wt.p, h->B.p, (float)p.alpha[2], p.beta[0], p.beta[1], k, p.inner_max_iter, inner_tol, h->dscal.p, st);
            prof.end(3, tc_wside_launches());
            prof.begin(0);
            tc_pass(h->tc, 0, i & 1, h->nranks == 1, h->H.p, h->C.p, (float)p.beta[2], (float)(p.beta[0] - p.beta[1]), (float)p.beta[1],
                    p.alpha[0], p.alpha[1], k, p.inner_max_iter, inner_tol, h->B.p, h->dscal.p, st);
            prof.end(0, tc_pass_launches(0));
            if (h->nranks > 1) {
                prof.begin(6);
                if (peer_x)
                    xchg_launch(h->xchg, h->B.p, (size_t)n * KPv, g64h_of(i), (size_t)GS, &h->dscal.p->loss_part, h->tc.packH, k, p.alpha[0],
                                p.alpha[1], h->sm_count, st);
                else {
                    allreduce(h, h->B.p, (size_t)n * KPv, g64h_of(i), GS, &h->dscal.p->loss_part, 1);
                    launch_finalize_gram(h, k, KPv, g64h_of(i), h->tc.packH, p.alpha[0], p.alpha[1]);
                }
                prof.end(6, 2);
            }
            if (i % p.trace == 0) error_block(false);
            continue;
        }
        // loss partials restart every iteration (only the last one before a block is read)
        CUDA_CHECK(cudaMemsetAsync(h->dscal.p, 0, 3 * sizeof(double), st));
        if (mse) {
            // ---------------- W update: A' ~ H' Wt
            if (!have_B) {
                prof.begin(2);
                if (tcx) tcx_contract_genes(h->tcx, h->tc, KPv, h->H.p, h->B.p, st);
                else launch_contract_genes(h, KPv, h->H.p, h->B.p);
                prof.end(2, tcx ? 3 : 1);
                if (h->nranks > 1) { prof.begin(6); allreduce(h, h->B.p, (size_t)n * KPv, nullptr, 0); prof.end(6, 1); }
            }
            prof.begin(3);
            launch_finalize_gram(h, k, KPv, h->G64h.p, h->pack.p, p.alpha[0], p.alpha[1]);
            if (tcx && !tcx_solve_simt)
                tcx_solve(h->tcx, h->Wt.p, h->B.p, h->pack.p, KPv, (float)p.alpha[2], k, n, p.inner_max_iter, inner_tol,
                          &h->dscal.p->sweeps_w, nullptr, st);
            else
                launch_solve(h, k, KPv, h->Wt.p, h->B.p, h->pack.p, (float)p.alpha[2], n, p.inner_max_iter, inner_tol,
                             &h->dscal.p->sweeps_w, nullptr);
            prof.end(3, 2);
            // ---------------- H update: A ~ Wt' H
            prof.begin(4);
            launch_gram(h, KPv, h->Wt.p, n, h->G64w.p);
            launch_finalize_gram(h, k, KPv, h->G64w.p, h->pack.p, p.beta[0], p.beta[1]);
            prof.end(4, 2);
            if (tcx && !tcx_solve_simt && tcx_slabs > 1) {
                // ---- wide engine, slab pipeline: the cells go in slabs over two auxiliary streams, each slab running
                // contraction -> solves -> A H' of the new H.  The contractions are HBM-bound and the solves latency-bound, so
                // the solves of one slab run under the contraction of the next (profile slots are wall shares here).
                prof.begin(0);
                tcx_split_w(h->tcx, h->tc, KPv, h->Wt.p, st);
                tcx_solve_prepare(h->tcx, h->pack.p, KPv, st);
                CUDA_CHECK(cudaMemsetAsync(h->B.p, 0, sizeof(float) * (size_t)n * KPv, st));
                dev_zero(h->G64h.p, sizeof(double) * (size_t)GS, st);
                CUDA_CHECK(cudaEventRecord(h->ev_ready, st));
                const int tiles = h->tcx.tiles_total, per = (tiles + tcx_slabs - 1) / tcx_slabs;
                const int n_sl = (tiles + per - 1) / per;
                while ((int)h->ev_slab.size() < 2 * n_sl) {
                    cudaEvent_t e;
                    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    h->ev_slab.push_back(e);
                }
                cudaStream_t qc = h->aux[0], qs = h->aux[1];
                CUDA_CHECK(cudaStreamWaitEvent(qc, h->ev_ready, 0));
                CUDA_CHECK(cudaStreamWaitEvent(qs, h->ev_ready, 0));
                // SWNE_TCX_TIMELINE: event marks around every kernel of iteration 3 (debug: where the two streams overlap)
                std::vector<std::pair<std::string, cudaEvent_t>> marks;
                const bool tl_on = slab_timeline && i == 3;
                auto mark = [&](cudaStream_t q, const char *what, int sl) {
                    if (!tl_on) return;
                    cudaEvent_t e;
                    cudaEventCreate(&e);
                    cudaEventRecord(e, q);
                    marks.emplace_back(std::string(what) + " " + std::to_string(sl), e);
                };
                mark(qc, "begin", 0);
                auto slab = [&](int sl, int &t0, int &nt, size_t &c0, int &cells) {
                    t0 = sl * per; nt = std::min(per, tiles - t0);
                    c0 = (size_t)t0 * TC_TILE; cells = std::min(m - (int)c0, nt * TC_TILE);
                };
                auto enqueue_cells = [&](int sl) {
                    int t0, nt, cells; size_t c0;
                    slab(sl, t0, nt, c0, cells);
                    mark(qc, "cells  start", sl);
                    tcx_cells_slab(h->tcx, h->tc, KPv, h->C.p, t0, nt, qc);
                    mark(qc, "cells  end  ", sl);
                    CUDA_CHECK(cudaEventRecord(h->ev_slab[2 * sl], qc));
                    // the solve of the slab, on the low-priority stream
                    CUDA_CHECK(cudaStreamWaitEvent(qs, h->ev_slab[2 * sl], 0));
                    mark(qs, "solve  start", sl);
                    tcx_solve_run(h->tcx, h->H.p + c0 * KPv, h->C.p + c0 * KPv, h->pack.p, KPv, (float)p.beta[2], k, cells, p.inner_max_iter,
                                  inner_tol, &h->dscal.p->sweeps_h, &h->dscal.p->loss_part, qs);
                    mark(qs, "solve  end  ", sl);
                    CUDA_CHECK(cudaEventRecord(h->ev_slab[2 * sl + 1], qs));
                    // H H' and the row sums of this slab's H (fp64 atomics into the zeroed sums), beside the next contraction
                    ops_of(KPv).gram(h->H.p + c0 * KPv, cells, h->G64h.p, h->G64h.p + KPv * KPv, qs);
                    CUDA_CHECK(cudaGetLastError());
                    mark(qs, "gramH  end  ", sl);
                };
                // contraction stream: C(0), C(1), then A H'(s) as soon as solve(s) is done, each followed by C(s + 2): the stream
                // always has HBM-bound work queued while the solves of the two slabs in between run beside it
                // (three_streams: A H' on its own stream, every C = A W queued up front -- the two contractions then share the SMs)
                cudaStream_t qg = three_streams ? h->aux[2] : qc;
                if (three_streams) CUDA_CHECK(cudaStreamWaitEvent(qg, h->ev_ready, 0));
                const int ahead = three_streams ? n_sl : 2;
                for (int sl = 0; sl < std::min(ahead, n_sl); ++sl) enqueue_cells(sl);
                for (int sl = 0; sl < n_sl; ++sl) {
                    int t0, nt, cells; size_t c0;
                    slab(sl, t0, nt, c0, cells);
                    CUDA_CHECK(cudaStreamWaitEvent(qg, h->ev_slab[2 * sl + 1], 0));
                    mark(qg, "genes  start", sl);
                    tcx_genes_slab(h->tcx, h->tc, KPv, h->H.p, h->B.p, t0, nt, sl & 3, qg, genes_regions == 2 ? 2 : 1);
                    mark(qg, "genes  end  ", sl);
                    if (sl + ahead < n_sl) enqueue_cells(sl + ahead);
                }
                for (int q = 0; q < 3; ++q) {
                    CUDA_CHECK(cudaEventRecord(h->ev_done[q], h->aux[q]));
                    CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_done[q], 0));
                }
                if (tl_on) {
                    CUDA_CHECK(cudaStreamSynchronize(st));
                    for (size_t q = 1; q < marks.size(); ++q) {
                        float ms = 0.f;
                        cudaEventElapsedTime(&ms, marks[0].second, marks[q].second);
                        fprintf(stderr, "[tcx timeline] %-16s %8.3f ms\n", marks[q].first.c_str(), ms);
                    }
                    for (auto &mk : marks) cudaEventDestroy(mk.second);
                }
                prof.end(0, 3 + 6 * tcx_slabs);
                if (h->nranks > 1) {
                    prof.begin(6);
                    if (xchg_setup(h, xchg_payload_bytes((size_t)n * SWNE_MAX_K, (size_t)SWNE_MAX_K * SWNE_MAX_K + SWNE_MAX_K)))
                        xchg_launch(h->xchg, h->B.p, (size_t)n * KPv, h->G64h.p, (size_t)GS, &h->dscal.p->loss_part, nullptr, k, 0.0, 0.0,
                                    h->sm_count, st);
                    else
                        allreduce(h, h->B.p, (size_t)n * KPv, h->G64h.p, GS, &h->dscal.p->loss_part, 1);
                    prof.end(6, 1);
                }
                have_B = true;      // A H' of the new H is in place for the next W update
            } else {
                prof.begin(0);
                if (tcx) tcx_contract_cells(h->tcx, h->tc, KPv, h->Wt.p, h->C.p, st);
                else launch_contract_cells(h, KPv, h->Wt.p, h->C.p);
                prof.end(0, tcx ? 3 : 1);
                prof.begin(1);
                if (tcx && !tcx_solve_simt)
                    tcx_solve(h->tcx, h->H.p, h->C.p, h->pack.p, KPv, (float)p.beta[2], k, m, p.inner_max_iter, inner_tol,
                              &h->dscal.p->sweeps_h, &h->dscal.p->loss_part, st);
                else
                    launch_solve(h, k, KPv, h->H.p, h->C.p, h->pack.p, (float)p.beta[2], m, p.inner_max_iter, inner_tol,
                                 &h->dscal.p->sweeps_h, &h->dscal.p->loss_part);
                prof.end(1, 1);
                prof.begin(4);
                launch_gram(h, KPv, h->H.p, m, h->G64h.p);
                prof.end(4, 1);
                if (h->nranks > 1) {
                    prof.begin(6);
                    allreduce(h, nullptr, 0, h->G64h.p, GS, &h->dscal.p->loss_part, 1);
                    prof.end(6, 1);
                }
            }
        } else {
            // ---------------- mkl: W update on the by-gene view, H update on the by-cell view
            prof.begin(5);
            launch_finalize_gram(h, k, KPv, h->G64h.p, h->pack.p, 0.0, 0.0);
            int kl_launches = 1;
            if (kl_sharded)
                kl_launches = launch_kl_sharded(h, k, KPv, h->Wt.p, h->H.p, m, pack_colsum(h->pack.p, KPv), p.alpha, n, p.inner_max_iter,
                                                inner_tol, &h->dscal.p->sweeps_w);
            else
                launch_kl(h, k, KPv, h->sp.gptr.p, h->sp.gidx.p, h->sp.gval.p, h->Wt.p, h->H.p, pack_colsum(h->pack.p, KPv), p.alpha, n,
                          p.inner_max_iter, inner_tol, &h->dscal.p->sweeps_w, h->dscal.p, 0);
            prof.end(5, 1 + kl_launches);
            prof.begin(4);
            launch_gram(h, KPv, h->Wt.p, n, h->G64w.p);
            launch_finalize_gram(h, k, KPv, h->G64w.p, h->pack.p, 0.0, 0.0);
            prof.end(4, 2);
            prof.begin(1);
            launch_kl(h, k, KPv, h->sp.cptr.p, h->sp.cidx.p, h->sp.cval.p, h->H.p, h->Wt.p, pack_colsum(h->pack.p, KPv), p.beta, m,
                      p.inner_max_iter, inner_tol, &h->dscal.p->sweeps_h, h->dscal.p,
                      (i % p.trace == 0 || i == p.max_iter - 1) ? 1 : 0);
            prof.end(1, 1);
            prof.begin(4);
            launch_gram(h, KPv, h->H.p, m, h->G64h.p);
            prof.end(4, 1);
            // row sums of H over all cells (the next gene side's sum of the fixed factor) and H H' of the penalties
            if (h->nranks > 1) { prof.begin(6); allreduce(h, nullptr, 0, h->G64h.p, GS); prof.end(6, 1); }
        }
        if (i % p.trace == 0) error_block(false);
    }
    if (i > 0 && (i - 1) % p.trace != 0 && rc == SWNE_OK) {
        --i;  // error_block reports iteration i+1
        error_block(true);
        ++i;
    } else if (mse && p.full_traces == 0 && i_e > 0 && rc == SWNE_OK && mkl_tr) {
        // final state coincides with the last block: fill in its mkl
        CUDA_CHECK(cudaMemsetAsync(&h->dscal.p->kl_part, 0, sizeof(double), st));
        ops_of(KPv).kl_trace(h->dA, h->ldA, n, m, h->H.p, h->Wt.p, k, &h->dscal.p->kl_part, st);
        CUDA_CHECK(cudaGetLastError());
        if (h->nranks > 1) allreduce(h, nullptr, 0, &h->dscal.p->kl_part, 1);
        CUDA_CHECK(cudaMemcpyAsync(&ht->sc, h->dscal.p, sizeof(FitScalars), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        double sum_ahat = 0;
        for (int a = 0; a < k; ++a) sum_ahat += ht->G64w[KPv * KPv + a] * ht->G64h[KPv * KPv + a];
        mkl_tr[i_e - 1] = (klcA - ht->sc.kl_part + sum_ahat) / N;
    }
    cudaEventRecord(e3, st);
    if (tc_overflow) {
        // H grew by more than the 4096x headroom of its fp16 image between two iterations: the tcgen05 results of the
        // last iterations are not trustworthy.  Nothing was written back yet (W/H still hold the caller's init): redo
        // the fit on the fp32 engine.  Loud in verbose mode, never silent about which engine produced the result.
        swne_nmf_result_t dummy;
        prof.finish(&dummy);
        if (p.verbose) fprintf(stderr, "swne_nmf: fp16 range overflow in the tcgen05 engine, refitting with the fp32 engine\n");
        swne_nmf_params_t p2 = *pp;
        p2.engine = SWNE_ENGINE_SIMT;
        return fit_impl(h, &p2, W, H, mse_tr, mkl_tr, terr_tr, epo_tr, trace_len, res, device_init);
    }

    // ---- results (a device-seeded fit may leave them on the device: FindNumFactors only needs the errors)
    h->last_KPv = KPv;
    h->last_k = k;
    if (W) download_factor_colmajor(h, h->Wt.p, W, k, KPv, n);
    if (H) download_factor_cellmajor(h, h->H.p, H, k, KPv, (size_t)m);
    cudaEventRecord(e4, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1); res->upload_ms = ms;
    cudaEventElapsedTime(&ms, e2, e3); res->loop_ms = ms;
    cudaEventElapsedTime(&ms, e3, e4); res->download_ms = ms;
    prof.finish(res);
    res->n_iteration = i;
    res->n_trace = i_e;
    // NNLM warns when the iteration cap ended the loop with the tolerance not reached; rel_tol < 0 asks for a fixed
    // number of iterations and is not a failure to converge
    res->not_converged = (i >= p.max_iter && p.rel_tol >= 0.0 && fabs(rel_err) > p.rel_tol) ? 1 : 0;
    if (rc != SWNE_OK) { set_error("interrupted by the progress callback"); return rc; }
    return res->not_converged ? SWNE_WARN_NOT_CONVERGED : SWNE_OK;
}

// ------------------------------------------------------------------------------------------------
// nnlm: one update() on the resident matrix
// ------------------------------------------------------------------------------------------------
static int nnlm_impl(swne_nmf_handle_s *h, int side, int k, const double *X, double *coef, const double *alpha3, int loss,
                     int max_iter, double rel_tol, long long *sweeps)
{
    h->last_k = 0;     // h->H is about to be reused
    if (!h->dA || h->n == 0) FAIL(SWNE_ERR_NO_MATRIX, "no matrix resident");
    if (!X || !coef) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    if (k < 1 || k > SWNE_MAX_K) FAIL(SWNE_ERR_BAD_ARG, "k must be in 1..%d", SWNE_MAX_K);
    if (side != 0 && side != 1) FAIL(SWNE_ERR_BAD_ARG, "side must be 0 (solve H) or 1 (solve W)");
    // cells sharded over ranks: side 0 is local (W replicated); side 1 sums H H' and A H' over the ranks (mse) or the Newton
    // sums of every coordinate step (mkl, launch_kl_sharded); every rank then solves all genes (identical results everywhere,
    // like the W update of a fit)
    if (max_iter <= 0) max_iter = 10000;
    if (rel_tol <= 0) rel_tol = 1e-12;
    double pen[3] = {0, 0, 0};
    if (alpha3) for (int q = 0; q < 3; ++q) pen[q] = alpha3[q];
    const int n = h->n, m = h->m, KPv = kp_of(k), GS = KPv * KPv + KPv;
    cudaStream_t st = h->stream;
    h->Wt.reserve((size_t)n * KPv); h->H.reserve((size_t)m * KPv);
    h->C.reserve((size_t)std::max(n, m) * KPv);
    h->pack.reserve((size_t)gram_pack_floats(KPv));
    h->G64w.reserve(GS); h->dscal.reserve(1);
    CUDA_CHECK(cudaMemsetAsync(h->dscal.p, 0, sizeof(FitScalars), st));
    float *fixed, *solve;
    int fixed_rows, cols;
    if (side == 0) {
        validate_init(X, (size_t)n * k, "x"); validate_init(coef, (size_t)k * m, "coefficients");
        upload_factor_colmajor(h, X, h->Wt.p, k, KPv, n);
        CUDA_CHECK(cudaStreamSynchronize(st));
        upload_factor_cellmajor(h, coef, h->H.p, k, KPv, (size_t)m);
        fixed = h->Wt.p; solve = h->H.p; fixed_rows = n; cols = m;
    } else {
        validate_init(X, (size_t)k * m, "x"); validate_init(coef, (size_t)n * k, "coefficients");
        upload_factor_cellmajor(h, X, h->H.p, k, KPv, (size_t)m);
        CUDA_CHECK(cudaStreamSynchronize(st));
        upload_factor_colmajor(h, coef, h->Wt.p, k, KPv, n);
        fixed = h->H.p; solve = h->Wt.p; fixed_rows = m; cols = n;
    }
    launch_gram(h, KPv, fixed, fixed_rows, h->G64w.p);
    if (side == 1 && h->nranks > 1) allreduce(h, nullptr, 0, h->G64w.p, GS);
    if (loss == SWNE_LOSS_MSE) {
        launch_finalize_gram(h, k, KPv, h->G64w.p, h->pack.p, pen[0], pen[1]);
        if (side == 0) launch_contract_cells(h, KPv, fixed, h->C.p);
        else {
            launch_contract_genes(h, KPv, fixed, h->C.p);
            if (h->nranks > 1) allreduce(h, h->C.p, (size_t)n * KPv, nullptr, 0);
        }
        launch_solve(h, k, KPv, solve, h->C.p, h->pack.p, (float)pen[2], cols, max_iter, (float)rel_tol,
                     &h->dscal.p->sweeps_h, nullptr);
    } else if (loss == SWNE_LOSS_MKL) {
        ensure_sparse(h);
        launch_finalize_gram(h, k, KPv, h->G64w.p, h->pack.p, 0.0, 0.0);
        if (side == 0)
            launch_kl(h, k, KPv, h->sp.cptr.p, h->sp.cidx.p, h->sp.cval.p, solve, fixed, pack_colsum(h->pack.p, KPv), pen, cols, max_iter,
                      (float)rel_tol, &h->dscal.p->sweeps_h, h->dscal.p, 0);
        else if (h->nranks > 1 || getenv("SWNE_KL_SHARD_STEP") || kl_rows_too_long(h, KPv))
            launch_kl_sharded(h, k, KPv, solve, fixed, fixed_rows, pack_colsum(h->pack.p, KPv), pen, cols, max_iter, (float)rel_tol,
                              &h->dscal.p->sweeps_h);
        else
            launch_kl(h, k, KPv, h->sp.gptr.p, h->sp.gidx.p, h->sp.gval.p, solve, fixed, pack_colsum(h->pack.p, KPv), pen, cols, max_iter,
                      (float)rel_tol, &h->dscal.p->sweeps_h, h->dscal.p, 0);
    } else {
        FAIL(SWNE_ERR_BAD_ARG, "loss must be mse or mkl");
    }
    FitScalars sc;
    CUDA_CHECK(cudaMemcpyAsync(&sc, h->dscal.p, sizeof(sc), cudaMemcpyDeviceToHost, st));
    if (side == 0) download_factor_cellmajor(h, solve, coef, k, KPv, (size_t)m);
    else download_factor_colmajor(h, solve, coef, k, KPv, n);
    if (sweeps) *sweeps = (long long)sc.sweeps_h;
    return SWNE_OK;
}

// ------------------------------------------------------------------------------------------------
// NNDSVD from a randomized truncated SVD of the resident matrix (R/nmf.R:12-47)
//   A (n x m) = U S V'.  Range finder on the gene side: Y = A Omega (n x l), power iterations with
//   CholeskyQR2 re-orthonormalisation, then the small problem B = Q'A (l x m) through its Gram.
// ------------------------------------------------------------------------------------------------
__global__ void fill_gaussian_kernel(float *x, size_t rows, int l, int KP, uint64_t seed)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * (size_t)KP) return;
    const int f = (int)(t % KP);
    float v = 0.f;
    if (f < l) {
        // splitmix64 -> two uniforms -> Box-Muller
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (t + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        const float u1 = ((float)(uint32_t)(z >> 40) + 1.f) * (1.f / 16777217.f);
        const float u2 = (float)(uint32_t)(z & 0xFFFFFFu) * (1.f / 16777216.f);
        v = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
    }
    x[t] = v;
}
// X <- X R^{-1} with R'R = X'X (Cholesky QR), twice.  The l x l factorisation runs on the device (one block), so the
// nine re-orthonormalisations of a seed cost no host synchronisation; a rank-deficient sketch raises bit 0 of *flag.
static void cholesky_qr2(swne_nmf_handle_s *h, int KPv, int l, float *X, int rows, bool cell_side, double *G64dev, float *Tdev,
                         unsigned *flag)
{
    for (int pass = 0; pass < 2; ++pass) {
        launch_gram(h, KPv, X, rows, G64dev);
        if (h->nranks > 1 && cell_side) allreduce(h, nullptr, 0, G64dev, KPv * KPv + KPv);
        seed_chol_inv_kernel<<<1, 64, 0, h->stream>>>(G64dev, KPv, l, Tdev, flag);
        ops_of(KPv).right_multiply(X, (size_t)rows, Tdev, l, h->stream);
        CUDA_CHECK(cudaGetLastError());
    }
}

// Randomized truncated SVD of the resident matrix, left on the device: U [n][KPl] (gene side, replicated on all ranks),
// V [m][KPl] (cell side, this rank's cells), singular values on the host.  A (n x m) = U S V'.  Range finder on the gene
// side: Y = A Omega (n x l), power iterations with CholeskyQR2, then the small problem B = Q'A (l x m) through its Gram.
// The 2 + 2q sketch products are the same skinny contractions as the fit's; with cells sharded over ranks the gene-side
// products and the cell-side Grams are all-reduced.
// mu != nullptr: the SVD of the row-centred matrix A - mu 1' (gene means mu [n] on the device; the ICA seed's whitening),
// through rank-one corrections of the sketch products -- the centred matrix is never formed.
struct RsvdResult { int l, KPl; std::vector<double> sv; };
static RsvdResult rsvd_device(swne_nmf_handle_s *h, int k, int oversample, int power_iters, uint64_t seed, float *Y, float *Z,
                              const float *mu = nullptr)
{
    const int n = h->n, m = h->m;
    if (oversample < 0) oversample = 10;
    if (power_iters < 0) power_iters = 4;
    int l = std::min(std::min(n, m), k + oversample);
    if (l > SWNE_MAX_K) l = SWNE_MAX_K;
    if (k > l) FAIL(SWNE_ERR_BAD_ARG, "k must be <= %d and <= min(n, m)", SWNE_MAX_K);
    const int KPv = kp_of(l), GS = KPv * KPv + KPv;
    cudaStream_t st = h->stream;
    h->G.reserve((size_t)KPv * KPv); h->G64w.reserve(GS); h->flags.reserve(1);
    CUDA_CHECK(cudaMemsetAsync(h->flags.p, 0, sizeof(unsigned), st));
    // split-fp16 tcgen05 contractions (fp32-level accuracy) when the shape allows, else the fp32 kernels
    const bool use_tc = n >= 128 && m >= 128 && l <= TCX_KP && !getenv("SWNE_NNSVD_SIMT");
    if (use_tc) {
        tc_prepare(h->tc, h->dA, h->ldA, n, m, (double)__builtin_bit_cast(float, h->stats.max_bits), h->sm_count, st);
        tcx_prepare(h->tcx, h->tc);
    }
    h->red.reserve(4 * 64 + 8);
    double *cs = h->red.p;                                    // KPv doubles: the vector of a rank-one centring correction
    const int wgrid_n = std::min(h->sm_count * 4, std::max(1, n / 4)), wgrid_m = std::min(h->sm_count * 4, std::max(1, m / 4));
    auto a_times = [&](const float *Zin, float *Yout) {       // Yout (n x l) = A Zin, summed over the ranks' cells
        if (use_tc) tcx_contract_genes(h->tcx, h->tc, KPv, Zin, Yout, st);
        else launch_contract_genes(h, KPv, Zin, Yout);
        if (h->nranks > 1) allreduce(h, Yout, (size_t)n * KPv, nullptr, 0);
        if (mu) {                                             // (A - mu 1') Z = A Z - mu (1'Z)
            CUDA_CHECK(cudaMemsetAsync(cs, 0, sizeof(double) * KPv, st));
            seed_wcolsum_kernel<<<wgrid_m, 256, 0, st>>>(Zin, (size_t)m, KPv, nullptr, cs);
            if (h->nranks > 1) allreduce(h, nullptr, 0, cs, KPv);
            seed_rank1_sub_kernel<<<(unsigned)(((size_t)n * KPv + 255) / 256), 256, 0, st>>>(Yout, (size_t)n, KPv, mu, cs);
        }
    };
    auto at_times = [&](const float *Yin, float *Zout) {      // Zout (m x l) = A' Yin  (local cells)
        if (use_tc) tcx_contract_cells(h->tcx, h->tc, KPv, Yin, Zout, st);
        else launch_contract_cells(h, KPv, Yin, Zout);
        if (mu) {                                             // (A - mu 1')' Y = A'Y - 1 (mu'Y)
            CUDA_CHECK(cudaMemsetAsync(cs, 0, sizeof(double) * KPv, st));
            seed_wcolsum_kernel<<<wgrid_n, 256, 0, st>>>(Yin, (size_t)n, KPv, mu, cs);
            seed_rank1_sub_kernel<<<(unsigned)(((size_t)m * KPv + 255) / 256), 256, 0, st>>>(Zout, (size_t)m, KPv, nullptr, cs);
        }
    };
    const size_t totz = (size_t)m * KPv;
    // different ranks draw different rows of the test matrix
    fill_gaussian_kernel<<<(unsigned)((totz + 255) / 256), 256, 0, st>>>(Z, (size_t)m, l, KPv, seed + 0x5851F42D4C957F2Dull * (uint64_t)h->rank);
    CUDA_CHECK(cudaGetLastError());
    a_times(Z, Y);                                          // Y = A Omega
    cholesky_qr2(h, KPv, l, Y, n, false, h->G64w.p, h->G.p, h->flags.p);
    for (int q = 0; q < power_iters; ++q) {
        at_times(Y, Z);                                     // Z = A' Q
        cholesky_qr2(h, KPv, l, Z, m, true, h->G64w.p, h->G.p, h->flags.p);
        a_times(Z, Y);                                      // Y = A Z
        cholesky_qr2(h, KPv, l, Y, n, false, h->G64w.p, h->G.p, h->flags.p);
    }
    at_times(Y, Z);                                         // Z = A' Q  (m x l) = B'
    launch_gram(h, KPv, Z, m, h->G64w.p);                   // B B' = Z'Z
    if (h->nranks > 1) allreduce(h, nullptr, 0, h->G64w.p, GS);
    std::vector<double> G(GS);
    unsigned bad = 0;
    CUDA_CHECK(cudaMemcpyAsync(G.data(), h->G64w.p, sizeof(double) * GS, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(&bad, h->flags.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));      // the only host synchronisation of the seed
    if (bad) FAIL(SWNE_ERR_BAD_ARG, "randomized SVD: rank deficient sketch (k too large for this matrix?)");
    std::vector<double> S(l * l), evec(l * l), eval(l);
    for (int a = 0; a < l; ++a)
        for (int b = 0; b < l; ++b) S[a * l + b] = G[a * KPv + b];
    small_jacobi_eigh(S.data(), l, eval.data(), evec.data());  // descending, evec columns
    // U = Q Uhat (n x l), V = Z Uhat S^-1 (m x l): right-multiply on device (two small uploads, stream ordered)
    std::vector<float> T(2 * KPv * KPv, 0.f);
    for (int a = 0; a < l; ++a)
        for (int b = 0; b < l; ++b) {
            T[a * KPv + b] = (float)evec[a * l + b];
            const double s = sqrt(std::max(eval[b], 0.0));
            T[KPv * KPv + a * KPv + b] = (s > 0) ? (float)(evec[a * l + b] / s) : 0.f;
        }
    h->G.reserve((size_t)2 * KPv * KPv);
    CUDA_CHECK(cudaMemcpyAsync(h->G.p, T.data(), sizeof(float) * T.size(), cudaMemcpyHostToDevice, st));
    ops_of(KPv).right_multiply(Y, (size_t)n, h->G.p, l, st);
    ops_of(KPv).right_multiply(Z, (size_t)m, h->G.p + KPv * KPv, l, st);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(st));      // T is a host temporary
    RsvdResult r;
    r.l = l; r.KPl = KPv;
    r.sv.resize(l);
    for (int f = 0; f < l; ++f) r.sv[f] = sqrt(std::max(eval[f], 0.0));
    return r;
}

// NNDSVD (R/nmf.R:12-47) of the device-resident triplets: Wt_out [n][KPd], H_out [m][KPd] hold the split factors, with
// RunNMF's zero replacement (R/nmf.R:121-127) fused in when fill_max > 0 (eps = 1e-6) and plain zeros otherwise.
static void nnsvd_split_device(swne_nmf_handle_s *h, const RsvdResult &r, int k, const float *U, const float *V, float *Wt_out,
                               float *H_out, int KPd, float eps, float fill_max, uint64_t seed)
{
    const int n = h->n, m = h->m, KPl = r.KPl;
    cudaStream_t st = h->stream;
    h->red.reserve(4 * 64 + 8);
    double *norms = h->red.p;      // [u+ | u- | v+ | v-] x KPl
    CUDA_CHECK(cudaMemsetAsync(norms, 0, sizeof(double) * 4 * KPl, st));
    seed_posneg_kernel<<<std::min(h->sm_count * 4, (n + 3) / 4), 256, 0, st>>>(U, (size_t)n, KPl, norms);
    seed_posneg_kernel<<<std::min(h->sm_count * 4, (m + 3) / 4), 256, 0, st>>>(V, (size_t)m, KPl, norms + 2 * KPl);
    CUDA_CHECK(cudaGetLastError());
    if (h->nranks > 1) allreduce(h, nullptr, 0, norms + 2 * KPl, 2 * KPl);      // V is sharded by cells, U replicated
    std::vector<double> hn(4 * KPl);
    CUDA_CHECK(cudaMemcpyAsync(hn.data(), norms, sizeof(double) * 4 * KPl, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    // per factor: which sign part is kept and the scale of the u and v sides (R/nmf.R:24-44; ties favour the positive part)
    std::vector<float> coef(2 * 64, 0.f);
    std::vector<int> mode(2 * 64, 1);
    for (int f = 0; f < k; ++f) {
        const double s = r.sv[f];
        if (f == 0) {
            coef[0] = coef[64] = (float)sqrt(s);
            mode[0] = mode[64] = 0;
            continue;
        }
        const double nup = sqrt(hn[f]), nun = sqrt(hn[KPl + f]), vp = sqrt(hn[2 * KPl + f]), vn = sqrt(hn[3 * KPl + f]);
        const double termp = nup * vp, termn = nun * vn;
        if (termp >= termn) {
            const double c = sqrt(s * termp);
            coef[f] = nup > 0 ? (float)(c / nup) : 0.f; coef[64 + f] = vp > 0 ? (float)(c / vp) : 0.f;
            mode[f] = mode[64 + f] = 1;
        } else {
            const double c = sqrt(s * termn);
            coef[f] = nun > 0 ? (float)(-c / nun) : 0.f; coef[64 + f] = vn > 0 ? (float)(-c / vn) : 0.f;
            mode[f] = mode[64 + f] = 2;
        }
    }
    h->G.reserve(4 * 64);
    float *dcoef = h->G.p;
    int *dmode = reinterpret_cast<int *>(h->G.p + 2 * 64);
    CUDA_CHECK(cudaMemcpyAsync(dcoef, coef.data(), sizeof(float) * 2 * 64, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(dmode, mode.data(), sizeof(int) * 2 * 64, cudaMemcpyHostToDevice, st));
    const size_t tw = (size_t)n * KPd, th = (size_t)m * KPd;
    seed_apply_kernel<<<(unsigned)((tw + 255) / 256), 256, 0, st>>>(U, KPl, Wt_out, KPd, (size_t)n, k, dcoef, dmode, eps, fill_max, seed ^ 0x57ull);
    seed_apply_kernel<<<(unsigned)((th + 255) / 256), 256, 0, st>>>(V, KPl, H_out, KPd, (size_t)m, k, dcoef + 64, dmode + 64, eps, fill_max,
                                                                   (seed ^ 0x48ull) + 0x9E3779B97F4A7C15ull * (uint64_t)h->rank);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(st));      // coef / mode are host temporaries
}

static int nnsvd_impl(swne_nmf_handle_s *h, int k, int oversample, int power_iters, uint64_t seed, double *W, double *H,
                      double *sv)
{
    h->last_k = 0;     // h->H is about to be reused
    if (!h->dA || h->n == 0) FAIL(SWNE_ERR_NO_MATRIX, "no matrix resident");
    if (!W || !H) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    const int n = h->n, m = h->m;
    if (k < 1 || k > std::min(n, m)) FAIL(SWNE_ERR_BAD_ARG, "k must be in 1..min(n,m)");
    const int lmax = std::min(std::min(std::min(n, m), k + (oversample < 0 ? 10 : oversample)), SWNE_MAX_K);
    const int KPl = kp_of(lmax), KPd = kp_of(k);
    h->B.reserve((size_t)n * KPl); h->C.reserve((size_t)m * KPl);
    h->Wt.reserve((size_t)n * KPd); h->H.reserve((size_t)m * KPd);
    const RsvdResult r = rsvd_device(h, k, oversample, power_iters, seed, h->B.p, h->C.p);
    nnsvd_split_device(h, r, k, h->B.p, h->C.p, h->Wt.p, h->H.p, KPd, 0.f, 0.f, seed);
    if (sv) for (int f = 0; f < k; ++f) sv[f] = r.sv[f];
    download_factor_colmajor(h, h->Wt.p, W, k, KPd, n);
    download_factor_cellmajor(h, h->H.p, H, k, KPd, (size_t)m);
    return SWNE_OK;
}

// ------------------------------------------------------------------------------------------------
// ica_init on the device (R/nmf.R:56-66).  icafast(t(A), nc = k) whitens with the top-k eigenpairs of the gene covariance
// of the centred data; those are the leading singular triplets of A - rowMeans(A) 1' = U S V', and the whitened data are
// sqrt(m) V.  So: centred randomized SVD (the nnsvd seed's machinery) -> FastICA fixed point on sqrt(m) V[:, :k]
// (one pass over m x k per iteration, the k x k algebra on the host) -> S = Y R, M = U diag(s / sqrt(m)) R ->
//   ica.fast = F:  W = M, H = S'        ica.fast = T (irlba(nv = 50) branch):  W = (A - rowMeans) S, H = S'
// followed by RunNMF's zero replacement (negative entries count as "< 1e-6") straight into the fit's layouts.
// Components come out up to sign and order, as they do from LAPACK in the reference.
// ------------------------------------------------------------------------------------------------
static void ica_seed_device(swne_nmf_handle_s *h, int k, bool ica_fast, uint64_t seed, int KPd, float fill_max, double m_glob,
                            float eps = 1e-6f)
{
    const int n = h->n, m = h->m;
    cudaStream_t st = h->stream;
    if (k > std::min(n, m)) FAIL(SWNE_ERR_BAD_ARG, "ica init needs k <= min(n, m)");
    if (ica_fast && k > 50) FAIL(SWNE_ERR_BAD_ARG, "ica.fast uses 50 principal components: k must be <= 50");
    // ---- gene means
    DevBuf<double> gs;
    DevBuf<float> mu;
    gs.reserve((size_t)n); mu.reserve((size_t)n);
    CUDA_CHECK(cudaMemsetAsync(gs.p, 0, sizeof(double) * n, st));
    {
        const int cpb = std::max(256, (m + h->sm_count * 4 - 1) / (h->sm_count * 4));
        dim3 grid((n + 255) / 256, (m + cpb - 1) / cpb);
        seed_gene_sums_kernel<<<grid, 256, 0, st>>>(h->dA, h->ldA, n, m, cpb, gs.p);
        if (h->nranks > 1) allreduce(h, nullptr, 0, gs.p, (size_t)n);
        seed_scale_to_float_kernel<<<(n + 255) / 256, 256, 0, st>>>(gs.p, mu.p, n, 1.0 / m_glob);
        CUDA_CHECK(cudaGetLastError());
    }
    // ---- centred truncated SVD: l = k + 10 components (ica.fast: the 50 of irlba(nv = 50), capped by the library's 64)
    const int over = ica_fast ? std::max(0, std::min(50, std::min(n, m)) - k) : 10;
    const int lmax = std::min(std::min(std::min(n, m), k + over), SWNE_MAX_K);
    const int KPl = kp_of(lmax), KPk = kp_of(k);
    h->B.reserve((size_t)n * std::max(KPl, KPd)); h->C.reserve((size_t)m * std::max(KPl, KPd));
    float *U = h->B.p, *V = h->C.p;
    const RsvdResult r = rsvd_device(h, k, over, -1, seed, U, V, mu.p);
    // ---- whitened data Y = sqrt(m) V[:, :k]   [m][KPk]
    DevBuf<float> Yb, small;
    DevBuf<double> acc;
    Yb.reserve((size_t)m * KPk); small.reserve(4 * 64 + (size_t)2 * 64 * 64); acc.reserve((size_t)KPk * KPk + KPk);
    float *dcoef = small.p;
    int *dmode = reinterpret_cast<int *>(small.p + 2 * 64);
    float *dR = small.p + 4 * 64;
    std::vector<float> coef(2 * 64, 1.f);
    std::vector<int> mode(2 * 64, 3);
    for (int f = 0; f < 64; ++f) coef[f] = (float)sqrt(m_glob);
    CUDA_CHECK(cudaMemcpyAsync(dcoef, coef.data(), sizeof(float) * 2 * 64, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(dmode, mode.data(), sizeof(int) * 2 * 64, cudaMemcpyHostToDevice, st));
    const size_t ty = (size_t)m * KPk;
    seed_apply_kernel<<<(unsigned)((ty + 255) / 256), 256, 0, st>>>(V, r.KPl, Yb.p, KPk, (size_t)m, k, dcoef, dmode, -INFINITY, 0.f, 0);
    CUDA_CHECK(cudaGetLastError());
    // ---- FastICA fixed point (icafast: alg "par", fun "logcosh", Rmat = I, maxit 50, tol 1e-4)
    std::vector<double> R(k * k, 0.0), Rn(k * k), T((size_t)KPk * KPk + KPk), RtR(k * k), evec(k * k), eval(k), P(k * k);
    for (int a = 0; a < k; ++a) R[a * k + a] = 1.0;
    std::vector<float> Rf((size_t)KPk * KPk);
    const int grid = std::min(h->sm_count * 2, std::max(1, (m + 31) / 32));
    double vtol = 1.0;
    for (int it = 0; it < 50 && vtol > 1e-4; ++it) {
        std::fill(Rf.begin(), Rf.end(), 0.f);
        for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) Rf[a * KPk + b] = (float)R[a * k + b];
        CUDA_CHECK(cudaMemcpyAsync(dR, Rf.data(), sizeof(float) * Rf.size(), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemsetAsync(acc.p, 0, sizeof(double) * T.size(), st));
        ops_of(KPk).ica_step(Yb.p, m, k, dR, acc.p, grid, st);
        CUDA_CHECK(cudaGetLastError());
        if (h->nranks > 1) allreduce(h, nullptr, 0, acc.p, T.size());
        CUDA_CHECK(cudaMemcpyAsync(T.data(), acc.p, sizeof(double) * T.size(), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        // R+ = Y'g / nobs - R diag(mean(1 - g^2)), then the symmetric orthogonalisation R+ (R+' R+)^(-1/2)
        for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) Rn[a * k + b] = T[a * KPk + b] / m_glob - R[a * k + b] * (T[(size_t)KPk * KPk + b] / m_glob);
        for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) {
                double sacc = 0.0;
                for (int q = 0; q < k; ++q) sacc += Rn[q * k + a] * Rn[q * k + b];
                RtR[a * k + b] = sacc;
            }
        small_jacobi_eigh(RtR.data(), k, eval.data(), evec.data());
        for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) {
                double sacc = 0.0;
                for (int q = 0; q < k; ++q) sacc += evec[a * k + q] * evec[b * k + q] / sqrt(std::max(eval[q], 1e-300));
                P[a * k + b] = sacc;
            }
        std::vector<double> Ro(k * k);
        for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) {
                double sacc = 0.0;
                for (int q = 0; q < k; ++q) sacc += Rn[a * k + q] * P[q * k + b];
                Ro[a * k + b] = sacc;
            }
        double mn = 1e300;
        for (int b = 0; b < k; ++b) {
            double sacc = 0.0;
            for (int a = 0; a < k; ++a) sacc += R[a * k + b] * Ro[a * k + b];
            mn = std::min(mn, fabs(sacc));
        }
        vtol = 1.0 - mn;
        R = Ro;
    }
    // ---- order by variance accounted for: vaf_c = sum_g M[g][c]^2 = sum_a R[a][c]^2 s_a^2 / m   (U orthonormal)
    std::vector<int> ix(k);
    std::vector<double> vaf(k, 0.0);
    for (int c = 0; c < k; ++c) {
        ix[c] = c;
        for (int a = 0; a < k; ++a) vaf[c] += R[a * k + c] * R[a * k + c] * r.sv[a] * r.sv[a] / m_glob;
    }
    std::stable_sort(ix.begin(), ix.end(), [&](int a, int b) { return vaf[a] > vaf[b]; });
    // ---- S = Y R[:, ix]  (in place), H = S' with the zero replacement
    std::fill(Rf.begin(), Rf.end(), 0.f);
    for (int a = 0; a < k; ++a)
        for (int c = 0; c < k; ++c) Rf[a * KPk + c] = (float)R[a * k + ix[c]];
    CUDA_CHECK(cudaMemcpyAsync(dR, Rf.data(), sizeof(float) * Rf.size(), cudaMemcpyHostToDevice, st));
    ops_of(KPk).right_multiply(Yb.p, (size_t)m, dR, k, st);
    for (int f = 0; f < 64; ++f) coef[f] = 1.f;
    CUDA_CHECK(cudaMemcpyAsync(dcoef, coef.data(), sizeof(float) * 2 * 64, cudaMemcpyHostToDevice, st));
    const size_t th = (size_t)m * KPd, tw = (size_t)n * KPd;
    if (!ica_fast) {
        seed_apply_kernel<<<(unsigned)((th + 255) / 256), 256, 0, st>>>(Yb.p, KPk, h->H.p, KPd, (size_t)m, k, dcoef, dmode, eps, fill_max,
                                                                       (seed ^ 0x48ull) + 0x9E3779B97F4A7C15ull * (uint64_t)h->rank);
        // W = M = U diag(s / sqrt(m)) R[:, ix]
        std::vector<float> T2((size_t)r.KPl * r.KPl, 0.f);
        for (int a = 0; a < k; ++a)
            for (int c = 0; c < k; ++c) T2[(size_t)a * r.KPl + c] = (float)(r.sv[a] / sqrt(m_glob) * R[a * k + ix[c]]);
        DevBuf<float> dT2;
        dT2.reserve(T2.size());
        CUDA_CHECK(cudaMemcpyAsync(dT2.p, T2.data(), sizeof(float) * T2.size(), cudaMemcpyHostToDevice, st));
        ops_of(r.KPl).right_multiply(U, (size_t)n, dT2.p, k, st);
        seed_apply_kernel<<<(unsigned)((tw + 255) / 256), 256, 0, st>>>(U, r.KPl, h->Wt.p, KPd, (size_t)n, k, dcoef, dmode, eps, fill_max,
                                                                       seed ^ 0x57ull);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaStreamSynchronize(st));
        dT2.release();
    } else {
        // W = (A - mu 1') S = A S - mu (1'S): one more skinny contraction on the signed S (before its replacement)
        const bool use_tc = n >= 128 && m >= 128 && KPk <= TCX_KP;
        float *Wraw = U;                                   // [n][KPk]; U is no longer needed
        if (use_tc) tcx_contract_genes(h->tcx, h->tc, KPk, Yb.p, Wraw, st);
        else launch_contract_genes(h, KPk, Yb.p, Wraw);
        if (h->nranks > 1) allreduce(h, Wraw, (size_t)n * KPk, nullptr, 0);
        h->red.reserve(4 * 64 + 8);
        CUDA_CHECK(cudaMemsetAsync(h->red.p, 0, sizeof(double) * KPk, st));
        seed_wcolsum_kernel<<<std::min(h->sm_count * 4, std::max(1, m / 4)), 256, 0, st>>>(Yb.p, (size_t)m, KPk, nullptr, h->red.p);
        if (h->nranks > 1) allreduce(h, nullptr, 0, h->red.p, KPk);
        seed_rank1_sub_kernel<<<(unsigned)(((size_t)n * KPk + 255) / 256), 256, 0, st>>>(Wraw, (size_t)n, KPk, mu.p, h->red.p);
        seed_apply_kernel<<<(unsigned)((tw + 255) / 256), 256, 0, st>>>(Wraw, KPk, h->Wt.p, KPd, (size_t)n, k, dcoef, dmode, eps, fill_max,
                                                                       seed ^ 0x57ull);
        seed_apply_kernel<<<(unsigned)((th + 255) / 256), 256, 0, st>>>(Yb.p, KPk, h->H.p, KPd, (size_t)m, k, dcoef, dmode, eps, fill_max,
                                                                       (seed ^ 0x48ull) + 0x9E3779B97F4A7C15ull * (uint64_t)h->rank);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
    gs.release(); mu.release(); Yb.release(); small.release(); acc.release();
}

// ------------------------------------------------------------------------------------------------
// RunNMF in one call (R/nmf.R:96-134): seed on the device -> zero replacement -> fit.  Nothing crosses the host
// between the seed and the fit.
// ------------------------------------------------------------------------------------------------
static int run_nmf_impl(swne_nmf_handle_s *h, const swne_nmf_params_t *pp, int init, uint64_t seed, double *W, double *H,
                        double *mse_tr, double *mkl_tr, double *terr_tr, double *epo_tr, int trace_len, swne_nmf_result_t *res)
{
    if (!h->dA || h->n == 0) FAIL(SWNE_ERR_NO_MATRIX, "no matrix resident: call swne_nmf_set_matrix_* first");
    if (!pp) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    if (init == SWNE_INIT_GIVEN) return fit_impl(h, pp, W, H, mse_tr, mkl_tr, terr_tr, epo_tr, trace_len, res);
    const int k = pp->k, n = h->n, m = h->m;
    if (k < 1 || k > SWNE_MAX_K) FAIL(SWNE_ERR_BAD_ARG, "k must be in 1..%d", SWNE_MAX_K);
    cudaStream_t st = h->stream;
    DeviceInitFn fn;
    if (init == SWNE_INIT_RANDOM) {
        // NNLM's own start: U(0,1) * 0.01.  W identical on every rank, H per shard.
        fn = [=](int KPv, double, double) {
            const size_t tw = (size_t)n * KPv, th = (size_t)m * KPv;
            seed_uniform_kernel<<<(unsigned)((tw + 255) / 256), 256, 0, st>>>(h->Wt.p, (size_t)n, k, KPv, 0.01f, seed ^ 0x57ull);
            seed_uniform_kernel<<<(unsigned)((th + 255) / 256), 256, 0, st>>>(h->H.p, (size_t)m, k, KPv, 0.01f,
                                                                            (seed ^ 0x48ull) + 0x9E3779B97F4A7C15ull * (uint64_t)h->rank);
            CUDA_CHECK(cudaGetLastError());
        };
    } else if (init == SWNE_INIT_NNSVD) {
        if (k > std::min(n, m)) FAIL(SWNE_ERR_BAD_ARG, "nnsvd init needs k <= min(n, m)");
        fn = [=](int KPv, double mean_a, double) {
            const int lmax = std::min(std::min(std::min(n, m), k + 10), SWNE_MAX_K);
            const int KPl = kp_of(lmax);
            h->B.reserve((size_t)n * std::max(KPl, KPv)); h->C.reserve((size_t)m * std::max(KPl, KPv));
            const RsvdResult r = rsvd_device(h, k, -1, -1, seed, h->B.p, h->C.p);
            // R/nmf.R:121-127: entries below 1e-6 -> 0, zeros -> runif(0, mean(A) / 100)
            nnsvd_split_device(h, r, k, h->B.p, h->C.p, h->Wt.p, h->H.p, KPv, 1e-6f, (float)(mean_a / 100.0), seed);
        };
    } else if (init == SWNE_INIT_ICA || init == SWNE_INIT_ICA_FAST) {
        fn = [=](int KPv, double mean_a, double m_glob) {
            ica_seed_device(h, k, init == SWNE_INIT_ICA_FAST, seed, KPv, (float)(mean_a / 100.0), m_glob);
        };
    } else {
        FAIL(SWNE_ERR_BAD_ARG, "Invalid initialization method");
    }
    return fit_impl(h, pp, W, H, mse_tr, mkl_tr, terr_tr, epo_tr, trace_len, res, &fn);
}

// ica_init (R/nmf.R:56-66) as a stand-alone call: the signed W (n x k) and H (k x m), before RunNMF's zero replacement
static int ica_init_impl(swne_nmf_handle_s *h, int k, int ica_fast, uint64_t seed, double *W, double *H)
{
    h->last_k = 0;     // h->H is about to be reused
    if (!h->dA || h->n == 0) FAIL(SWNE_ERR_NO_MATRIX, "no matrix resident");
    if (!W || !H) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    if (k < 1 || k > SWNE_MAX_K) FAIL(SWNE_ERR_BAD_ARG, "k must be in 1..%d", SWNE_MAX_K);
    const int n = h->n, m = h->m, KPd = kp_of(k);
    h->Wt.reserve((size_t)n * KPd); h->H.reserve((size_t)m * KPd);
    double m_glob = (double)m;
    if (h->nranks > 1) {
        h->red.reserve(8);
        CUDA_CHECK(cudaMemcpyAsync(h->red.p, &m_glob, sizeof(double), cudaMemcpyHostToDevice, h->stream));
        allreduce(h, nullptr, 0, h->red.p, 1);
        CUDA_CHECK(cudaMemcpyAsync(&m_glob, h->red.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    ica_seed_device(h, k, ica_fast != 0, seed, KPd, 0.f, m_glob, -INFINITY);
    download_factor_colmajor(h, h->Wt.p, W, k, KPd, n);
    download_factor_cellmajor(h, h->H.p, H, k, KPd, (size_t)m);
    return SWNE_OK;
}

// ------------------------------------------------------------------------------------------------
// FindNumFactors (R/nmf.R:157-205) on the resident matrix, as one library call: for every k two mse fits from NNLM's
// random start (R/nmf.R:167-168; `loss` only selects the error metric) -- on A and on
// matrix(A[sample(length(A))], nrow(A)) (R/nmf.R:163) -- and the reconstruction errors (R/nmf.R:170-181).  A, the
// permuted copy, every seed, W and H stay on the device; only 2 x |k.range| doubles come back.
// ------------------------------------------------------------------------------------------------
__global__ void fnf_keys_kernel(uint64_t *keys, size_t count, uint64_t seed)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (t + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    keys[t] = z ^ (z >> 31);
}
// compact [m][n] <-> padded [m][ldA]
__global__ void fnf_repad_kernel(const float *src, size_t src_ld, float *dst, size_t dst_ld, int n, size_t m)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * dst_ld) return;
    const size_t j = t / dst_ld, g = t % dst_ld;
    dst[t] = (g < (size_t)n && g < src_ld) ? src[j * src_ld + g] : 0.f;
}

static int find_num_factors_impl(swne_nmf_handle_s *h, const int *k_range, int nk, int loss_metric, int max_iter, uint64_t seed,
                                 int engine, double *k_err, int *k_pick, double *err_diff)
{
    struct Invalidate { swne_nmf_handle_s *h; ~Invalidate() { h->last_k = 0; } } invalidate_resident_h{h};   // the last fit is of the permuted matrix
    if (!h->dA || h->n == 0) FAIL(SWNE_ERR_NO_MATRIX, "no matrix resident");
    if (!k_range || !k_err) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    // R/nmf.R:158's rule belongs to the selection; a replica that only wants the errors of its share of k.range (k_pick and
    // err_diff NULL) may pass fewer values
    if (nk < 1 || (nk < 3 && (k_pick || err_diff))) FAIL(SWNE_ERR_BAD_ARG, "k.range sequence must have at least 3 values");
    if (loss_metric != SWNE_LOSS_MSE && loss_metric != SWNE_LOSS_MKL) FAIL(SWNE_ERR_BAD_ARG, "Invalid loss function");
    if (h->nranks > 1) FAIL(SWNE_ERR_UNSUPPORTED, "FindNumFactors runs its fits as replicas: call it on a single-GPU handle per k subset");
    for (int c = 0; c < nk; ++c)
        if (k_range[c] < 1 || k_range[c] > SWNE_MAX_K) FAIL(SWNE_ERR_BAD_ARG, "k must be in 1..%d", SWNE_MAX_K);
    const int n = h->n, m = h->m;
    cudaStream_t st = h->stream;
    if (max_iter <= 0) max_iter = 250;
    // ---- the permuted matrix: sort the entries by random 64-bit keys
    const size_t count = (size_t)n * m;
    DevBuf<float> Arand, vin, vout;
    DevBuf<uint64_t> kin, kout;
    DevBuf<unsigned char> tmp;
    Arand.reserve((size_t)m * h->ldA);
    kin.reserve(count); kout.reserve(count); vout.reserve(count);
    const float *vals = h->dA;
    if (h->ldA != (size_t)n) {
        vin.reserve(count);
        fnf_repad_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(h->dA, h->ldA, vin.p, (size_t)n, n, (size_t)m);
        vals = vin.p;
    }
    fnf_keys_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(kin.p, count, seed ^ 0xA5A5A5A5ull);
    CUDA_CHECK(cudaGetLastError());
    size_t tmp_bytes = 0;
    if (count > 0x7fffffffull) FAIL(SWNE_ERR_UNSUPPORTED, "FindNumFactors: more than 2^31 - 1 matrix entries");
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kin.p, kout.p, vals, vout.p, (int)count, 0, 64, st));
    tmp.reserve(tmp_bytes);
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kin.p, kout.p, vals, vout.p, (int)count, 0, 64, st));
    fnf_repad_kernel<<<(unsigned)(((size_t)m * h->ldA + 255) / 256), 256, 0, st>>>(vout.p, (size_t)n, Arand.p, h->ldA, n, (size_t)m);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(st));
    kin.release(); kout.release(); vin.release(); vout.release(); tmp.release();

    swne_nmf_params_t p;
    swne_nmf_default_params(&p, 0);
    p.loss = SWNE_LOSS_MSE; p.max_iter = max_iter; p.engine = engine; p.full_traces = -1;
    float *const dA_orig = h->dA;
    const bool own_orig = h->own_A;
    int rc_all = SWNE_OK;
    try {
        for (int which = 0; which < 2; ++which) {
            if (which == 1) {
                // the permutation keeps every statistic of the matrix (sum, sum of squares, nnz, max): only the data pointer changes
                h->dA = Arand.p;
                h->own_A = false;
                h->tc.ready = false; h->sp.ready = false;
            }
            for (int c = 0; c < nk; ++c) {
                p.k = k_range[c];
                swne_nmf_result_t res;
                const int rc = run_nmf_impl(h, &p, SWNE_INIT_RANDOM, seed + 7919ull * (uint64_t)(2 * k_range[c] + which), nullptr, nullptr, nullptr, nullptr,
                                            nullptr, nullptr, 0, &res);
                if (rc < 0) { rc_all = rc; throw ArgError{rc}; }
                h->red.reserve(8);
                CUDA_CHECK(cudaMemsetAsync(h->red.p, 0, 2 * sizeof(double), st));
                ops_of(h->last_KPv).recon_error(h->dA, h->ldA, n, m, h->Wt.p, h->H.p, p.k, h->red.p, std::min(m, h->sm_count * 8), st);
                CUDA_CHECK(cudaGetLastError());
                double r2[2];
                CUDA_CHECK(cudaMemcpyAsync(r2, h->red.p, sizeof(r2), cudaMemcpyDeviceToHost, st));
                CUDA_CHECK(cudaStreamSynchronize(st));
                k_err[(size_t)which * nk + c] = r2[loss_metric == SWNE_LOSS_MSE ? 0 : 1] / ((double)n * m);
            }
        }
    } catch (...) {
        h->dA = dA_orig; h->own_A = own_orig; h->tc.ready = false; h->sp.ready = false;
        Arand.release();
        throw;
    }
    h->dA = dA_orig; h->own_A = own_orig; h->tc.ready = false; h->sp.ready = false;
    Arand.release();
    // ---- the pick (R/nmf.R:186-194): first k where the permuted data gain more from the next factor than the real data
    int pick = -1;
    double best = 1e300;
    int best_i = 0;
    for (int c = 0; c + 1 < nk; ++c) {
        const double d = (k_err[c] - k_err[c + 1]) - (k_err[nk + c] - k_err[nk + c + 1]);
        if (err_diff) err_diff[c] = d;
        if (d < 0 && pick < 0) pick = c;
        if (d < best) { best = d; best_i = c; }
    }
    if (pick < 0) pick = best_i + 1;
    if (k_pick) *k_pick = k_range[pick];
    return rc_all;
}

// ------------------------------------------------------------------------------------------------
// consumers of H
// ------------------------------------------------------------------------------------------------
__global__ void gram64_kernel(const double *__restrict__ H, int k, int m, double *__restrict__ G /* k*k + k */)
{
    // one block per pair chunk: thread (a,b) strides over cells; small k, setup-time kernel
    const int pr = blockIdx.x;  // pair index or k*k + a for column sums
    const int a = pr / k, b = pr % k;
    __shared__ double scratch[32];
    double s = 0.0;
    if (pr < k * k) {
        for (int j = threadIdx.x; j < m; j += blockDim.x) s += H[(size_t)j * k + a] * H[(size_t)j * k + b];
    } else {
        const int c = pr - k * k;
        for (int j = threadIdx.x; j < m; j += blockDim.x) s += H[(size_t)j * k + c];
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) G[pr] = s;
}

// H == NULL: the H the most recent fit left on this handle's device (fp32 [cell][KP]); widened to the kernels' fp64 [cell][k]
// layout on the device -- the same values the host copy of that H holds, so the results are identical, without the upload
static void resident_h_to_f64(swne_nmf_handle_s *h, int k, int m, double *dH)
{
    if (h->last_k == 0 || !h->H.p) FAIL(SWNE_ERR_BAD_ARG, "H is NULL and no fit has left its H on the device (or another call has reused the buffer since)");
    if (k != h->last_k || m != h->m)
        FAIL(SWNE_ERR_BAD_ARG, "H is NULL: k and m must be those of the resident fit (k = %d, m = %d)", h->last_k, h->m);
    const size_t tot = (size_t)m * k;
    factor_out_cellmajor_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->H.p, dH, k, h->last_KPv, (size_t)m);
    CUDA_CHECK(cudaGetLastError());
}

static int sample_coords_impl(swne_nmf_handle_s *h, const double *H, int k, int m, const double *coords, double alpha,
                              int n_pull, double *out)
{
    if (!coords || !out) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    if (k < 1 || k > 64 || m < 1) FAIL(SWNE_ERR_BAD_ARG, "bad k or m");
    // R/swne_plotting.R:71-72
    if (n_pull > 0 && n_pull < 3) n_pull = 3;
    if (n_pull <= 0 || n_pull > k) n_pull = k;
    cudaStream_t st = h->stream;
    DevBuf<double> dH, dC, dO;
    dH.reserve((size_t)k * m); dC.reserve(2 * k); dO.reserve(2 * (size_t)m);
    if (H) CUDA_CHECK(cudaMemcpyAsync(dH.p, H, sizeof(double) * (size_t)k * m, cudaMemcpyHostToDevice, st));
    else resident_h_to_f64(h, k, m, dH.p);
    CUDA_CHECK(cudaMemcpyAsync(dC.p, coords, sizeof(double) * 2 * k, cudaMemcpyHostToDevice, st));
    sample_coords_kernel<<<(m + 127) / 128, 128, 2 * k * sizeof(double), st>>>(dH.p, k, m, dC.p, alpha, n_pull, dO.p);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(out, dO.p, sizeof(double) * 2 * (size_t)m, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    dH.release(); dC.release(); dO.release();
    return SWNE_OK;
}

static int factor_distance_impl(swne_nmf_handle_s *h, const double *H, int k, int m, int distance, double *out)
{
    if (!out) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    if (k < 1 || k > 64 || m < 1) FAIL(SWNE_ERR_BAD_ARG, "bad k or m");
    if (distance < 0 || distance > 2) FAIL(SWNE_ERR_BAD_ARG, "distance must be 0 cosine, 1 pearson or 2 euclidean");
    // the distance runs over ALL cells: a rank of a sharded fit only holds its own (gather H on the host and pass it)
    if (!H && h->nranks > 1) FAIL(SWNE_ERR_UNSUPPORTED, "factor distance from the resident H needs all cells on one GPU");
    cudaStream_t st = h->stream;
    DevBuf<double> dH, dG;
    dH.reserve((size_t)k * m); dG.reserve(k * k + k);
    if (H) CUDA_CHECK(cudaMemcpyAsync(dH.p, H, sizeof(double) * (size_t)k * m, cudaMemcpyHostToDevice, st));
    else resident_h_to_f64(h, k, m, dH.p);
    gram64_kernel<<<k * k + k, 256, 0, st>>>(dH.p, k, m, dG.p);
    CUDA_CHECK(cudaGetLastError());
    std::vector<double> G(k * k + k);
    CUDA_CHECK(cudaMemcpyAsync(G.data(), dG.p, sizeof(double) * G.size(), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    dH.release(); dG.release();
    for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b) {
            const double gab = G[a * k + b], gaa = G[a * k + a], gbb = G[b * k + b];
            double d;
            if (distance == 0) {
                d = 1.0 - gab / sqrt(gaa * gbb);
            } else if (distance == 1) {
                const double ma = G[k * k + a] / m, mb = G[k * k + b] / m;
                const double cov = gab - m * ma * mb, va = gaa - m * ma * ma, vb = gbb - m * mb * mb;
                d = sqrt(std::max(0.0, 2.0 * (1.0 - cov / sqrt(va * vb))));
            } else {
                d = sqrt(std::max(0.0, gaa + gbb - 2.0 * gab));
            }
            out[a + (size_t)k * b] = d;
        }
    return SWNE_OK;
}

static int recon_error_impl(swne_nmf_handle_s *h, const double *W, const double *H, int k, double *mse, double *kl)
{
    h->last_k = 0;     // h->H is about to be reused
    if (!h->dA || h->n == 0) FAIL(SWNE_ERR_NO_MATRIX, "no matrix resident");
    if (!W || !H) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    if (k < 1 || k > SWNE_MAX_K) FAIL(SWNE_ERR_BAD_ARG, "k out of range");
    const int n = h->n, m = h->m, KPv = kp_of(k);
    cudaStream_t st = h->stream;
    h->Wt.reserve((size_t)n * KPv); h->H.reserve((size_t)m * KPv); h->red.reserve(8);
    upload_factor_colmajor(h, W, h->Wt.p, k, KPv, n);
    CUDA_CHECK(cudaStreamSynchronize(st));
    upload_factor_cellmajor(h, H, h->H.p, k, KPv, (size_t)m);
    CUDA_CHECK(cudaMemsetAsync(h->red.p, 0, 2 * sizeof(double), st));
    const int grid = std::min(m, h->sm_count * 8);
    ops_of(KPv).recon_error(h->dA, h->ldA, n, m, h->Wt.p, h->H.p, k, h->red.p, grid, st);
    CUDA_CHECK(cudaGetLastError());
    if (h->nranks > 1) allreduce(h, nullptr, 0, h->red.p, 2);
    double r[2];
    CUDA_CHECK(cudaMemcpyAsync(r, h->red.p, sizeof(r), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    double mg = (double)m;
    if (h->nranks > 1) {
        double gl = (double)m;
        CUDA_CHECK(cudaMemcpyAsync(h->red.p, &gl, sizeof(double), cudaMemcpyHostToDevice, st));
        allreduce(h, nullptr, 0, h->red.p, 1);
        CUDA_CHECK(cudaMemcpyAsync(&mg, h->red.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
    const double N = (double)n * mg;
    if (mse) *mse = r[0] / N;
    if (kl) *kl = r[1] / N;
    return SWNE_OK;
}

}  // namespace swne

// ==================================================================================================
// extern "C"
// ==================================================================================================
using namespace swne;

#define API_BEGIN try {
#define API_END                                            \
    }                                                      \
    catch (const swne::ArgError &e) { return e.code; }     \
    catch (const swne::CudaError &) { return SWNE_ERR_CUDA; } \
    catch (const std::bad_alloc &) { swne::set_error("host allocation failed"); return SWNE_ERR_CUDA; } \
    catch (...) { swne::set_error("unknown C++ exception"); return SWNE_ERR_CUDA; }

extern "C" {

const char *swne_nmf_last_error(void) { return g_err; }
int swne_nmf_abi_version(void) { return SWNE_NMF_ABI_VERSION; }

int swne_nmf_device_count(int *count)
{
    API_BEGIN
    if (!count) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    CUDA_CHECK(cudaGetDeviceCount(count));
    return SWNE_OK;
    API_END
}

void swne_nmf_default_params(swne_nmf_params_t *p, int k)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->k = k;
    p->loss = SWNE_LOSS_MSE;
    p->max_iter = 500;
    p->rel_tol = 1e-4;
    p->inner_max_iter = 0;
    p->inner_rel_tol = 1e-9;
    p->trace = 0;
    p->verbose = 0;
    p->engine = SWNE_ENGINE_AUTO;
    p->full_traces = 0;
    p->profile = 0;
    p->nnlm_variant = 0;
}

int swne_nmf_create(swne_nmf_handle_t *out, int device)
{
    API_BEGIN
    if (!out) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    int cnt = 0;
    CUDA_CHECK(cudaGetDeviceCount(&cnt));
    if (cnt < 1) FAIL(SWNE_ERR_CUDA, "no CUDA device visible: libswne_nmf has no CPU fallback");
    if (device < 0 || device >= cnt) FAIL(SWNE_ERR_BAD_ARG, "device %d out of range (0..%d)", device, cnt - 1);
    DeviceGuard g(device);
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) FAIL(SWNE_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    swne_nmf_handle_s *h = new swne_nmf_handle_s();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    // aux[0] and aux[2] carry the HBM-bound contractions of the slab pipeline at high priority, aux[1] the latency-bound solves at
    // low priority: when an SM frees resources the block scheduler places a waiting contraction CTA before more solve CTAs
    int prio_least = 0, prio_greatest = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    for (int q = 0; q < 3; ++q) {
        CUDA_CHECK(cudaStreamCreateWithPriority(&h->aux[q], cudaStreamNonBlocking, q == 1 ? prio_least : prio_greatest));
        CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_done[q], cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    *out = h;
    return SWNE_OK;
    API_END
}

int swne_nmf_destroy(swne_nmf_handle_t h)
{
    API_BEGIN
    if (!h) return SWNE_OK;
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->stream);
    xchg_release(h, true);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    h->A_store.release();
    h->sp.cptr.release(); h->sp.cidx.release(); h->sp.gptr.release(); h->sp.gidx.release();
    h->sp.cval.release(); h->sp.gval.release(); h->sp.ajt.release();
    tc_release(h->tc);
    tcx_release(h->tcx);
    h->Wt.release(); h->H.release(); h->C.release(); h->B.release(); h->G.release(); h->pack.release();
    h->G64w.release(); h->G64h.release(); h->stage.release(); h->dstats.release(); h->dscal.release(); h->red.release(); h->flags.release();
    if (h->pinned) cudaFreeHost(h->pinned);
    for (int q = 0; q < 3; ++q) {
        if (h->aux[q]) { cudaStreamSynchronize(h->aux[q]); cudaStreamDestroy(h->aux[q]); }
        if (h->ev_done[q]) cudaEventDestroy(h->ev_done[q]);
        if (q == 0) for (cudaEvent_t e : h->ev_slab) cudaEventDestroy(e);
    }
    if (h->ev_ready) cudaEventDestroy(h->ev_ready);
    cudaStreamDestroy(h->stream);
    delete h;
    return SWNE_OK;
    API_END
}

int swne_nmf_set_progress(swne_nmf_handle_t h, swne_nmf_progress_fn fn, void *user)
{
    if (!h) { set_error("NULL handle"); return SWNE_ERR_BAD_ARG; }
    h->progress = fn; h->progress_user = user;
    return SWNE_OK;
}

int swne_nmf_comm_unique_id(void *id128)
{
    API_BEGIN
    if (!id128) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    load_nccl();
    ncclUniqueId id;
    NCCL_CHECK(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return SWNE_OK;
    API_END
}

int swne_nmf_comm_init(swne_nmf_handle_t h, const void *id128, int rank, int nranks)
{
    API_BEGIN
    if (!h || !id128) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) FAIL(SWNE_ERR_BAD_ARG, "bad rank/nranks");
    DeviceGuard g(h->device);
    if (nranks == 1) { h->rank = 0; h->nranks = 1; return SWNE_OK; }
    load_nccl();
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NCCL_CHECK(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    h->rank = rank; h->nranks = nranks;
    return SWNE_OK;
    API_END
}

int swne_nmf_comm_info(swne_nmf_handle_t h, int *rank, int *nranks, int *peer_exchange)
{
    API_BEGIN
    if (!h) FAIL(SWNE_ERR_BAD_ARG, "NULL handle");
    if (rank) *rank = h->rank;
    if (nranks) *nranks = h->nranks;
    if (peer_exchange) *peer_exchange = h->xchg.ready ? 1 : 0;
    return SWNE_OK;
    API_END
}

// The solvers read the Gram's triangular coefficients from __constant__ memory, which is per device, not per handle:
// compute entry points of handles on the same device are serialised (different devices run concurrently).
static std::mutex &device_mutex(int device)
{
    static std::mutex m[64];
    return m[device & 63];
}
#define HANDLE_GUARD(h)                                               \
    if (!(h)) FAIL(SWNE_ERR_BAD_ARG, "NULL handle");                  \
    std::lock_guard<std::mutex> _dev_lock(device_mutex((h)->device)); \
    DeviceGuard _guard((h)->device)

int swne_nmf_set_matrix_dense_f64(swne_nmf_handle_t h, const double *A, int n, int m)
{
    API_BEGIN
    HANDLE_GUARD(h);
    set_dense<double>(h, A, n, m);
    return SWNE_OK;
    API_END
}
int swne_nmf_set_matrix_dense_f32(swne_nmf_handle_t h, const float *A, int n, int m)
{
    API_BEGIN
    HANDLE_GUARD(h);
    set_dense<float>(h, A, n, m);
    return SWNE_OK;
    API_END
}
int swne_nmf_set_matrix_csc_f64(swne_nmf_handle_t h, const int *p, const int *i, const double *x, int n, int m)
{
    API_BEGIN
    HANDLE_GUARD(h);
    set_csc<double>(h, p, i, x, n, m);
    return SWNE_OK;
    API_END
}
int swne_nmf_set_matrix_csc_f32(swne_nmf_handle_t h, const int *p, const int *i, const float *x, int n, int m)
{
    API_BEGIN
    HANDLE_GUARD(h);
    set_csc<float>(h, p, i, x, n, m);
    return SWNE_OK;
    API_END
}
int swne_nmf_set_matrix_csc_counts(swne_nmf_handle_t h, const int *p, const int *i, const double *x, int n, int m,
                                   const double *cell_scale, const double *gene_scale, int transform)
{
    API_BEGIN
    HANDLE_GUARD(h);
    set_csc<double>(h, p, i, x, n, m, cell_scale, gene_scale, transform);
    return SWNE_OK;
    API_END
}
int swne_nmf_set_matrix_dense_f32_device(swne_nmf_handle_t h, const float *dA, int n, int m)
{
    API_BEGIN
    HANDLE_GUARD(h);
    if (!dA) FAIL(SWNE_ERR_BAD_ARG, "A is NULL");
    cudaPointerAttributes attr;
    CUDA_CHECK(cudaPointerGetAttributes(&attr, dA));
    if (attr.type != cudaMemoryTypeDevice) FAIL(SWNE_ERR_BAD_ARG, "A is not a device pointer");
    if (n % 4 == 0 && ((uintptr_t)dA & 15) == 0) {
        begin_matrix(h, n, m, false);
        h->A_store.release();
        h->dA = const_cast<float *>(dA);   // borrowed: the caller keeps it alive and unchanged
        h->own_A = false;
    } else {
        begin_matrix(h, n, m, true);
        CUDA_CHECK(cudaMemsetAsync(h->dA, 0, (size_t)m * h->ldA * sizeof(float), h->stream));
        CUDA_CHECK(cudaMemcpy2DAsync(h->dA, h->ldA * sizeof(float), dA, (size_t)n * sizeof(float), (size_t)n * sizeof(float), m,
                                     cudaMemcpyDeviceToDevice, h->stream));
    }
    const int grid = std::min(m, h->sm_count * 8);
    ingest_dense_kernel<float><<<grid, 256, 0, h->stream>>>(h->dA, h->ldA, nullptr, h->ldA, n, m, h->dstats.p);
    CUDA_CHECK(cudaGetLastError());
    check_stats(h);
    return SWNE_OK;
    API_END
}

int swne_nmf_matrix_info(swne_nmf_handle_t h, int *n, int *m, long long *nnz, double *sum, double *sumsq)
{
    API_BEGIN
    if (!h) FAIL(SWNE_ERR_BAD_ARG, "NULL handle");
    if (!h->dA || h->n == 0) FAIL(SWNE_ERR_NO_MATRIX, "no matrix resident");
    if (n) *n = h->n;
    if (m) *m = h->m;
    if (nnz) *nnz = (long long)h->stats.nnz;
    if (sum) *sum = h->stats.sum;
    if (sumsq) *sumsq = h->stats.sumsq;
    return SWNE_OK;
    API_END
}

int swne_nmf_fit(swne_nmf_handle_t h, const swne_nmf_params_t *p, double *W, double *H, double *mse, double *mkl,
                 double *target_loss, double *average_epochs, int trace_len, swne_nmf_result_t *res)
{
    API_BEGIN
    HANDLE_GUARD(h);
    swne_nmf_result_t local;
    return fit_impl(h, p, W, H, mse, mkl, target_loss, average_epochs, trace_len, res ? res : &local);
    API_END
}

int swne_nmf_fit_dense_f64(swne_nmf_handle_t h, const swne_nmf_params_t *p, const double *A, int n, int m, double *W,
                           double *H, double *mse, double *mkl, double *target_loss, double *average_epochs,
                           int trace_len, swne_nmf_result_t *res)
{
    API_BEGIN
    HANDLE_GUARD(h);
    FitEvents ev;       // destroyed on every exit path
    cudaEventRecord(ev.e[0], h->stream);
    set_dense<double>(h, A, n, m);
    cudaEventRecord(ev.e[1], h->stream);
    swne_nmf_result_t local;
    swne_nmf_result_t *r = res ? res : &local;
    const int rc = fit_impl(h, p, W, H, mse, mkl, target_loss, average_epochs, trace_len, r);
    float ms = 0;
    cudaEventElapsedTime(&ms, ev.e[0], ev.e[1]);
    r->upload_ms += ms;
    return rc;
    API_END
}

int swne_nmf_fit_csc_f64(swne_nmf_handle_t h, const swne_nmf_params_t *p, const int *cp, const int *ci, const double *cx,
                         int n, int m, double *W, double *H, double *mse, double *mkl, double *target_loss,
                         double *average_epochs, int trace_len, swne_nmf_result_t *res)
{
    API_BEGIN
    HANDLE_GUARD(h);
    set_csc<double>(h, cp, ci, cx, n, m);
    swne_nmf_result_t local;
    return fit_impl(h, p, W, H, mse, mkl, target_loss, average_epochs, trace_len, res ? res : &local);
    API_END
}

int swne_run_nmf(swne_nmf_handle_t h, const swne_nmf_params_t *p, int init, uint64_t seed, double *W, double *H, double *mse,
                 double *mkl, double *target_loss, double *average_epochs, int trace_len, swne_nmf_result_t *res)
{
    API_BEGIN
    HANDLE_GUARD(h);
    swne_nmf_result_t local;
    return run_nmf_impl(h, p, init, seed, W, H, mse, mkl, target_loss, average_epochs, trace_len, res ? res : &local);
    API_END
}

int swne_find_num_factors(swne_nmf_handle_t h, const int *k_range, int nk, int loss_metric, int max_iter, uint64_t seed, int engine,
                          double *k_err, int *k_pick, double *err_diff)
{
    API_BEGIN
    HANDLE_GUARD(h);
    return find_num_factors_impl(h, k_range, nk, loss_metric, max_iter, seed, engine, k_err, k_pick, err_diff);
    API_END
}

int swne_nnlm(swne_nmf_handle_t h, int side, int k, const double *X, double *coef, const double *alpha3, int loss,
              int max_iter, double rel_tol, long long *sweeps)
{
    API_BEGIN
    HANDLE_GUARD(h);
    return nnlm_impl(h, side, k, X, coef, alpha3, loss, max_iter, rel_tol, sweeps);
    API_END
}

int swne_nnsvd_init(swne_nmf_handle_t h, int k, int oversample, int power_iters, uint64_t seed, double *W, double *H,
                    double *singular_values)
{
    API_BEGIN
    HANDLE_GUARD(h);
    return nnsvd_impl(h, k, oversample, power_iters, seed, W, H, singular_values);
    API_END
}

int swne_ica_init(swne_nmf_handle_t h, int k, int ica_fast, uint64_t seed, double *W, double *H)
{
    API_BEGIN
    HANDLE_GUARD(h);
    return ica_init_impl(h, k, ica_fast, seed, W, H);
    API_END
}

int swne_zero_replace(double *x, size_t len, double eps, const double *fill, size_t n_fill, size_t *n_zero)
{
    API_BEGIN
    if (!x && len) FAIL(SWNE_ERR_BAD_ARG, "NULL argument");
    // branch-free passes: about half of an NNDSVD seed is zero, in no predictable pattern
    size_t z = 0;
    for (size_t i = 0; i < len; ++i) {
        const double v = (x[i] < eps) ? 0.0 : x[i];
        x[i] = v;
        z += (v == 0.0) ? 1u : 0u;
    }
    if (fill) {
        if (z > n_fill) FAIL(SWNE_ERR_BAD_ARG, "fill vector shorter than the number of zeros");
        if (z) {
            size_t q = 0;
            const size_t last = z - 1;
            for (size_t i = 0; i < len; ++i) {
                const double v = x[i];
                const bool isz = (v == 0.0);
                x[i] = isz ? fill[q < last ? q : last] : v;
                q += isz ? 1u : 0u;
            }
        }
    }
    if (n_zero) *n_zero = z;
    return SWNE_OK;
    API_END
}

int swne_sample_coords(swne_nmf_handle_t h, const double *H, int k, int m, const double *H_coords, double alpha_exp,
                       int n_pull, double *out)
{
    API_BEGIN
    HANDLE_GUARD(h);
    return sample_coords_impl(h, H, k, m, H_coords, alpha_exp, n_pull, out);
    API_END
}

int swne_factor_distance(swne_nmf_handle_t h, const double *H, int k, int m, int distance, double *out)
{
    API_BEGIN
    HANDLE_GUARD(h);
    return factor_distance_impl(h, H, k, m, distance, out);
    API_END
}

int swne_recon_error(swne_nmf_handle_t h, const double *W, const double *H, int k, double *mse, double *kl)
{
    API_BEGIN
    HANDLE_GUARD(h);
    return recon_error_impl(h, W, H, k, mse, kl);
    API_END
}

}  // extern "C"
