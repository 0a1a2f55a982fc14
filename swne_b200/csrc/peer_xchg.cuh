// peer_xchg.cuh -- the per-iteration exchange of the cell-sharded fit over NVLink peer memory:
//   A H' partial (n x KP fp32)  +  H H' and row sums (fp64)  +  loss partial  ->  summed over the ranks, on every rank
// (fused engine: KP = 32, plus the GramPack of the next W update; wide engine's slab pipeline: KP = k rounded up to 8)
// as a one-shot all-reduce written for this payload (a few hundred KB, <= 8 ranks on one NVSwitch box), instead of three
// grouped ncclAllReduce calls plus a finalize kernel:
//   push    every rank stores its payload into slot [parity][rank] of EVERY rank's exchange region (its own included) with
//           16-byte stores through the peer mappings, fences at system scope, and its last CTA publishes the epoch number in
//           flag [parity][rank] of every region;
//   reduce  every rank waits until the flags of all ranks in ITS region show the epoch, sums the slots in rank order (so all
//           ranks get bit-identical sums, hence bit-identical W), writes B, the fp64 sums and the loss back in place, and its
//           last CTA builds the GramPack of the next W update (what finalize_gram_kernel did in a launch of its own).
// Regions are double-buffered by epoch parity: a rank can run at most one exchange ahead of the slowest one (its next push
// needs the slowest rank's flag of the current epoch first), so two buffers suffice.  The regions are plain cudaMalloc
// memory shared through CUDA IPC handles, which the ranks swap with one ncclAllGather at set-up; NCCL stays in charge of
// the rare collectives (statistics, stopping decisions, traces).
// A rank that never shows up would leave the others spinning: the wait gives up after 5 s of globaltimer, raises a flag
// the host turns into SWNE_ERR_NCCL at the next error block, and lets the kernel finish.
#pragma once
#include <vector>

#include "common.cuh"
#include "pass_tc.cuh"

namespace swne {

struct PeerXchg {
    bool ready = false;
    bool tried = false;               // set-up attempted for the current payload size (a failed attempt is not repeated)
    int nranks = 0, rank = 0;
    size_t payload = 0;               // bytes per slot (capacity: the largest payload of this matrix), multiple of 16
    size_t flags_off = 0, total = 0;  // flags: [2][nranks] uint64 behind the slots; then the two ticket / error ints
    uint8_t *region = nullptr;        // this rank's region
    uint8_t **peers_dev = nullptr;    // device array [nranks]: base of every rank's region as mapped here
    std::vector<void *> opened;       // peer mappings (to close)
    unsigned long long epoch = 0;
};

constexpr int XCHG_THREADS = 256;

// payload layout of a slot: [nB4 float4 of B][nG2 double2 of the fp64 sums][one double2: loss partial, 0]
__global__ void __launch_bounds__(XCHG_THREADS) xchg_push_kernel(uint8_t *const *__restrict__ peers, int nranks, int rank, size_t payload,
                                                                 int parity, unsigned long long epoch, size_t flags_off,
                                                                 const float4 *__restrict__ B, size_t nB4, const double2 *__restrict__ G,
                                                                 size_t nG2, const double *__restrict__ loss, int *ticket)
{
    const size_t slot_off = ((size_t)parity * nranks + rank) * payload;
    const size_t total16 = nB4 + nG2 + 1;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total16; idx += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        if (idx < nB4) {
            const float4 f = __ldcg(B + idx);          // written by the pass kernel's red.add (L2)
            v = make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w));
        } else if (idx < nB4 + nG2) {
            const double2 d = __ldcg(G + (idx - nB4));
            v = make_uint4((unsigned)__double2loint(d.x), (unsigned)__double2hiint(d.x), (unsigned)__double2loint(d.y),
                           (unsigned)__double2hiint(d.y));
        } else {
            const double l = loss ? __ldcg(loss) : 0.0;
            v = make_uint4((unsigned)__double2loint(l), (unsigned)__double2hiint(l), 0u, 0u);
        }
        for (int p = 0; p < nranks; ++p) {
            // start with the next rank so that the ranks do not all hit the same destination at once
            const int q = (rank + 1 + p) % nranks;
            reinterpret_cast<uint4 *>(peers[q] + slot_off)[idx] = v;
        }
    }
    __threadfence_system();
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        s_last = (t == (int)gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        // every CTA fenced its stores before taking a ticket: publish the epoch to all ranks
        __threadfence_system();
        if (threadIdx.x < nranks) {
            unsigned long long *flag =
                reinterpret_cast<unsigned long long *>(peers[threadIdx.x] + flags_off) + (size_t)parity * nranks + rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
        }
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// sums the slots of this rank's region in rank order into B / G / loss, then (last CTA) the GramPack of G.
// err: int raised when a peer's flag did not arrive in time.  The dynamic shared memory is not used; s_inv[32] is static.
__global__ void __launch_bounds__(XCHG_THREADS) xchg_reduce_kernel(const uint8_t *__restrict__ region, int nranks, size_t payload, int parity,
                                                                   unsigned long long epoch, size_t flags_off, float4 *__restrict__ B,
                                                                   size_t nB4, double2 *__restrict__ G, size_t nG2, double *__restrict__ loss,
                                                                   int *ticket, int *err, float *pack_next, int k, double pen0, double pen1)
{
    __shared__ int s_flag;
    __shared__ float s_inv[TC_KP];
    if (threadIdx.x < nranks) {
        const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(region + flags_off) + (size_t)parity * nranks + threadIdx.x;
        const unsigned long long t0 = gtimer();
        unsigned long long seen = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
            if (seen >= epoch) break;
            if (gtimer() - t0 > 5000000000ull) { atomicOr(err, 1); break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    const uint8_t *base = region + (size_t)parity * nranks * payload;
    const size_t total16 = nB4 + nG2 + 1;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total16; idx += (size_t)gridDim.x * blockDim.x) {
        if (idx < nB4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < nranks; ++r) {
                const float4 v = __ldcg(reinterpret_cast<const float4 *>(base + (size_t)r * payload) + idx);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            B[idx] = acc;
        } else {
            double2 acc = make_double2(0.0, 0.0);
            for (int r = 0; r < nranks; ++r) {
                const double2 v = __ldcg(reinterpret_cast<const double2 *>(base + (size_t)r * payload) + idx);
                acc.x += v.x; acc.y += v.y;
            }
            if (idx < nB4 + nG2) G[idx - nB4] = acc;
            else if (loss) *loss = acc.x;
        }
    }
    // (pack_next == nullptr: the wide engine, whose W update builds its GramPack itself at another row stride)
    if (pack_next != nullptr && tc_last_cta(ticket, &s_flag))
        tc_finalize_gram(reinterpret_cast<const double *>(G), pack_next, k, pen0, pen1, s_inv);
}

// host side ------------------------------------------------------------------------------------------------------------
static inline size_t xchg_payload_bytes(size_t nB_floats, size_t nG_doubles) { return nB_floats * 4 + nG_doubles * 8 + 16; }

static inline void xchg_launch(PeerXchg &x, float *B, size_t nB_floats, double *G, size_t nG_doubles, double *loss, float *pack_next, int k,
                               double pen0, double pen1, int sm_count, cudaStream_t st)
{
    const unsigned long long epoch = ++x.epoch;
    const int parity = (int)(epoch & 1);
    int *ticket = reinterpret_cast<int *>(x.region + x.flags_off + 2 * (size_t)x.nranks * 8);
    const size_t nB4 = nB_floats / 4, nG2 = nG_doubles / 2;
    const int grid = (int)std::min<size_t>((size_t)sm_count, (nB4 + nG2 + XCHG_THREADS) / XCHG_THREADS);
    xchg_push_kernel<<<grid, XCHG_THREADS, 0, st>>>(x.peers_dev, x.nranks, x.rank, x.payload, parity, epoch, x.flags_off,
                                                    reinterpret_cast<const float4 *>(B), nB4, reinterpret_cast<const double2 *>(G), nG2, loss,
                                                    ticket);
    xchg_reduce_kernel<<<grid, XCHG_THREADS, 0, st>>>(x.region, x.nranks, x.payload, parity, epoch, x.flags_off, reinterpret_cast<float4 *>(B),
                                                      nB4, reinterpret_cast<double2 *>(G), nG2, loss, ticket + 1, ticket + 2, pack_next, k, pen0,
                                                      pen1);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace swne
