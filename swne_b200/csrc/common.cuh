// common.cuh -- error handling and small device helpers shared by all kernels of libswne_nmf.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/swne_nmf.h"

#define TINY_NUM_F 1e-16f
#define TINY_NUM_D 1e-16

namespace swne {

void set_error(const char *fmt, ...);

struct CudaError {
    cudaError_t e;
    const char *file;
    int line;
};

#define CUDA_CHECK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            swne::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                            cudaGetErrorString(_e));                                              \
            throw swne::CudaError{_e, __FILE__, __LINE__};                                        \
        }                                                                                         \
    } while (0)

static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }
static inline size_t round_up_sz(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of one double per thread; result valid in thread 0.  blockDim.x <= 1024.
__device__ __forceinline__ double block_sum(double v, double *s_scratch /* >= 32 doubles */)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_scratch[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? s_scratch[threadIdx.x] : 0.0;
    if (wid == 0) v = warp_sum(v);
    return v;
}

struct MatStats {
    double sum, sumsq, klc;          // sum a, sum a^2, sum (a+tiny) log(a+tiny) - a
    unsigned long long nnz;
    unsigned int bad;                // bit0 negative, bit1 non finite, bit2 CSC index out of range
    unsigned int max_bits;           // float bits of the largest entry (entries are >= 0)
};

// device-side scalars of one fit; copied to the host at every error block
struct FitScalars {
    double loss_part;        // sum_j ( h_j' WW h_j - 2 h_j' c_j )   (mse via the Gram identity)
    double kl_part;          // sum_nz a * log(ahat + tiny)
    double sq_part;          // sum_nz ( a^2 - 2 a ahat )            (mse of an mkl fit)
    unsigned long long sweeps_w;
    unsigned long long sweeps_h;
    double pad[3];
};

}  // namespace swne
