// pass_tc.cuh -- fused one-pass tcgen05 engine (SWNE_ENGINE_TC).  Placeholder until the kernel lands:
// tc_supported() is false so SWNE_ENGINE_AUTO resolves to the fp32 CUDA-core engine and an explicit
// SWNE_ENGINE_TC request fails loudly with SWNE_ERR_UNSUPPORTED.
#pragma once
#include "common.cuh"

namespace swne {

struct TcMatrix {
    bool ready = false;
};

static inline bool tc_supported(int, int, int) { return false; }
static inline void tc_prepare(TcMatrix &, const float *, size_t, int, int, int, int, cudaStream_t) {}
static inline void tc_pass(TcMatrix &, const float *, float *, const float *, const float *, float, int, int, float, float *,
                           double *, FitScalars *, cudaStream_t)
{
}
static inline int tc_pass_launches() { return 1; }
static inline void tc_release(TcMatrix &) {}

}  // namespace swne
