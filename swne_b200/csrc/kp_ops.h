// kp_ops.h -- launch table of the kernels that are templated on KP (k padded to a multiple of 8).
// Each KP is compiled in its own translation unit (kernels_kp.cu with -DKP_VALUE=...) so the build
// parallelises over cores; swne_nmf.cu picks the table with kp_ops(KP).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace swne {

struct FitScalars;

struct KpOps {
    void (*gram)(const float *X, int rows, double *G64, double *colsum64, cudaStream_t st);
    void (*contract_cells)(const float *A, size_t ldA, int n, int m, const float *Wt, float *C, cudaStream_t st);
    void (*contract_genes)(const float *A, size_t ldA, int n, int m, const float *H, float *B, int cells_per_chunk, int chunks,
                           cudaStream_t st);
    // pack: GramPack of this side (see kernels_simt.cuh); its Grams are copied into the constant bank first
    void (*solve)(float *X, const float *C, const float *pack, float l1, int k, int cols, int max_iter, float rel_tol,
                  unsigned long long *sweeps, double *loss_part, cudaStream_t st);
    void (*kl)(const int *ptr, const int *idx, const float *val, float *ajt, float *X, const float *Y, const float *sumY,
               float b0, float b1, float b2, int k, int cols, int max_iter, float rel_tol, unsigned long long *sweeps,
               double *kl_part, double *sq_part, int err, int max_nnz, int avg_nnz, const int *order, const int *sorted_cnt_host,
               const float *Yt, size_t ld, cudaStream_t st);   // Yt: factor-major copy of Y (kl_shard_transpose) or null
    // gene side of the KL update with the cells sharded over ranks (kernels_simt.cuh, kl_shard_*): init of A-hat and the
    // per-gene state, the rank's partial Newton sums of coordinate kk (part: float2 per gene), the replicated step, sweep end
    void (*kl_shard_init)(const int *ptr, const int *idx, float *ajt, const float *X, const float *Y, int k, int cols, float *dprev,
                          float *sumX, int *active, cudaStream_t st);
    void (*kl_shard_transpose)(const float *Y, int rows, float *Yt, size_t ld, cudaStream_t st);   // Yt[f][r] = Y[r][f]
    void (*kl_shard_accum)(const int *ptr, const int *idx, const float *val, float *ajt, const float *Yt, size_t ld,
                           const float *dprev, const int *active, int kk, int kprev, int cols, float *part, cudaStream_t st);
    void (*kl_shard_update)(const float *part, float *X, const float *sumY, float *sumX, float *dprev, int *active, float b0,
                            float b1, float b2, int kk, int cols, float rel_tol, int flags, cudaStream_t st);
    void (*kl_shard_sweep_end)(int *active, int cols, unsigned long long *sweeps, unsigned *n_active, cudaStream_t st);
    void (*kl_trace)(const float *A, size_t ldA, int n, int m, const float *H, const float *Wt, int k, double *kl_part,
                     cudaStream_t st);
    void (*recon_error)(const float *A, size_t ldA, int n, int m, const float *Wt, const float *H, int k, double *out2, int grid,
                        cudaStream_t st);
    void (*right_multiply)(float *X, size_t rows, const float *T, int l, cudaStream_t st);
    // one FastICA pass over the whitened data (ica_step_kernel): out[KP*KP + KP] doubles, zeroed by the caller
    void (*ica_step)(const float *Y, int rows, int nc, const float *R, double *out, int grid, cudaStream_t st);
};

const KpOps *kp_ops(int KP);  // nullptr when KP is not one of 8,16,...,64

}  // namespace swne
