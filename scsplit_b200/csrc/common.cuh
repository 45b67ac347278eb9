// Shared declarations of the scsplit_b200 device library (internal; the public surface is include/scsplit_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/scsplit_b200.h"

namespace scs {

void set_error(const char* fmt, ...);

#define SCS_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            scs::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

#define SCS_CHECK(cond, ...)              \
    do {                                  \
        if (!(cond)) {                    \
            scs::set_error(__VA_ARGS__);  \
            return 1;                     \
        }                                 \
    } while (0)

#define SCS_TRY(expr)        \
    do {                     \
        int _r = (expr);     \
        if (_r) return _r;   \
    } while (0)

// Entry code word of both orientations: (index << 2) | (sel << 1) | dup
//   sel = 0: alt-allele entry, 1: ref-allele entry
//   dup = 1: a ref entry whose (SNV, cell) also has an alt entry (skipped by pattern counts)
// cell-major ("E") stream: index = SNV; gather row of the log table = code >> 1 = 2*v + sel
// SNV-major  ("M") stream: index = cell; gather row of W          = code >> 2 = c
struct Stream {
    int64_t nrows = 0;     // E: N cells;  M: 2V pseudo-rows (2v = alt entries of SNV v, 2v+1 = ref entries)
    int64_t n_light = 0;   // entries with count == 1 (4 bytes each)
    int64_t n_heavy = 0;   // entries with count  > 1 (8 bytes each)
    int64_t* lptr = nullptr;   // [nrows+1]
    int64_t* hptr = nullptr;   // [nrows+1]
    uint32_t* lcode = nullptr; // [n_light]
    uint32_t* hcode = nullptr; // [n_heavy]
    int32_t* hcnt = nullptr;   // [n_heavy]
};

struct NcclApi;  // dlopen'ed function table (comm.cu)

struct EmState {
    int R = 0, K = 0, KP = 0, max_iter = 0;
    int64_t stats_stride = 0;   // doubles per restart in `stats`
    double* T = nullptr;        // [R][2V][KP]  log2 P (row 2v) and log2(1-P) (row 2v+1)
    double* P = nullptr;        // [R][V][KP]
    double* L = nullptr;        // [R][N][KP]
    double* W = nullptr;        // [R][N][KP]
    double* stats = nullptr;    // [R][2V*KP + KP + 2]: S (alt/ref posterior-weighted sums), colsum[KP], LL, pad
    double* lps = nullptr;      // [R][KP]
    double* partial = nullptr;  // [R][nblk][KP+1] per-block column mass + LL (fixed-order reduction)
    double* hist = nullptr;     // [R][max_iter]
    double* ll_prev = nullptr;  // [R]
    int32_t* iters = nullptr;   // [R]
    int32_t* done = nullptr;    // [R]
    int nblk = 0;
};

}  // namespace scs

struct scs_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    int64_t V = 0, N = 0;
    int64_t nnz_ref = 0, nnz_alt = 0, nnz_union = 0;
    int64_t device_bytes = 0;
    scs::Stream E;   // cell-major
    scs::Stream M;   // SNV-major (2V pseudo-rows)
    int64_t* sum_alt = nullptr;  // [V] integer row sums (global after comm_init)
    int64_t* sum_ref = nullptr;  // [V]
    double* kalt = nullptr;      // [V]
    scs::EmState em;
    // multi-GPU
    void* comm = nullptr;        // ncclComm_t
    int rank = 0, world = 1;
    // instrumentation
    int64_t launches = 0;
    int timing = 0;
    int variant = 0;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_pending;  // (kind, (start, stop))
    cudaEvent_t ev_iter[2] = {nullptr, nullptr};   // brackets the launches of the last scs_em_iterate
    double last_iter_ms = 0.0;
    double t_sum[2] = {0, 0};
    int64_t t_cnt[2] = {0, 0};
};

namespace scs {

template <typename T>
int dev_alloc(scs_ctx* ctx, T** p, int64_t n) {
    *p = nullptr;
    if (n <= 0) n = 1;
    SCS_CUDA(cudaMalloc((void**)p, (size_t)n * sizeof(T)));
    ctx->device_bytes += n * (int64_t)sizeof(T);
    return 0;
}
template <typename T>
void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

inline int round_even(int k) { return (k + 1) & ~1; }

// layout.cu
int build_layout(scs_ctx* ctx, const int64_t* ref_indptr, const int32_t* ref_indices, const void* ref_data,
                 const int64_t* alt_indptr, const int32_t* alt_indices, const void* alt_data, int data_bytes);
void free_stream(Stream& s);

// em.cu
int launch_spmm(scs_ctx* ctx, int kind /*0=E,1=M*/, int KP, int R, const double* X, int64_t x_stride, double* Y,
                int64_t y_stride, const int32_t* done);
int em_free(scs_ctx* ctx);

// comm.cu
int comm_allreduce_f64(scs_ctx* ctx, double* buf_dev, int64_t count);
int comm_allreduce_i64(scs_ctx* ctx, int64_t* buf_dev, int64_t count);
int comm_destroy(scs_ctx* ctx);

}  // namespace scs
