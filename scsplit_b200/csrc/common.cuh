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

// Device-side bounds checks of the gather / scatter indices (compute-sanitizer is not available on the pool's boxes): compiled
// in with -DSCS_BOUNDS_CHECK (python -m scsplit_b200.build --bounds -> libscsplit_b200_bounds.so, selected with
// SCSPLIT_B200_LIB); a violated check traps the kernel, which the next CUDA call reports.
#ifdef SCS_BOUNDS_CHECK
#include <assert.h>
#define SCS_ASSERT(cond) assert(cond)
#else
#define SCS_ASSERT(cond) ((void)0)
#endif

#define SCS_TRY(expr)        \
    do {                     \
        int _r = (expr);     \
        if (_r) return _r;   \
    } while (0)

// Entry code word of both orientations: (index << 1) | dup
//   dup = 1: a ref-allele entry whose (SNV, cell) also has an alt-allele entry (skipped by pattern counts)
// "sel-major" numbering: pseudo-row / gather row g = sel * V + v with sel = 0 (alt) or 1 (ref).
// cell-major ("E") stream: rows = N cells, index = g; gather row of the log table T[2V][KP] = code >> 1
//                          (T[v] = log2 P[v], T[V+v] = log2(1-P[v])); entries of a cell ascend in g.
// SNV-major  ("M") stream: rows = 2V pseudo-rows g (alt rows first), index = cell; gather row of W[N][KP] = code >> 1;
//                          entries of a pseudo-row ascend in cell.  S[g] = sum of W rows: S[v] = altW, S[V+v] = refW.
struct Stream {
    int64_t nrows = 0;
    int64_t xrows = 0;     // rows of the gathered operand (E: 2V, M: N)
    int64_t n_light = 0;   // entries with count == 1 (4 bytes each)
    int64_t n_heavy = 0;   // entries with count  > 1 (8 bytes each)
    int64_t* lptr = nullptr;   // [nrows+1]
    int64_t* hptr = nullptr;   // [nrows+1]
    uint32_t* lcode = nullptr; // [n_light]
    uint32_t* hcode = nullptr; // [n_heavy]
    int32_t* hcnt = nullptr;   // [n_heavy]
};

// Tile-major copy of a stream's count-1 entries for one table width KP (built on first use, DESIGN.md "v2"):
// the gathered operand is cut into tiles of TR rows that fit shared memory; unit u = tile * nrows + row owns the
// entries tcode[tptr[u] .. tptr[u+1]) of `row` that fall into `tile`, stored as 16-bit row offsets inside the tile.
struct TiledStream {
    int KP = 0;
    int TR = 0;            // operand rows per tile
    int ntiles = 0;
    int64_t nunits = 0;    // ntiles * nrows
    int64_t* tptr = nullptr;   // [nunits + 1]
    uint16_t* tcode = nullptr; // [n_light (+pad)]
    // static work plan of k_spmm_tiled (built once per KP): CTA j runs the segments seg_ptr[j]..seg_ptr[j+1]); a segment
    // is a run of units inside ONE tile; gb cuts it into `ngroups` contiguous runs of equal cost for the row groups
    int ncta = 0, ngroups = 0, nseg = 0;
    int32_t* seg_ptr = nullptr;   // [ncta + 1]
    int32_t* seg_tile = nullptr;  // [nseg]
    int64_t* gb = nullptr;        // [nseg][ngroups + 1] unit boundaries
};

// Lock-step stream of the fixed-point E-step (estep_lk.cuh): the units (cell, table tile) of a tile sorted by length and dealt
// 32 at a time to "tasks" -- 32 units of one padded length, one per lane -- whose codes lie lane-interleaved: stream row x
// holds one 32-bit word (two 16-bit codes) for each of the 32 lanes.
struct LockstepStream {
    int lanes = 0;              // 16-byte chunks per table row: 4 (64-byte rows, K <= 9) or 8 (128-byte rows, K <= 17)
    int TR = 0, ntiles = 0;
    int64_t ntasks = 0, nrows = 0;
    uint32_t* words = nullptr;      // [nrows][32]
    int64_t* task_row = nullptr;    // [ntasks + 1] first stream row of each task (tasks in tile-major order)
    int32_t* task_cell = nullptr;   // [ntasks][32] cell of each lane's unit, -1 = the lane has none
    int32_t* tile_task = nullptr;   // [ntiles + 1] first task of each tile
    int ncta = 0;
    int64_t* cta_row = nullptr;     // [ncta + 1] stream-row ranges of the CTAs (equal cost; a tile boundary inside a range is charged)
    int32_t* cta_tile = nullptr;    // [ncta] tile of the first row of each range
    int32_t* xheavy = nullptr;      // [N] entries of each cell with count > Q_REP_MAX (k_bayes adds them in fp64)
};

struct NcclApi;  // dlopen'ed function table (comm.cu)

// Which kernel does a restart's M-step sums: the sparse kernel (saturated posteriors) or the tiled SpMM.  The decision is a
// function of the restart's own history only -- `sat[r]` latches once the restart's dense-row count has fallen to the
// saturation threshold, `dense` is the count of the current iteration -- so a restart takes the same path (hence gives
// the same bits) whatever else is batched with it.  dense == nullptr: no gating (the kernel always runs).
struct Gate {
    const double* dense = nullptr;   // [r * stride]: posterior rows of restart r that do not pack into a sparse word
    int64_t stride = 0;
    double thr = 0.0;
    const int32_t* sat = nullptr;    // [R] latched "saturated" flags (k_bayes)
};
__device__ __forceinline__ bool gate_takes_sparse(const Gate& g, int r) {
    return (g.sat && g.sat[r]) || g.dense[(int64_t)r * g.stride] <= g.thr;
}

struct EmState {
    int R = 0, K = 0, KP = 0, max_iter = 0;
    int64_t stats_stride = 0;   // doubles per restart in `stats`
    double* T = nullptr;        // [R][2V][KP]  log2 P (row 2v) and log2(1-P) (row 2v+1)
    double* P = nullptr;        // [R][V][KP]
    double* L = nullptr;        // [R][N][KP]
    double* W = nullptr;        // [R][N][KP]
    double* stats = nullptr;    // [R][2V*KP + KP + 2]: S (alt/ref posterior-weighted sums), colsum[KP], LL, pad
    double* gstats = nullptr;   // the same summed over ranks (cells sharded over GPUs); nullptr on a single GPU
    double* lps = nullptr;      // [R][KP]
    double* partial = nullptr;  // [R][nblk][KP+2] per-block column mass, LL, dense-path cell count (fixed-order reduction)
    uint64_t* Tq = nullptr;     // [R][2V][2 * q_lanes] fixed-point copy of T (estep_q.cuh); nullptr = the E-step runs in fp64
    int32_t* qbad = nullptr;    // [R] the restart's table does not fit the fixed-point format: its E-step runs in fp64
    int q_lanes = 0, q_F = 0;
    // lock-step E-step of a 16-value-wide table (K = 10..17): Tq is kept as TWO tables of 64-byte rows, [2][R][2V][8] -- values
    // 0..8 (the ninth in byte pairs, as for K = 9) and values 9..16 -- and the kernel makes one pass over the stream per table
    int q_split = 0;
    unsigned long long* Lq = nullptr;   // [R][N][KP] integer E-step totals of the lock-step kernel (estep_lk.cuh); k_bayes leaves them at zero
    int32_t* sat = nullptr;     // [R] latched: the restart's dense-row count has been at or below the saturation threshold
    uint32_t* hq = nullptr;     // [R][h_stride] packed (state, 1 - weight) of a saturated posterior row, or SPARSE_DENSE
    int64_t h_stride = 0;       // N rounded up to 4 (16-byte multiples for the bulk copy)
    int32_t* sched = nullptr;   // [R][2] row tickets / finished-CTA count of the sparse M-step, then [R] k_finalize and [R] k_bayes block counts (self-resetting)
    int sparse_only = 0;        // host hint from the last poll: every restart is saturated -> launch only the sparse M-step
    double* hist = nullptr;     // [R][max_iter]
    double* ll_prev = nullptr;  // [R]
    int32_t* iters = nullptr;   // [R]
    int32_t* done = nullptr;    // [R]
    int nblk = 0;
    // One EM iteration (E-step + M-step launches) captured as a CUDA graph, per launch set: [check_conv][sparse_only].  A set is
    // captured once it has been launched GRAPH_AFTER times the plain way: the first pass does the lazy builds and attribute
    // calls a capture cannot hold, and capture + instantiation only pay back over a dozen or so replays (a restart of a
    // handful of iterations, e.g. one end-to-end call of bench.py, never captures).
    cudaGraphExec_t graph[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int graph_warm[2][2] = {{0, 0}, {0, 0}};
    int graph_launches[2][2] = {{0, 0}, {0, 0}};   // kernels inside each captured iteration
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
    scs::TiledStream TE, TM;   // tile-major copies for the current KP
    scs::TiledStream QE;       // tile-major cell-major stream of the fixed-point E-step (estep_q.cuh): KP field = lanes per row
    scs::LockstepStream LK;    // lock-step stream of the fixed-point E-step (estep_lk.cuh), the default form
    int q_state[2] = {0, 0};   // per row width (4 / 8 lanes): 0 = not tried, 1 = built, -1 = the pool does not fit the format
    int q_F = 0;               // fraction bits of the fixed-point tables of this pool
    uint64_t* qpart = nullptr; // per-(unit, lane) integer sums of the fixed-point E-step
    int64_t qpart_words = 0;
    double* part = nullptr;    // per-(tile,row) partial sums of the tiled SpMM
    int64_t part_doubles = 0;
    uint32_t smem_attr_set[3] = {0, 0, 0};  // bit KP/2: this device already has the large-shared-memory attribute of k_spmm_tiled<KP> / k_mstep_sparse<KP> / k_spmm_tiled<KP, packed>
    int64_t* m_bp = nullptr;     // [(nblocks+1)][2V] entry pointers of the blocked sparse M-step (pools whose cells do not fit one block)
    int m_bp_blocks = 0;
    int64_t m_bp_cells = 0;
    int32_t* m_order = nullptr;  // [2V] pseudo-rows of the M stream by descending length (job order of the sparse M-step; built on first use)
    int64_t* sum_alt = nullptr;  // [V] integer row sums (global after comm_init)
    int64_t* sum_ref = nullptr;  // [V]
    double* kalt = nullptr;      // [V]
    scs::EmState em;
    // multi-GPU
    void* comm = nullptr;        // ncclComm_t
    int rank = 0, world = 1;
    int64_t entries_global = 0;               // stored entries of the whole pool (all ranks)
    cudaStream_t comm_stream = nullptr;       // high-priority stream of the statistics all-reduce (cells sharded over ranks)
    cudaEvent_t ev_comm[10] = {};             // [0..7] M-step chunk done (main stream), [8] posterior tail ready, [9] all-reduce done (comm stream)
    int32_t* m_order_chunks = nullptr;        // [2V] job order of the sparse M-step inside each of the M_CHUNKS row ranges
    std::vector<int32_t> h_done;              // done flags as of the last poll (which restarts the collectives may leave out)
    // instrumentation
    int64_t launches = 0;
    int timing = 0;              // 0 = off, n = time every n-th E-step and every n-th M-step
    int64_t phase_seen[2] = {0, 0};   // E-steps / M-steps seen since timing was switched on
    bool sample_on = true;       // whether the launches issued right now belong to a sampled step
    int64_t m_samples = 0;       // sampled M-steps (an M-step is up to three launches)
    int variant = 0;
    int use_graphs = 1;          // replay one EM iteration as a CUDA graph (single GPU, kernel timing off)
    int32_t* h_poll = nullptr;   // pinned [2 * R]: stop flags and saturation latches as of the last poll
    int h_poll_cap = 0;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_pending;  // (kind, (start, stop))
    cudaEvent_t ev_iter[2] = {nullptr, nullptr};   // brackets the launches of the last scs_em_iterate
    double last_iter_ms = 0.0;
    double t_sum[5] = {0, 0, 0, 0, 0};   // E-step sums (all launches), M tiled SpMM, M sparse kernel, statistics all-reduce pieces, k_bayes
    int64_t t_cnt[5] = {0, 0, 0, 0, 0};
};

namespace scs {

// All device memory comes from the device's default stream-ordered pool on the context's stream (the release
// threshold is raised in scs_create), so creating and destroying contexts back to back does not pay cudaMalloc.
template <typename T>
int dev_alloc(scs_ctx* ctx, T** p, int64_t n) {
    *p = nullptr;
    if (n <= 0) n = 1;
    SCS_CUDA(cudaMallocAsync((void**)p, (size_t)n * sizeof(T), ctx->stream));
    ctx->device_bytes += n * (int64_t)sizeof(T);
    return 0;
}
template <typename T>
void dev_free(scs_ctx* ctx, T*& p) {
    if (p) cudaFreeAsync(p, ctx->stream);
    p = nullptr;
}

inline int round_even(int k) { return (k + 1) & ~1; }

// Stream-ordered temporaries of one build step: everything allocated through it goes back to the pool when the scope
// ends, on every path out (the early error returns included).
struct Scratch {
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t s) : st(s) {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    template <typename T>
    int get(T** p, int64_t n) {
        *p = nullptr;
        void* q = nullptr;
        SCS_CUDA(cudaMallocAsync(&q, (size_t)(n > 0 ? n : 1) * sizeof(T), st));
        ptrs.push_back(q);
        *p = (T*)q;
        return 0;
    }
    int get_bytes(void** p, size_t bytes) {
        *p = nullptr;
        SCS_CUDA(cudaMallocAsync(p, bytes ? bytes : 1, st));
        ptrs.push_back(*p);
        return 0;
    }
    ~Scratch() {
        for (void* q : ptrs) cudaFreeAsync(q, st);
    }
};

// layout.cu
int build_layout(scs_ctx* ctx, const int64_t* ref_indptr, const int32_t* ref_indices, const void* ref_data,
                 const int64_t* alt_indptr, const int32_t* alt_indices, const void* alt_data, int data_bytes);
void free_stream(scs_ctx* ctx, Stream& s);
void free_tiled(scs_ctx* ctx, TiledStream& t);
int build_tiled(scs_ctx* ctx, const Stream& s, TiledStream& t, int KP, int TR, int pad, int ncta, int ngroups, int order_gpw = 0);
// rows by descending light-entry count, ties by index; with nchunks > 1 that order inside each of nchunks equal row ranges
int build_row_order(scs_ctx* ctx, const Stream& s, int32_t** order, int nchunks = 1);
// per-(block of `block` gathered rows, stream row) entry pointers: bp[b * nrows + row], b = 0 .. nblocks
int build_block_ptrs(scs_ctx* ctx, const Stream& s, int nblocks, int64_t block, int64_t** bp);
// tile-major stream of the fixed-point E-step; *ok = 0 when the pool's cells are too deep for a useful fraction (F < 40)
int build_qtiled(scs_ctx* ctx, const Stream& s, TiledStream& t, int lanes, int TR, int ncta, int ngroups, int* F, int* ok);
// lock-step stream of the fixed-point E-step (same conditions as build_qtiled)
int build_lockstep(scs_ctx* ctx, const Stream& s, LockstepStream& t, int lanes, int TR, int ncta, int* F, int* ok);
void free_lockstep(scs_ctx* ctx, LockstepStream& t);

// em.cu
int launch_spmm(scs_ctx* ctx, int kind /*0=E,1=M*/, int KP, int R, const double* X, int64_t x_stride, double* Y,
                int64_t y_stride, const int32_t* done, Gate gate = Gate());
int em_free(scs_ctx* ctx);
// host [rows][K] <-> device [rows][KP] through a staging buffer and a (un)pitch kernel (strided 2-D copies of small rows are slow)
int upload_pitched(scs_ctx* ctx, double* dst_dev, const double* host, int64_t rows, int K, int KP);
int download_pitched(scs_ctx* ctx, double* host, const double* src_dev, int64_t rows, int K, int KP);

// comm.cu
int comm_allreduce_f64(scs_ctx* ctx, double* buf_dev, int64_t count);
int comm_allreduce_i64(scs_ctx* ctx, int64_t* buf_dev, int64_t count);
int comm_allreduce_segments(scs_ctx* ctx, const double* send, double* recv, int64_t stride, int64_t count, int nseg, const int32_t* skip,
                            cudaStream_t stream);
int comm_destroy(scs_ctx* ctx);

}  // namespace scs
