// EM driver: restart-batched run_EM (scSplit:148-162) with the stop rule evaluated on the device, plus the
// step-level E/M entry points used by the parity tests.
#include <algorithm>

#include "em_kernels.cuh"
#include "spmm_tiled.cuh"
#include "mstep_sparse.cuh"
#include "estep_lk.cuh"

namespace scs {

// compile-time state count (padded to even) dispatch
#define SCS_DISPATCH_KP(KP_, ...)                                     \
    switch (KP_) {                                                    \
        case 2: { constexpr int KPc = 2; __VA_ARGS__; } break;        \
        case 4: { constexpr int KPc = 4; __VA_ARGS__; } break;        \
        case 6: { constexpr int KPc = 6; __VA_ARGS__; } break;        \
        case 8: { constexpr int KPc = 8; __VA_ARGS__; } break;        \
        case 10: { constexpr int KPc = 10; __VA_ARGS__; } break;      \
        case 12: { constexpr int KPc = 12; __VA_ARGS__; } break;      \
        case 14: { constexpr int KPc = 14; __VA_ARGS__; } break;      \
        case 16: { constexpr int KPc = 16; __VA_ARGS__; } break;      \
        case 18: { constexpr int KPc = 18; __VA_ARGS__; } break;      \
        case 20: { constexpr int KPc = 20; __VA_ARGS__; } break;      \
        case 22: { constexpr int KPc = 22; __VA_ARGS__; } break;      \
        case 24: { constexpr int KPc = 24; __VA_ARGS__; } break;      \
        case 26: { constexpr int KPc = 26; __VA_ARGS__; } break;      \
        case 28: { constexpr int KPc = 28; __VA_ARGS__; } break;      \
        case 30: { constexpr int KPc = 30; __VA_ARGS__; } break;      \
        case 32: { constexpr int KPc = 32; __VA_ARGS__; } break;      \
        default: scs::set_error("unsupported KP=%d", KP_); return 1;  \
    }

static int timing_begin(scs_ctx* ctx, cudaEvent_t* a, cudaEvent_t* b) {
    if (!ctx->timing || !ctx->sample_on || ctx->ev_pending.size() >= 8192) { *a = *b = nullptr; return 0; }
    for (cudaEvent_t* e : {a, b}) {
        if (!ctx->ev_pool.empty()) { *e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
        else SCS_CUDA(cudaEventCreate(e));
    }
    SCS_CUDA(cudaEventRecord(*a, ctx->stream));
    return 0;
}
// E-steps (phase 0) and M-steps (phase 1) are sampled every ctx->timing-th time, so the event records cost the timed loop
// next to nothing; launches outside the EM loop are always sampled
struct PhaseSample {
    scs_ctx* ctx;
    PhaseSample(scs_ctx* c, int phase) : ctx(c) {
        if (c->timing) c->sample_on = (c->phase_seen[phase]++ % c->timing) == 0;
        if (c->timing && c->sample_on && phase == 1) c->m_samples++;
    }
    ~PhaseSample() { ctx->sample_on = true; }
};
static int timing_end(scs_ctx* ctx, int kind, cudaEvent_t a, cudaEvent_t b) {
    if (!a) return 0;
    SCS_CUDA(cudaEventRecord(b, ctx->stream));
    ctx->ev_pending.push_back({kind, {a, b}});
    return 0;
}

int launch_spmm(scs_ctx* ctx, int kind, int KP, int R, const double* X, int64_t x_stride, double* Y, int64_t y_stride,
                const int32_t* done, Gate gate) {
    const Stream& s = kind == 0 ? ctx->E : ctx->M;
    cudaEvent_t ea, eb;
    SCS_TRY(timing_begin(ctx, &ea, &eb));
    int variant = ctx->variant;
    if (variant == 0 || variant == 3) variant = tiled_supported(ctx, kind, KP) ? 2 : 1;
    if (variant == 2) {
        SCS_TRY(launch_spmm_tiled(ctx, kind, KP, R, X, x_stride, Y, y_stride, done, gate));
    } else {
        const int WPB = 8;
        dim3 grid((unsigned)((s.nrows + WPB - 1) / WPB), (unsigned)R);
        SCS_DISPATCH_KP(KP, (k_spmm_rows<KPc><<<grid, WPB * 32, 0, ctx->stream>>>(s.nrows, s.lptr, s.lcode, s.hptr, s.hcode, s.hcnt, X,
                                                                                  x_stride, Y, y_stride, done)));
        ctx->launches++;
    }
    SCS_TRY(timing_end(ctx, kind, ea, eb));
    SCS_CUDA(cudaGetLastError());
    return 0;
}

bool tiled_supported(scs_ctx*, int, int KP) { return KP >= 2 && KP <= 32 && (KP % 2) == 0; }

int launch_spmm_tiled(scs_ctx* ctx, int kind, int KP, int R, const double* X, int64_t x_stride, double* Y, int64_t y_stride,
                      const int32_t* done, Gate gate) {
    const Stream& s = kind == 0 ? ctx->E : ctx->M;
    TiledStream& t = kind == 0 ? ctx->TE : ctx->TM;
    const int TR = tile_rows_for(KP, s.xrows);
    const int ctas = ctx->sm_count > 0 ? ctx->sm_count : 148;
    // the E stream (one E-step per iteration for the whole run) gets the packed, build-time-ordered form where the table
    // width has one; the M stream is gathered through the tiled kernel only while posteriors are soft and stays as it is
    // (units much longer than the order pass handles -- deep coverage -- keep the plain form, which does not depend on the order)
    const int64_t ntiles_est = (s.xrows + TR - 1) / TR;
    const bool packed = kind == 0 && tiled_packed_for(KP) && s.n_light / std::max<int64_t>(1, ntiles_est * s.nrows) <= 160;
    const int ngroups = (TILED_THREADS / 32) * groups_per_warp_for(KP, packed);
    if (t.KP != KP || t.TR != TR || t.ncta != ctas || t.ngroups != ngroups)
        SCS_TRY(build_tiled(ctx, s, t, KP, TR, unit_pad_for(KP), ctas, ngroups, packed ? groups_per_warp_for(KP, true) : 0));
    const int64_t part_stride = t.nunits * KP;
    if (ctx->part_doubles < part_stride * R) {
        dev_free(ctx, ctx->part);
        SCS_TRY(dev_alloc(ctx, &ctx->part, part_stride * R));
        ctx->part_doubles = part_stride * R;
    }
    const size_t smem = ((((size_t)(TR + 1) * KP * 8) + 127) & ~(size_t)127) + 16;
    dim3 grid((unsigned)ctas, (unsigned)R);
    dim3 cgrid((unsigned)((s.nrows * (KP / 2) + 255) / 256), (unsigned)R);
    SCS_DISPATCH_KP(KP, {
        constexpr bool CAN_PACK = tiled_packed_for(KPc);
        const int slot = (packed && CAN_PACK) ? 2 : 0;            // smem_attr_set[0] plain, [2] packed instance
        if (!(ctx->smem_attr_set[slot] >> (KPc / 2) & 1u)) {      // per context: the attribute belongs to the device
            if (packed && CAN_PACK)
                SCS_CUDA(cudaFuncSetAttribute(k_spmm_tiled<KPc, CAN_PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES + 1024));
            else
                SCS_CUDA(cudaFuncSetAttribute(k_spmm_tiled<KPc, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES + 1024));
            ctx->smem_attr_set[slot] |= 1u << (KPc / 2);
        }
        if (packed && CAN_PACK)
            k_spmm_tiled<KPc, CAN_PACK><<<grid, TILED_THREADS, smem, ctx->stream>>>(s.nrows, TR, s.xrows, t.tptr, t.tcode, t.seg_ptr, t.seg_tile,
                                                                                 t.gb, X, x_stride, ctx->part, part_stride, done, gate);
        else
            k_spmm_tiled<KPc, false><<<grid, TILED_THREADS, smem, ctx->stream>>>(s.nrows, TR, s.xrows, t.tptr, t.tcode, t.seg_ptr, t.seg_tile,
                                                                              t.gb, X, x_stride, ctx->part, part_stride, done, gate);
        k_spmm_combine<KPc><<<cgrid, 256, 0, ctx->stream>>>(s.nrows, t.ntiles, ctx->part, part_stride, s.hptr, s.hcode, s.hcnt, X,
                                                          x_stride, Y, y_stride, done, gate);
    });
    ctx->launches += 2;
    return 0;
}

// Fixed-point E-step (estep_q.cuh): tile-major stream with inline counts, built on first use per row width.
// Returns the lanes per row (4 / 8) when this pool and K can use it, else 0.
// Two forms of the fixed-point E-step.  Lock-step (estep_lk.cuh: a lane per entry, integer totals by RED, no combine kernel) is
// the default: 84 us against 112 us at config 3 for the row-group form (estep_q.cuh: a group of lanes per entry, per-unit sums
// + k_combine_q).  Tables of 10..17 values per row are read as two tables of 64-byte rows in two passes (EmState::q_split);
// one pass over 128-byte rows (SCS_Q_WIDE=1) measured 509 us on a config-4 shard (25k cells, 67 table tiles, 44-code units: the
// flush -- two 8 x 8 transposes, 17 REDs -- comes round every 22 stream rows), the row-group form 492 us.
// SCS_Q_GROUPS=1 selects the row-group form for every width (A/B measurements).
static bool lk_wanted(int) { return std::getenv("SCS_Q_GROUPS") == nullptr; }
static bool lk_split(int lanes) { return lanes == 8 && std::getenv("SCS_Q_WIDE") == nullptr; }

static int q_prepare(scs_ctx* ctx, int K) {
    if (ctx->variant != 0 || std::getenv("SCS_NO_Q")) return 0;
    const int lanes = q_lanes_for(K);
    if (!lanes) return 0;
    const int idx = lanes == 4 ? 0 : 1;
    if (ctx->q_state[idx] < 0) return 0;
    const int ctas = ctx->sm_count > 0 ? ctx->sm_count : 148;
    const int ngroups = (TILED_THREADS / 32) * (32 / lanes);
    const int TR = q_tile_rows(lanes, ctx->E.xrows);
    if (lk_wanted(lanes)) {
        const int sl = lk_split(lanes) ? 4 : lanes;       // lanes (16-byte chunks) of the rows the stream addresses
        const int STR = q_tile_rows(sl, ctx->E.xrows);
        if (ctx->LK.lanes != sl || ctx->LK.TR != STR || ctx->LK.ncta != ctas || !ctx->LK.words) {
            int ok = 0, F = 0;
            if (build_lockstep(ctx, ctx->E, ctx->LK, sl, STR, ctas, &F, &ok)) return -1;
            ctx->q_state[idx] = ok ? 1 : -1;
            if (!ok) return 0;
            ctx->q_F = F;
        }
        return lanes;
    }
    if (ctx->QE.KP != lanes || ctx->QE.TR != TR || ctx->QE.ncta != ctas || ctx->QE.ngroups != ngroups || !ctx->QE.tcode) {
        int ok = 0, F = 0;
        if (build_qtiled(ctx, ctx->E, ctx->QE, lanes, TR, ctas, ngroups, &F, &ok)) return -1;
        ctx->q_state[idx] = ok ? 1 : -1;
        if (!ok) return 0;
        ctx->q_F = F;
    }
    return lanes;
}

static int launch_estep_lk(scs_ctx* ctx, const int32_t* done) {
    EmState& em = ctx->em;
    const LockstepStream& t = ctx->LK;
    const int lanes = t.lanes;                           // of the rows the stream addresses (4 when a wide table is split)
    const size_t smem = ((((size_t)(t.TR + 2) * lanes * 16) + 127) & ~(size_t)127) + 16;
    if (!(ctx->smem_attr_set[2] >> (26 + lanes / 8) & 1u)) {
        if (lanes == 4) SCS_CUDA(cudaFuncSetAttribute(k_estep_lk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES));
        else SCS_CUDA(cudaFuncSetAttribute(k_estep_lk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES));
        ctx->smem_attr_set[2] |= 1u << (26 + lanes / 8);
    }
    LkStreamView v;
    v.words = t.words; v.task_row = t.task_row; v.task_cell = t.task_cell; v.tile_task = t.tile_task; v.cta_row = t.cta_row; v.cta_tile = t.cta_tile;
    v.ntiles = t.ntiles; v.TR = t.TR; v.xrows = ctx->E.xrows;
    dim3 grid((unsigned)t.ncta, (unsigned)em.R);
    const int64_t xq_stride = 2 * ctx->V * 2 * lanes, lq_stride = ctx->N * em.KP;
    if (em.q_split) {                                    // values 0..8 from the first table, 9..K-1 from the second
        const int64_t split_stride = (int64_t)em.R * 2 * ctx->V * 8;
        k_estep_lk<4><<<grid, lk_threads(4), smem, ctx->stream>>>(v, em.Tq, xq_stride, em.Lq, lq_stride, 9, em.KP, 0, done, em.qbad);
        k_estep_lk<4><<<grid, lk_threads(4), smem, ctx->stream>>>(v, em.Tq + split_stride, xq_stride, em.Lq, lq_stride, em.K - 9, em.KP, 9, done, em.qbad);
        ctx->launches += 2;
    } else {
        if (lanes == 4) k_estep_lk<4><<<grid, lk_threads(4), smem, ctx->stream>>>(v, em.Tq, xq_stride, em.Lq, lq_stride, em.K, em.KP, 0, done, em.qbad);
        else k_estep_lk<8><<<grid, lk_threads(8), smem, ctx->stream>>>(v, em.Tq, xq_stride, em.Lq, lq_stride, em.K, em.KP, 0, done, em.qbad);
        ctx->launches += 1;
    }
    SCS_CUDA(cudaGetLastError());
    return 0;
}

static int launch_estep_q(scs_ctx* ctx, const int32_t* done) {
    EmState& em = ctx->em;
    if (em.Lq) return launch_estep_lk(ctx, done);
    TiledStream& t = ctx->QE;
    const int lanes = em.q_lanes;
    const int64_t part_stride = t.nunits * q_part_words(lanes);
    if (ctx->qpart_words < part_stride * em.R) {
        dev_free(ctx, ctx->qpart);
        SCS_TRY(dev_alloc(ctx, &ctx->qpart, part_stride * em.R));
        ctx->qpart_words = part_stride * em.R;
    }
    const size_t smem = ((((size_t)(t.TR + 2) * lanes * 16) + 127) & ~(size_t)127) + 16;
    const int64_t xq_stride = 2 * ctx->V * 2 * lanes;
    dim3 grid((unsigned)t.ncta, (unsigned)em.R);
    dim3 cgrid((unsigned)((ctx->N * lanes + 63) / 64), (unsigned)em.R);       // 256 threads = 64 (cell, lane) pairs x 4 tile classes
    const double scale_inv = ldexp(1.0, -em.q_F);
    if (!(ctx->smem_attr_set[2] >> (16 + lanes) & 1u)) {
        if (lanes == 4) SCS_CUDA(cudaFuncSetAttribute(k_estep_q<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES));
        else SCS_CUDA(cudaFuncSetAttribute(k_estep_q<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES));
        ctx->smem_attr_set[2] |= 1u << (16 + lanes);
    }
    if (lanes == 4) {
        k_estep_q<4><<<grid, TILED_THREADS, smem, ctx->stream>>>(t.TR, ctx->E.xrows, t.tptr, reinterpret_cast<const uint32_t*>(t.tcode), t.seg_ptr,
                                                                 t.seg_tile, t.gb, em.Tq, xq_stride, ctx->qpart, part_stride, done, em.qbad);
        k_combine_q<4><<<cgrid, 256, 0, ctx->stream>>>(ctx->N, em.K, em.KP, t.ntiles, ctx->qpart, part_stride, scale_inv, ctx->E.hptr, ctx->E.hcode,
                                                       ctx->E.hcnt, em.T, 2 * ctx->V * em.KP, em.L, ctx->N * em.KP, done, em.qbad);
    } else {
        k_estep_q<8><<<grid, TILED_THREADS, smem, ctx->stream>>>(t.TR, ctx->E.xrows, t.tptr, reinterpret_cast<const uint32_t*>(t.tcode), t.seg_ptr,
                                                                 t.seg_tile, t.gb, em.Tq, xq_stride, ctx->qpart, part_stride, done, em.qbad);
        k_combine_q<8><<<cgrid, 256, 0, ctx->stream>>>(ctx->N, em.K, em.KP, t.ntiles, ctx->qpart, part_stride, scale_inv, ctx->E.hptr, ctx->E.hcode,
                                                       ctx->E.hcnt, em.T, 2 * ctx->V * em.KP, em.L, ctx->N * em.KP, done, em.qbad);
    }
    ctx->launches += 2;
    SCS_CUDA(cudaGetLastError());
    return 0;
}

int em_free(scs_ctx* ctx) {
    EmState& em = ctx->em;
    for (auto& row : em.graph)
        for (auto& g : row)
            if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    dev_free(ctx, em.Tq); dev_free(ctx, em.qbad); dev_free(ctx, em.gstats); dev_free(ctx, em.Lq);
    ctx->h_done.clear();
    dev_free(ctx, em.T); dev_free(ctx, em.P); dev_free(ctx, em.L); dev_free(ctx, em.W); dev_free(ctx, em.stats); dev_free(ctx, em.lps);
    dev_free(ctx, em.hq); dev_free(ctx, em.sched);
    dev_free(ctx, em.partial); dev_free(ctx, em.hist); dev_free(ctx, em.ll_prev); dev_free(ctx, em.iters); dev_free(ctx, em.done);
    em = EmState();
    return 0;
}

static __global__ void k_selftest_exp2(const double* __restrict__ x, double* __restrict__ y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = exp2_bf(x[i]);
}

static __global__ void k_fill_u32(uint32_t* __restrict__ p, int64_t n, uint32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// dst[row][KP] <- src[row][K] (pad columns zero): small rows make a strided cudaMemcpy2D slow, a kernel does not care
__global__ void k_repitch(int64_t rows, int K, int KP, const double* __restrict__ src, double* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * KP) return;
    const int64_t row = i / KP;
    const int k = (int)(i - row * KP);
    dst[i] = k < K ? src[row * K + k] : 0.0;
}

// host [rows][K] -> device [rows][KP], stream-ordered (the staging buffer returns to the pool when the copy is done)
int upload_pitched(scs_ctx* ctx, double* dst, const double* host, int64_t rows, int K, int KP) {
    double* tmp = nullptr;
    SCS_CUDA(cudaMallocAsync((void**)&tmp, (size_t)rows * K * 8, ctx->stream));
    SCS_CUDA(cudaMemcpyAsync(tmp, host, (size_t)rows * K * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_repitch<<<(unsigned)((rows * KP + 255) / 256), 256, 0, ctx->stream>>>(rows, K, KP, tmp, dst);
    ctx->launches++;
    SCS_CUDA(cudaFreeAsync(tmp, ctx->stream));
    SCS_CUDA(cudaGetLastError());
    return 0;
}

// dst[row][K] <- src[row][KP]
__global__ void k_unpitch(int64_t rows, int K, int KP, const double* __restrict__ src, double* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * K) return;
    const int64_t row = i / K;
    dst[i] = src[row * KP + (i - row * K)];
}

// device [rows][KP] -> host [rows][K], stream-ordered
int download_pitched(scs_ctx* ctx, double* host, const double* src, int64_t rows, int K, int KP) {
    double* tmp = nullptr;
    SCS_CUDA(cudaMallocAsync((void**)&tmp, (size_t)rows * K * 8, ctx->stream));
    k_unpitch<<<(unsigned)((rows * K + 255) / 256), 256, 0, ctx->stream>>>(rows, K, KP, src, tmp);
    ctx->launches++;
    SCS_CUDA(cudaMemcpyAsync(host, tmp, (size_t)rows * K * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(cudaFreeAsync(tmp, ctx->stream));
    SCS_CUDA(cudaGetLastError());
    return 0;
}

static TableOut table_out(scs_ctx* ctx, int r0) {
    EmState& em = ctx->em;
    TableOut t;
    t.T = em.T + (int64_t)r0 * 2 * ctx->V * em.KP;
    if (em.Tq) {
        t.Tq = em.Tq + (int64_t)r0 * 2 * ctx->V * 2 * (em.q_split ? 4 : em.q_lanes);
        t.qlanes = em.q_lanes;
        t.split = em.q_split;
        t.split_stride = (int64_t)em.R * 2 * ctx->V * 8;
        t.qscale = ldexp(1.0, em.q_F);
        t.qbad = em.qbad + r0;
        t.qbad_next = em.qbad + em.R + r0;
    }
    return t;
}

static int tables_from_P(scs_ctx* ctx, int r0, int nr) {
    EmState& em = ctx->em;
    const int rpb = finalize_rows_per_block(em.KP);
    dim3 grid((unsigned)((ctx->V + rpb - 1) / rpb), (unsigned)nr);
    if (em.qbad) SCS_CUDA(cudaMemsetAsync(em.qbad + r0, 0, (size_t)nr * 4, ctx->stream));
    k_finalize<<<grid, FIN_THREADS, 0, ctx->stream>>>(ctx->V, em.K, em.KP, nullptr, 0, nullptr, em.P + (int64_t)r0 * ctx->V * em.KP, nullptr,
                                                      table_out(ctx, r0), nullptr, PriorArgs{});
    ctx->launches++;
    SCS_CUDA(cudaGetLastError());
    return 0;
}

// ---- the five launches of one iteration -----------------------------------------------------------------
static int run_e(scs_ctx* ctx, const int32_t* done) {
    EmState& em = ctx->em;
    PhaseSample sample(ctx, 0);
    const int64_t V = ctx->V, N = ctx->N;
    LSource src;
    if (em.Tq) {
        // fixed-point sums (k_estep_q, folded over the tiles into L by k_combine_q) for every restart whose table fits the format;
        // k_bayes gathers the rows of the (rare) restarts flagged by k_finalize from the fp64 table itself
        cudaEvent_t ea, eb;
        SCS_TRY(timing_begin(ctx, &ea, &eb));
        SCS_TRY(launch_estep_q(ctx, done));
        SCS_TRY(timing_end(ctx, 0, ea, eb));
        src.qbad = em.qbad;
        src.lptr = ctx->E.lptr; src.hptr = ctx->E.hptr; src.lcode = ctx->E.lcode; src.hcode = ctx->E.hcode; src.hcnt = ctx->E.hcnt;
        src.T = em.T;
        src.t_stride = 2 * V * em.KP;
        if (em.Lq) {
            src.Lq = em.Lq;
            src.xheavy = ctx->LK.xheavy;
            src.scale_inv = ldexp(1.0, -em.q_F);
            src.rep_max = Q_REP_MAX;
        }
    } else {
        SCS_TRY(launch_spmm(ctx, 0, em.KP, em.R, em.T, 2 * V * em.KP, em.L, N * em.KP, done));
    }
    cudaEvent_t ba, bb;                                  // the Bayes kernel (kind 4)
    SCS_TRY(timing_begin(ctx, &ba, &bb));
    dim3 grid((unsigned)em.nblk, (unsigned)em.R);
    SCS_DISPATCH_KP(em.KP, (k_bayes<KPc><<<grid, 1024, 0, ctx->stream>>>(N, em.K, src, em.L, em.W, em.lps, em.partial, em.nblk, em.hq,
                                                                              em.h_stride, done, em.stats + 2 * V * em.KP,
                                                                              em.stats_stride, em.sched + 3 * em.R, em.sat,
                                                                              (double)std::max<int64_t>(4, N / 128))));
    ctx->launches += 1;
    SCS_TRY(timing_end(ctx, 4, ba, bb));
    SCS_CUDA(cudaGetLastError());
    return 0;
}

// Row ranges of the M-step sums whose all-reduce may overlap the next range's kernel (cells sharded over ranks).  Measured on
// 8 x B200 at config 4 (25k cells per rank, profiles/r02_bench_8gpu_chunked.json vs r02_bench_8gpu.json): 4 ranges LOSE to
// one -- 1.00 ms against 0.92 ms per iteration.  The sparse M-step is a persistent kernel that fills every SM (768-1024 threads, ~210 KB of shared
// memory per CTA), so an NCCL kernel on the second stream only gets SMs as a range's CTAs drain, and then holds them against
// the next range's CTAs: the M-step went 155 -> 202 us and four latency-bound 4 MB collectives cost more than one of 17 MB.
// The default is therefore ONE range (plus the posterior tail, which does overlap the M-step); SCS_M_CHUNKS=n re-enables it.
constexpr int M_CHUNKS_MAX = 8;
static int m_chunks_wanted() {
    const char* e = std::getenv("SCS_M_CHUNKS");
    const int n = e ? atoi(e) : 1;
    return n < 1 ? 1 : n > M_CHUNKS_MAX ? M_CHUNKS_MAX : n;
}

// One sum all-reduce of a piece of the statistics of every live restart, stats -> gstats, on the communication stream, after
// `after` (an event of the main stream) has fired.  Timed (kind 3) when the step is sampled.
static int reduce_piece(scs_ctx* ctx, int64_t off, int64_t count, cudaEvent_t after) {
    EmState& em = ctx->em;
    SCS_CUDA(cudaEventRecord(after, ctx->stream));
    SCS_CUDA(cudaStreamWaitEvent(ctx->comm_stream, after, 0));
    cudaEvent_t ea = nullptr, eb = nullptr;
    if (ctx->timing && ctx->sample_on && ctx->ev_pending.size() < 8192) {
        for (cudaEvent_t* e : {&ea, &eb}) {
            if (!ctx->ev_pool.empty()) { *e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
            else SCS_CUDA(cudaEventCreate(e));
        }
        SCS_CUDA(cudaEventRecord(ea, ctx->comm_stream));
    }
    const int32_t* skip = (int)ctx->h_done.size() == em.R ? ctx->h_done.data() : nullptr;
    SCS_TRY(comm_allreduce_segments(ctx, em.stats + off, em.gstats + off, em.stats_stride, count, em.R, skip, ctx->comm_stream));
    if (ea) {
        SCS_CUDA(cudaEventRecord(eb, ctx->comm_stream));
        ctx->ev_pending.push_back({3, {ea, eb}});
    }
    return 0;
}

static int run_m(scs_ctx* ctx, const int32_t* done, int track, int check_conv) {
    EmState& em = ctx->em;
    PhaseSample sample(ctx, 1);
    const int64_t V = ctx->V, N = ctx->N;
    const bool multi = ctx->world > 1;
    // M-step sums: the sparse kernel when (nearly) all posterior rows are saturated, else the tiled SpMM; both are launched
    // and the dense-path cell count in the statistics tail (written by the last block of k_bayes, or by k_reduce_stats on the step-level path) gates which one does the work
    const SparsePlan plan = (ctx->variant == 0 || ctx->variant == 3) ? sparse_plan_for(N, em.KP) : SparsePlan();
    const int sparse_threads = plan.threads;
    const bool sparse = sparse_threads > 0;
    Gate gate;
    // N/32: measured at config 3 (tools/soft_phase_probe.py) -- the run's second iteration still has 14 % unsaturated cells,
    // where the sparse kernel (391 us) loses to the tiled SpMM (363 us); from 2.7 % on (third iteration) it wins
    if (sparse) {
        gate.dense = em.stats + 2 * V * em.KP + em.KP + 1;
        gate.stride = em.stats_stride;
        gate.thr = (double)std::max<int64_t>(16, N / 32);
        gate.sat = em.sat;
    }
    // Cells sharded over ranks: the sums of every rank are added by NCCL, piece by piece on a second stream, while the next
    // piece is still being computed -- first the posterior tail (column mass, LL; ready since k_bayes), then M_CHUNKS row
    // ranges of S, each as soon as its kernel has finished.  k_finalize waits for the last piece.
    const int64_t tail_off = 2 * V * em.KP;
    if (multi) SCS_TRY(reduce_piece(ctx, tail_off, em.KP + 2, ctx->ev_comm[8]));
    // (the sparse kernel is correct for ANY number of unsaturated cells, only slower; once a poll has seen the latch of every
    // live restart set, the tiled pair -- which those restarts can no longer take -- is not launched at all)
    if (!(sparse && em.sparse_only))
        SCS_TRY(launch_spmm(ctx, 1, em.KP, em.R, em.W, N * em.KP, em.stats, em.stats_stride, done, gate));
    // (more than one range only on request, and only from ~16 M entries per rank on; decided from the pool's global size so
    // that every rank issues the same collectives)
    const int nchunks = multi ? (int)std::min<int64_t>(m_chunks_wanted(), std::max<int64_t>(1, ctx->entries_global / ctx->world / (16 << 20))) : 1;
    const int64_t nrows = ctx->M.nrows;
    // rows [lo, hi) of chunk c: the keys of build_row_order put row i into chunk i * nchunks / nrows
    auto chunk_lo = [&](int c) { return ((int64_t)c * nrows + nchunks - 1) / nchunks; };
    if (sparse) {
        cudaEvent_t ea = nullptr, eb = nullptr;
        SCS_TRY(timing_begin(ctx, &ea, &eb));
        int32_t*& order = nchunks > 1 ? ctx->m_order_chunks : ctx->m_order;
        if (!order) SCS_TRY(build_row_order(ctx, ctx->M, &order, nchunks));
        // pools whose cells do not fit the kernel's shared-memory table in one piece: one launch per block of cells, each adding
        // its share of every pseudo-row (entry pointers per block are built on first use)
        if (plan.nblocks > 1 && (!ctx->m_bp || ctx->m_bp_blocks != plan.nblocks || ctx->m_bp_cells != plan.block_cells)) {
            dev_free(ctx, ctx->m_bp);
            SCS_TRY(build_block_ptrs(ctx, ctx->M, plan.nblocks, plan.block_cells, &ctx->m_bp));
            ctx->m_bp_blocks = plan.nblocks;
            ctx->m_bp_cells = plan.block_cells;
        }
        dim3 grid((unsigned)(ctx->sm_count > 0 ? ctx->sm_count : 148), (unsigned)em.R);
        for (int c = 0; c < nchunks; ++c) {
            const int64_t lo = chunk_lo(c), hi = chunk_lo(c + 1);
            for (int b = 0; b < plan.nblocks; ++b) {
                CellBlock blk;
                blk.lo = (int64_t)b * plan.block_cells;
                blk.n = std::min<int64_t>(plan.block_cells, em.h_stride - blk.lo);
                if (plan.nblocks > 1) { blk.lo_ptr = ctx->m_bp + (int64_t)b * nrows; blk.hi_ptr = ctx->m_bp + (int64_t)(b + 1) * nrows; }
                blk.accumulate = b > 0;
                const size_t smem = sparse_smem_bytes(blk.n, em.KP, sparse_threads);
                SCS_DISPATCH_KP(em.KP, {
                    if (!(ctx->smem_attr_set[1] >> (KPc / 2) & 1u)) {
                        SCS_CUDA(cudaFuncSetAttribute(k_mstep_sparse<KPc>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES + 1024));
                        ctx->smem_attr_set[1] |= 1u << (KPc / 2);
                    }
                    k_mstep_sparse<KPc><<<grid, sparse_threads, smem, ctx->stream>>>(nrows, hi - lo, N, em.K, order + lo, ctx->M.lptr, ctx->M.lcode,
                                                                                   ctx->M.hptr, ctx->M.hcode, ctx->M.hcnt, em.hq, em.h_stride, em.W,
                                                                                   em.stats, em.stats_stride, gate, done, em.sched, blk);
                });
                ctx->launches++;
            }
            if (multi) SCS_TRY(reduce_piece(ctx, lo * em.KP, (hi - lo) * em.KP, ctx->ev_comm[c]));
        }
        SCS_TRY(timing_end(ctx, 2, ea, eb));
    } else if (multi) {                                  // (same pieces as a rank whose shard takes the sparse kernel)
        for (int c = 0; c < nchunks; ++c) SCS_TRY(reduce_piece(ctx, chunk_lo(c) * em.KP, (chunk_lo(c + 1) - chunk_lo(c)) * em.KP, ctx->ev_comm[c]));
    }
    if (multi) {
        SCS_CUDA(cudaEventRecord(ctx->ev_comm[9], ctx->comm_stream));
        SCS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm[9], 0));
    }
    const double* S = multi ? em.gstats : em.stats;
    const int rpb = finalize_rows_per_block(em.KP);
    dim3 grid((unsigned)((V + rpb - 1) / rpb), (unsigned)em.R);
    PriorArgs prior;
    prior.tail_off = tail_off;
    prior.lps = em.lps; prior.hist = em.hist; prior.max_iter = em.max_iter; prior.ll_prev = em.ll_prev;
    prior.iters = em.iters; prior.done = em.done; prior.count = em.sched + 2 * em.R;
    prior.track = track; prior.check_conv = check_conv;
    k_finalize<<<grid, FIN_THREADS, 0, ctx->stream>>>(V, em.K, em.KP, S, em.stats_stride, ctx->kalt, nullptr, em.P, table_out(ctx, 0), done, prior);
    ctx->launches += 1;
    SCS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace scs

using namespace scs;

extern "C" {

int scs_selftest_exp2(scs_ctx* ctx, const double* x, double* y, int64_t n) {
    SCS_CHECK(ctx, "null context");
    SCS_CHECK(x && y && n >= 0, "bad arguments");
    if (n == 0) return 0;
    SCS_CUDA(cudaSetDevice(ctx->device));
    double *dx = nullptr, *dy = nullptr;
    SCS_TRY(dev_alloc(ctx, &dx, n));
    SCS_TRY(dev_alloc(ctx, &dy, n));
    SCS_CUDA(cudaMemcpyAsync(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_selftest_exp2<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dx, dy, n);
    ctx->launches++;
    SCS_CUDA(cudaMemcpyAsync(y, dy, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, dx);
    dev_free(ctx, dy);
    return 0;
}

int scs_em_end(scs_ctx* ctx) {
    SCS_CHECK(ctx, "null context");
    SCS_CUDA(cudaSetDevice(ctx->device));
    return em_free(ctx);
}

int scs_em_footprint(scs_ctx* ctx, int K, int64_t out[2]) {
    SCS_CHECK(ctx && out, "null argument");
    SCS_CHECK(K >= 1 && K <= SCS_MAX_K, "K must be in [1, %d] (got %d)", SCS_MAX_K, K);
    SCS_CUDA(cudaSetDevice(ctx->device));
    const int64_t V = ctx->V, N = ctx->N, KP = round_even(K);
    int64_t b = 0;
    b += 2 * V * KP * 8 + V * KP * 8 + 2 * N * KP * 8;                     // T, P, L, W
    b += (2 * V * KP + KP + 2) * 8 * (ctx->world > 1 ? 2 : 1);              // statistics (and their sum over ranks)
    b += 4 * N + SCS_MAX_ITER * 8 + 64 * (KP + 2) * 8 * 8;                  // packed posteriors, LL history, block partials
    const int lanes = q_lanes_for(K);
    if (lanes) {
        const int64_t ntq = (2 * V + q_tile_rows(lanes, 2 * V) - 1) / q_tile_rows(lanes, 2 * V);
        // fixed-point table; integer totals (lock-step form) or per-unit sums (row-group form)
        b += 2 * V * 2 * lanes * 8 + (lk_wanted(lanes) ? N * KP * 8 : ntq * N * q_part_words(lanes) * 8);
    }
    const int64_t nte = (2 * V + tile_rows_for((int)KP, 2 * V) - 1) / tile_rows_for((int)KP, 2 * V);
    const int64_t ntm = (N + tile_rows_for((int)KP, N) - 1) / tile_rows_for((int)KP, N);
    b += std::max(lanes ? (int64_t)0 : nte * N, ntm * 2 * V) * KP * 8;      // per-unit sums of the fp64 tiled SpMM (one buffer for both)
    b += (N + 3 * V) * KP * 8 + 4 * N;                                      // seeding: one-hot W, S, P, labels
    out[0] = b;
    size_t free_b = 0, total_b = 0;
    SCS_CUDA(cudaMemGetInfo(&free_b, &total_b));
    uint64_t reserved = 0, used = 0;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
        free_b += (size_t)(reserved - used);
    out[1] = (int64_t)free_b;
    return 0;
}

int scs_em_begin(scs_ctx* ctx, int R, int K, const double* P0, int max_iter) {
    SCS_CHECK(ctx, "null context");
    SCS_CHECK(R >= 1, "R must be >= 1");
    SCS_CHECK(K >= 1 && K <= SCS_MAX_K, "K must be in [1, %d] (got %d)", SCS_MAX_K, K);
    SCS_CHECK(max_iter >= 1, "max_iter must be >= 1");
    SCS_CHECK(P0, "P0 is null");
    SCS_CUDA(cudaSetDevice(ctx->device));
    em_free(ctx);
    EmState& em = ctx->em;
    const int64_t V = ctx->V, N = ctx->N;
    em.R = R; em.K = K; em.KP = round_even(K); em.max_iter = max_iter;
    const int KP = em.KP;
    em.stats_stride = 2 * V * KP + KP + 2;
    {   // Bayes blocks per restart: every block gets the same number of cells, so pick a count that fills whole waves of
        // one block per SM (R * nblk close to a multiple of the SM count), never more blocks than 64-cell chunks
        const int64_t nchunks = (N + bayes_cells_per_block(KP) - 1) / bayes_cells_per_block(KP);
        const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
        int64_t best = std::min<int64_t>(nchunks, std::max<int64_t>(1, sms / R));
        for (int waves = 1; waves <= 8 && best < nchunks; ++waves) {
            const int64_t b = std::min<int64_t>(nchunks, (int64_t)waves * sms / R);
            if (b < 1) continue;
            best = b;
            if ((double)(b * R) >= 0.9 * waves * sms) break;
        }
        em.nblk = (int)best;
    }
    {
        const int lanes = q_prepare(ctx, K);
        if (lanes < 0) return 1;
        if (lanes) {
            em.q_lanes = lanes;
            em.q_F = ctx->q_F;
            SCS_TRY(dev_alloc(ctx, &em.Tq, (int64_t)R * 2 * V * 2 * lanes));
            SCS_TRY(dev_alloc(ctx, &em.qbad, 2 * R));   // [R] flags, [R] their staging inside k_finalize
            SCS_CUDA(cudaMemsetAsync(em.qbad, 0, (size_t)R * 8, ctx->stream));
            if (lk_wanted(lanes)) {
                em.q_split = lk_split(lanes) ? 1 : 0;
                SCS_TRY(dev_alloc(ctx, &em.Lq, (int64_t)R * N * em.KP));
                SCS_CUDA(cudaMemsetAsync(em.Lq, 0, (size_t)R * N * em.KP * 8, ctx->stream));
            }
        }
    }
    SCS_TRY(dev_alloc(ctx, &em.T, (int64_t)R * 2 * V * KP));
    SCS_TRY(dev_alloc(ctx, &em.P, (int64_t)R * V * KP));
    SCS_TRY(dev_alloc(ctx, &em.L, (int64_t)R * N * KP));
    SCS_TRY(dev_alloc(ctx, &em.W, (int64_t)R * N * KP));
    SCS_TRY(dev_alloc(ctx, &em.stats, (int64_t)R * em.stats_stride));
    if (ctx->world > 1) {
        SCS_TRY(dev_alloc(ctx, &em.gstats, (int64_t)R * em.stats_stride));
        SCS_CUDA(cudaMemsetAsync(em.gstats, 0, (size_t)R * em.stats_stride * 8, ctx->stream));
    }
    SCS_TRY(dev_alloc(ctx, &em.lps, (int64_t)R * KP));
    SCS_TRY(dev_alloc(ctx, &em.partial, (int64_t)R * em.nblk * (KP + 2)));
    em.h_stride = (N + 3) / 4 * 4;
    SCS_TRY(dev_alloc(ctx, &em.hq, (int64_t)R * em.h_stride));
    k_fill_u32<<<(unsigned)((R * em.h_stride + 255) / 256), 256, 0, ctx->stream>>>(em.hq, R * em.h_stride, SPARSE_DENSE);
    ctx->launches++;
    SCS_TRY(dev_alloc(ctx, &em.done, 2 * R));      // [R] stop flags, [R] saturation latches: one copy per poll
    em.sat = em.done + R;
    SCS_CUDA(cudaMemsetAsync(em.done, 0, (size_t)R * 8, ctx->stream));
    SCS_TRY(dev_alloc(ctx, &em.sched, 4 * R));   // [R][2] sparse M-step tickets, [R] k_finalize and [R] k_bayes block counts
    SCS_CUDA(cudaMemsetAsync(em.sched, 0, (size_t)R * 16, ctx->stream));
    SCS_TRY(dev_alloc(ctx, &em.hist, (int64_t)R * max_iter));
    SCS_TRY(dev_alloc(ctx, &em.ll_prev, R));
    SCS_TRY(dev_alloc(ctx, &em.iters, R));
    cudaStream_t st = ctx->stream;
    SCS_CUDA(cudaMemsetAsync(em.W, 0, (size_t)R * N * KP * 8, st));
    SCS_CUDA(cudaMemsetAsync(em.L, 0, (size_t)R * N * KP * 8, st));
    SCS_CUDA(cudaMemsetAsync(em.stats, 0, (size_t)R * em.stats_stride * 8, st));
    SCS_CUDA(cudaMemsetAsync(em.hist, 0, (size_t)R * max_iter * 8, st));
    SCS_CUDA(cudaMemsetAsync(em.iters, 0, (size_t)R * 4, st));
    SCS_TRY(upload_pitched(ctx, em.P, P0, (int64_t)R * V, K, KP));
    // lP_s = [log2(1/K)] * K (scSplit:91); sum_log_likelihood = [1, 2] (scSplit:155) -> previous LL = 2
    std::vector<double> h_lps((size_t)R * KP, 0.0), h_prev((size_t)R, 2.0);
    const double l0 = log2(1.0 / (double)K);
    for (int r = 0; r < R; ++r)
        for (int k = 0; k < K; ++k) h_lps[(size_t)r * KP + k] = l0;
    SCS_CUDA(cudaMemcpyAsync(em.lps, h_lps.data(), h_lps.size() * 8, cudaMemcpyHostToDevice, st));
    SCS_CUDA(cudaMemcpyAsync(em.ll_prev, h_prev.data(), h_prev.size() * 8, cudaMemcpyHostToDevice, st));
    SCS_TRY(tables_from_P(ctx, 0, R));
    SCS_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int scs_estep(scs_ctx* ctx) {
    SCS_CHECK(ctx && ctx->em.R, "EM state not initialised (call scs_em_begin)");
    SCS_CUDA(cudaSetDevice(ctx->device));
    SCS_TRY(run_e(ctx, nullptr));
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int scs_mstep(scs_ctx* ctx) {
    SCS_CHECK(ctx && ctx->em.R, "EM state not initialised (call scs_em_begin)");
    SCS_CUDA(cudaSetDevice(ctx->device));
    EmState& em = ctx->em;
    // column mass of whatever W currently holds (it may have been injected with scs_em_set)
    dim3 grid((unsigned)em.nblk, (unsigned)em.R);
    SCS_DISPATCH_KP(em.KP, (k_colsum_partial<KPc><<<grid, 1024, 0, ctx->stream>>>(ctx->N, em.K, em.W, em.partial, em.nblk, em.hq,
                                                                                       em.h_stride)));
    k_reduce_stats<<<em.R, 256, 0, ctx->stream>>>(em.KP, em.KP + 2, em.partial, em.nblk, em.stats, em.stats_stride, 2 * ctx->V * em.KP, 0, nullptr);
    ctx->launches += 2;
    SCS_TRY(run_m(ctx, nullptr, 0, 0));
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Capture the launches of one iteration (run_e + run_m) for launch set [check_conv][sparse_only] and instantiate the graph.
// Nothing executes.  The set must have been launched the plain way before (lazy builds, function attributes): graph_warm says so.
// A failure switches graph replay off for the context and leaves plain launches in charge.
static int em_capture_set(scs_ctx* ctx, int a, int b) {
    EmState& em = ctx->em;
    if (em.graph[a][b]) return 0;
    const int32_t* done = a ? em.done : nullptr;
    const int64_t l0 = ctx->launches;
    const int64_t seen0 = ctx->phase_seen[0], seen1 = ctx->phase_seen[1];
    const int hint0 = em.sparse_only, timing0 = ctx->timing;
    em.sparse_only = b;                                  // what run_m looks at to leave the tiled pair out
    ctx->timing = 0;                                     // no event records inside a capture
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
    int rc = 0;
    if (e == cudaSuccess) {
        rc = run_e(ctx, done);
        if (!rc) rc = run_m(ctx, done, 1, a);
        e = cudaStreamEndCapture(ctx->stream, &g);
    }
    em.sparse_only = hint0;
    ctx->timing = timing0;
    ctx->phase_seen[0] = seen0;
    ctx->phase_seen[1] = seen1;
    const int nlaunch = (int)(ctx->launches - l0);
    ctx->launches = l0;
    cudaGraphExec_t x = nullptr;
    if (!rc && e == cudaSuccess && g) e = cudaGraphInstantiate(&x, g, 0);
    if (g) cudaGraphDestroy(g);
    if (rc || e != cudaSuccess || !x) {
        cudaGetLastError();
        ctx->use_graphs = 0;
        return rc;
    }
    em.graph[a][b] = x;
    em.graph_launches[a][b] = nlaunch;
    return 0;
}
// Graph replay of an iteration: single-rank contexts only.  (A multi-rank iteration -- NCCL all-reduces and the fork / join of
// the communication stream -- captures and replays correctly too, measured on 2 GPUs, but gains nothing: 0.1986 ms against
// 0.1985 ms per iteration at config 3, the iteration waits on the collective, not on the host.)  SCS_NO_GRAPH switches it off.
static bool graphs_allowed(const scs_ctx* ctx) { return ctx->world == 1 && !std::getenv("SCS_NO_GRAPH"); }
// the lazy builds of set [a][b] are done once the set itself or (for b = 1, a subset of the launches of b = 0) set [a][0] has run
static bool em_set_is_warm(const EmState& em, int a, int b) { return em.graph_warm[a][b] > 0 || (b == 1 && em.graph_warm[a][0] > 0); }

// One iteration: the launches of run_e + run_m, replayed from a CUDA graph where one exists for the current launch set.
static int em_one_iteration(scs_ctx* ctx, const int32_t* done, int check_conv) {
    EmState& em = ctx->em;
    constexpr int GRAPH_AFTER = 8;   // plain iterations of an EM state before graphs are captured by themselves (scs_em_capture does it at once)
    // an iteration whose kernels are being timed (every ctx->timing-th) is launched the plain way: events cannot sit inside a replay
    const bool sampled = ctx->timing && (ctx->phase_seen[0] % ctx->timing) == 0;
    const bool graphs = ctx->use_graphs && graphs_allowed(ctx) && !sampled;
    const int a = check_conv ? 1 : 0, b = em.sparse_only ? 1 : 0;
    if (graphs && !em.graph[a][b] && em_set_is_warm(em, a, b) &&
        em.graph_warm[a][0] + em.graph_warm[a][1] >= GRAPH_AFTER)
        SCS_TRY(em_capture_set(ctx, a, b));
    if (graphs && ctx->use_graphs && em.graph[a][b]) {
        SCS_CUDA(cudaGraphLaunch(em.graph[a][b], ctx->stream));
        ctx->launches += em.graph_launches[a][b];
        if (ctx->timing) { ctx->phase_seen[0]++; ctx->phase_seen[1]++; }     // (what run_e / run_m would have counted)
        return 0;
    }
    SCS_TRY(run_e(ctx, done));
    SCS_TRY(run_m(ctx, done, 1, check_conv));
    em.graph_warm[a][b]++;
    return 0;
}

int scs_em_capture(scs_ctx* ctx, int check_conv) {
    SCS_CHECK(ctx && ctx->em.R, "EM state not initialised (call scs_em_begin)");
    SCS_CUDA(cudaSetDevice(ctx->device));
    EmState& em = ctx->em;
    if (!ctx->use_graphs || !graphs_allowed(ctx)) return 0;
    const int a = check_conv ? 1 : 0;
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < 2 && ctx->use_graphs; ++b)
        if (em_set_is_warm(em, a, b)) SCS_TRY(em_capture_set(ctx, a, b));
    return 0;
}

int scs_em_iterate(scs_ctx* ctx, int n_iter, int check_conv) {
    SCS_CHECK(ctx && ctx->em.R, "EM state not initialised (call scs_em_begin)");
    SCS_CUDA(cudaSetDevice(ctx->device));
    EmState& em = ctx->em;
    const int32_t* done = check_conv ? em.done : nullptr;
    if (ctx->h_poll_cap < 2 * em.R) {
        if (ctx->h_poll) cudaFreeHost(ctx->h_poll);
        ctx->h_poll = nullptr;
        SCS_CUDA(cudaMallocHost((void**)&ctx->h_poll, (size_t)2 * em.R * 4));
        ctx->h_poll_cap = 2 * em.R;
    }
    const int32_t *h_done = ctx->h_poll, *h_sat = ctx->h_poll + em.R;
    if (!check_conv) ctx->h_done.clear();     // every restart takes part in the collectives of a fixed-length run
    const int POLL = 8;   // iterations between host polls of the device-side convergence flags and saturation latches
    if (!ctx->ev_iter[0]) { SCS_CUDA(cudaEventCreate(&ctx->ev_iter[0])); SCS_CUDA(cudaEventCreate(&ctx->ev_iter[1])); }
    SCS_CUDA(cudaEventRecord(ctx->ev_iter[0], ctx->stream));
    for (int i = 0; i < n_iter; ++i) {
        SCS_TRY(em_one_iteration(ctx, done, check_conv));
        if ((i + 1) % POLL == 0 || (check_conv && i + 1 == n_iter)) {
            SCS_CUDA(cudaMemcpyAsync(ctx->h_poll, em.done, (size_t)em.R * 8, cudaMemcpyDeviceToHost, ctx->stream));
            SCS_CUDA(cudaStreamSynchronize(ctx->stream));
            bool all = true, saturated = true;
            for (int r = 0; r < em.R; ++r) {
                all = all && h_done[r];
                if (!(check_conv && h_done[r])) saturated = saturated && h_sat[r];
            }
            em.sparse_only = saturated ? 1 : 0;
            if (check_conv) ctx->h_done.assign(h_done, h_done + em.R);   // (identical on every rank: functions of all-reduced sums)
            if (check_conv && all) break;
        }
    }
    SCS_CUDA(cudaEventRecord(ctx->ev_iter[1], ctx->stream));
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    SCS_CUDA(cudaEventElapsedTime(&ms, ctx->ev_iter[0], ctx->ev_iter[1]));
    ctx->last_iter_ms = (double)ms;
    return 0;
}

int scs_em_last_ms(scs_ctx* ctx, double* ms) {
    SCS_CHECK(ctx && ms, "null argument");
    *ms = ctx->last_iter_ms;
    return 0;
}

int scs_em_run(scs_ctx* ctx) {
    SCS_CHECK(ctx && ctx->em.R, "EM state not initialised (call scs_em_begin)");
    return scs_em_iterate(ctx, ctx->em.max_iter, 1);
}

int scs_em_status(scs_ctx* ctx, int32_t* iters, int32_t* converged, double* ll) {
    SCS_CHECK(ctx && ctx->em.R, "EM state not initialised (call scs_em_begin)");
    SCS_CUDA(cudaSetDevice(ctx->device));
    EmState& em = ctx->em;
    std::vector<int32_t> it((size_t)em.R);
    SCS_CUDA(cudaMemcpyAsync(it.data(), em.iters, (size_t)em.R * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < em.R; ++r) {
        if (iters) iters[r] = it[r];
        if (converged) converged[r] = it[r] < em.max_iter ? 1 : 0;   // scSplit:162
    }
    if (ll) SCS_CUDA(cudaMemcpyAsync(ll, em.ll_prev, (size_t)em.R * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int scs_em_info(scs_ctx* ctx, int64_t out[6]) {
    SCS_CHECK(ctx && ctx->em.R, "EM state not initialised (call scs_em_begin)");
    SCS_CHECK(out, "null output");
    SCS_CUDA(cudaSetDevice(ctx->device));
    EmState& em = ctx->em;
    out[0] = em.Tq ? em.q_lanes : 0;
    out[1] = em.Tq ? em.q_F : 0;
    out[2] = em.Lq ? ctx->LK.ntiles : em.Tq ? ctx->QE.ntiles : ctx->TE.ntiles;
    out[3] = 0;
    out[4] = em.Lq ? (em.q_split ? 3 : 2) : em.Tq ? 1 : 0;
    out[5] = em.Lq ? ctx->LK.ntasks : 0;
    if (em.Tq) {
        std::vector<int32_t> bad((size_t)em.R);
        SCS_CUDA(cudaMemcpyAsync(bad.data(), em.qbad, (size_t)em.R * 4, cudaMemcpyDeviceToHost, ctx->stream));
        SCS_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int r = 0; r < em.R; ++r) out[3] += bad[r] != 0;
    }
    return 0;
}

int scs_em_get(scs_ctx* ctx, int restart, int what, double* out) {
    SCS_CHECK(ctx && ctx->em.R, "EM state not initialised (call scs_em_begin)");
    EmState& em = ctx->em;
    SCS_CHECK(restart >= 0 && restart < em.R, "restart %d out of range", restart);
    SCS_CHECK(out, "null output");
    SCS_CUDA(cudaSetDevice(ctx->device));
    const int64_t V = ctx->V, N = ctx->N;
    const int K = em.K, KP = em.KP;
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    switch (what) {
        case SCS_P: SCS_TRY(download_pitched(ctx, out, em.P + (int64_t)restart * V * KP, V, K, KP)); break;
        case SCS_W: SCS_TRY(download_pitched(ctx, out, em.W + (int64_t)restart * N * KP, N, K, KP)); break;
        case SCS_L: SCS_TRY(download_pitched(ctx, out, em.L + (int64_t)restart * N * KP, N, K, KP)); break;
        case SCS_LPS: SCS_CUDA(cudaMemcpyAsync(out, em.lps + (int64_t)restart * KP, (size_t)K * 8, cudaMemcpyDeviceToHost, ctx->stream)); break;
        case SCS_LLHIST: SCS_CUDA(cudaMemcpyAsync(out, em.hist + (int64_t)restart * em.max_iter, (size_t)em.max_iter * 8, cudaMemcpyDeviceToHost, ctx->stream)); break;
        case SCS_STATS: SCS_CUDA(cudaMemcpyAsync(out, (em.gstats ? em.gstats : em.stats) + (int64_t)restart * em.stats_stride, (size_t)em.stats_stride * 8, cudaMemcpyDeviceToHost, ctx->stream)); break;
        default: SCS_CHECK(false, "unknown array id %d", what);
    }
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int scs_em_set(scs_ctx* ctx, int restart, int what, const double* in) {
    SCS_CHECK(ctx && ctx->em.R, "EM state not initialised (call scs_em_begin)");
    EmState& em = ctx->em;
    SCS_CHECK(restart >= 0 && restart < em.R, "restart %d out of range", restart);
    SCS_CHECK(in, "null input");
    SCS_CUDA(cudaSetDevice(ctx->device));
    const int64_t V = ctx->V, N = ctx->N;
    const int K = em.K, KP = em.KP;
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    switch (what) {
        case SCS_P:
            SCS_TRY(upload_pitched(ctx, em.P + (int64_t)restart * V * KP, in, V, K, KP));
            SCS_TRY(tables_from_P(ctx, restart, 1));
            break;
        case SCS_W: SCS_TRY(upload_pitched(ctx, em.W + (int64_t)restart * N * KP, in, N, K, KP)); break;
        case SCS_L: SCS_TRY(upload_pitched(ctx, em.L + (int64_t)restart * N * KP, in, N, K, KP)); break;
        case SCS_LPS: SCS_CUDA(cudaMemcpyAsync(em.lps + (int64_t)restart * KP, in, (size_t)K * 8, cudaMemcpyHostToDevice, ctx->stream)); break;
        default: SCS_CHECK(false, "array id %d cannot be set", what);
    }
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"
