// Seeding and post-EM reductions of the `scSplit run` path:
//   scs_seed_af         hard one-hot M-step, scSplit:137-145
//   scs_pattern_counts  non-zero pattern counts for the initialiser's trimming loop, scSplit:117-118,124-125
//   scs_doublet_scores  define_doublet, scSplit:210-225
//   scs_refine_scores   refine_doublets, scSplit:243-244
//   scs_cluster_counts  distinguishing_alleles reductions + ternary P/A code, scSplit:269-282
#include "em_kernels.cuh"

namespace scs {

__global__ void k_onehot(int64_t n, int KP, const int32_t* __restrict__ labels, double* __restrict__ W) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over R*N cells
    if (i >= n) return;
    const int lab = labels[i];
    double* w = W + i * KP;
    for (int k = 0; k < KP; ++k) w[k] = (k == lab) ? 1.0 : 0.0;
}

// out[c] = sum over alt entries of cell c: cnt * (1 - P[v, label[c]])
__global__ void __launch_bounds__(256) k_refine_scores(int64_t N, int64_t V, int K, const int64_t* __restrict__ lptr, const uint32_t* __restrict__ lcode,
                                                       const int64_t* __restrict__ hptr, const uint32_t* __restrict__ hcode,
                                                       const int32_t* __restrict__ hcnt, const double* __restrict__ P,
                                                       const int32_t* __restrict__ labels, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= N) return;
    const int lab = labels[c];
    double acc = 0.0;
    if (lab >= 0 && lab < K) {
        for (int64_t e = lptr[c] + lane; e < lptr[c + 1]; e += 32) {
            const int64_t g = (int64_t)(lcode[e] >> 1);          // alt entries are the gather rows g < V
            if (g < V) acc += 1.0 - P[g * K + lab];
        }
        for (int64_t h = hptr[c] + lane; h < hptr[c + 1]; h += 32) {
            const int64_t g = (int64_t)(hcode[h] >> 1);
            if (g < V) acc += (double)hcnt[h] * (1.0 - P[g * K + lab]);
        }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[c] = acc;
}

// partial[blk][k] = sum over the block's cells of mask[c] * L[c,k]   (fixed order)
template <int KP>
__global__ void __launch_bounds__(CELL_BLOCK) k_masked_colsum(int64_t N, const double* __restrict__ L, const uint8_t* __restrict__ mask,
                                                       double* __restrict__ partial) {
    __shared__ double s_red[8 * (KP + 1)];
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[KP + 1];
#pragma unroll
    for (int k = 0; k <= KP; ++k) v[k] = 0.0;
    if (c < N && mask[c]) {
        const double m = (double)mask[c];
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] = m * L[c * KP + k];
    }
    block_reduce_store<KP + 1>(v, s_red, partial + (int64_t)blockIdx.x * (KP + 1));
}

// one warp per SNV: union-pattern entries whose cell is kept
__global__ void __launch_bounds__(256) k_row_pattern_counts(int64_t V, const int64_t* __restrict__ lptr, const uint32_t* __restrict__ lcode,
                                                            const int64_t* __restrict__ hptr, const uint32_t* __restrict__ hcode,
                                                            const uint8_t* __restrict__ keep_rows, const uint8_t* __restrict__ keep_cols,
                                                            int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (v >= V) return;
    if (keep_rows && !keep_rows[v]) {   // a dropped SNV: nobody reads its count, and the later passes of the trimming loop keep
        if (lane == 0) out[v] = 0;      // only a few per cent of the rows, so skipping them is what makes those passes cheap
        return;
    }
    int64_t n = 0;
    for (int sel = 0; sel < 2; ++sel) {
        const int64_t pr = (int64_t)sel * V + v;
        for (int64_t e = lptr[pr] + lane; e < lptr[pr + 1]; e += 32) {
            const uint32_t code = lcode[e];
            n += !(code & 1u) && (!keep_cols || keep_cols[code >> 1]);
        }
        for (int64_t e = hptr[pr] + lane; e < hptr[pr + 1]; e += 32) {
            const uint32_t code = hcode[e];
            n += !(code & 1u) && (!keep_cols || keep_cols[code >> 1]);
        }
    }
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) out[v] = n;
}

// one warp per cell: union-pattern entries whose SNV is kept
__global__ void __launch_bounds__(256) k_col_pattern_counts(int64_t N, int64_t V, const int64_t* __restrict__ lptr, const uint32_t* __restrict__ lcode,
                                                            const int64_t* __restrict__ hptr, const uint32_t* __restrict__ hcode,
                                                            const uint8_t* __restrict__ keep_rows, const uint8_t* __restrict__ keep_cols,
                                                            int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= N) return;
    if (keep_cols && !keep_cols[c]) {
        if (lane == 0) out[c] = 0;
        return;
    }
    int64_t n = 0;
    for (int64_t e = lptr[c] + lane; e < lptr[c + 1]; e += 32) {
        const uint32_t code = lcode[e];
        const int64_t g = (int64_t)(code >> 1);
        n += !(code & 1u) && (!keep_rows || keep_rows[g >= V ? g - V : g]);
    }
    for (int64_t e = hptr[c] + lane; e < hptr[c + 1]; e += 32) {
        const uint32_t code = hcode[e];
        const int64_t g = (int64_t)(code >> 1);
        n += !(code & 1u) && (!keep_rows || keep_rows[g >= V ? g - V : g]);
    }
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) out[c] = n;
}

// one warp per SNV: integer per-cluster sums of alt (pseudo-row v) and ref (V+v) counts over labelled cells
__global__ void __launch_bounds__(256) k_cluster_counts(int64_t V, int K, const int64_t* __restrict__ lptr, const uint32_t* __restrict__ lcode,
                                                        const int64_t* __restrict__ hptr, const uint32_t* __restrict__ hcode,
                                                        const int32_t* __restrict__ hcnt, const int32_t* __restrict__ labels,
                                                        int64_t* __restrict__ n_ref, int64_t* __restrict__ n_alt) {
    __shared__ unsigned long long cnt[8][2][SCS_MAX_K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    for (int i = lane; i < 2 * SCS_MAX_K; i += 32) (&cnt[warp][0][0])[i] = 0ull;
    __syncwarp();
    if (v < V) {
        for (int sel = 0; sel < 2; ++sel) {
            const int64_t pr = (int64_t)sel * V + v;
            for (int64_t e = lptr[pr] + lane; e < lptr[pr + 1]; e += 32) {
                const int lab = labels[lcode[e] >> 1];
                if (lab >= 0 && lab < K) atomicAdd(&cnt[warp][sel][lab], 1ull);
            }
            for (int64_t e = hptr[pr] + lane; e < hptr[pr + 1]; e += 32) {
                const int lab = labels[hcode[e] >> 1];
                if (lab >= 0 && lab < K) atomicAdd(&cnt[warp][sel][lab], (unsigned long long)hcnt[e]);
            }
        }
        __syncwarp();
        for (int k = lane; k < K; k += 32) {
            n_alt[v * K + k] = (int64_t)cnt[warp][0][k];
            n_ref[v * K + k] = (int64_t)cnt[warp][1][k];
        }
    }
}

// scSplit:279-281: (N_alt >= 10) - ((N_ref >= 10) & (N_alt == 0))
__global__ void k_pa_codes(int64_t n, const int64_t* __restrict__ n_ref, const int64_t* __restrict__ n_alt, int8_t* __restrict__ pa) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pa[i] = (int8_t)((n_alt[i] >= 10) - ((n_ref[i] >= 10) && (n_alt[i] == 0)));
}

#define SCS_DISPATCH_KP2(KP_, ...)                                   \
    switch (KP_) {                                                   \
        case 2: { constexpr int KPc = 2; __VA_ARGS__; } break;       \
        case 4: { constexpr int KPc = 4; __VA_ARGS__; } break;       \
        case 6: { constexpr int KPc = 6; __VA_ARGS__; } break;       \
        case 8: { constexpr int KPc = 8; __VA_ARGS__; } break;       \
        case 10: { constexpr int KPc = 10; __VA_ARGS__; } break;     \
        case 12: { constexpr int KPc = 12; __VA_ARGS__; } break;     \
        case 14: { constexpr int KPc = 14; __VA_ARGS__; } break;     \
        case 16: { constexpr int KPc = 16; __VA_ARGS__; } break;     \
        case 18: { constexpr int KPc = 18; __VA_ARGS__; } break;     \
        case 20: { constexpr int KPc = 20; __VA_ARGS__; } break;     \
        case 22: { constexpr int KPc = 22; __VA_ARGS__; } break;     \
        case 24: { constexpr int KPc = 24; __VA_ARGS__; } break;     \
        case 26: { constexpr int KPc = 26; __VA_ARGS__; } break;     \
        case 28: { constexpr int KPc = 28; __VA_ARGS__; } break;     \
        case 30: { constexpr int KPc = 30; __VA_ARGS__; } break;     \
        case 32: { constexpr int KPc = 32; __VA_ARGS__; } break;     \
        default: scs::set_error("unsupported KP=%d", KP_); return 1; \
    }

}  // namespace scs

using namespace scs;

extern "C" {

int scs_seed_af(scs_ctx* ctx, int R, int K, const int32_t* labels, double* P0_out) {
    SCS_CHECK(ctx, "null context");
    SCS_CHECK(R >= 1 && K >= 1 && K <= SCS_MAX_K, "bad R=%d or K=%d", R, K);
    SCS_CHECK(labels && P0_out, "null argument");
    SCS_CUDA(cudaSetDevice(ctx->device));
    const int64_t V = ctx->V, N = ctx->N;
    const int KP = round_even(K);
    const int64_t stride = 2 * V * KP;
    cudaStream_t st = ctx->stream;
    Scratch s(ctx->stream);
    int32_t* d_lab; double *d_W, *d_S, *d_P;
    SCS_TRY(s.get(&d_lab, (int64_t)R * N));
    SCS_TRY(s.get(&d_W, (int64_t)R * N * KP));
    SCS_TRY(s.get(&d_S, (int64_t)R * stride));
    SCS_TRY(s.get(&d_P, (int64_t)R * V * KP));
    SCS_CUDA(cudaMemcpyAsync(d_lab, labels, (size_t)R * N * 4, cudaMemcpyHostToDevice, st));
    k_onehot<<<(unsigned)(((int64_t)R * N + 255) / 256), 256, 0, st>>>((int64_t)R * N, KP, d_lab, d_W);
    ctx->launches++;
    SCS_TRY(launch_spmm(ctx, 1, KP, R, d_W, N * KP, d_S, stride, nullptr));
    if (ctx->world > 1) SCS_TRY(comm_allreduce_f64(ctx, d_S, (int64_t)R * stride));
    const int rpb = finalize_rows_per_block(KP);
    dim3 grid((unsigned)((V + rpb - 1) / rpb), (unsigned)R);
    k_finalize<<<grid, FIN_THREADS, 0, st>>>(V, K, KP, d_S, stride, ctx->kalt, nullptr, d_P, TableOut{}, nullptr, PriorArgs{});
    ctx->launches++;
    SCS_TRY(download_pitched(ctx, P0_out, d_P, (int64_t)R * V, K, KP));
    SCS_CUDA(cudaStreamSynchronize(st));
    SCS_CUDA(cudaGetLastError());
    return 0;
}

int scs_pattern_counts(scs_ctx* ctx, const uint8_t* keep_rows, const uint8_t* keep_cols, int64_t* row_cnt, int64_t* col_cnt) {
    SCS_CHECK(ctx, "null context");
    SCS_CUDA(cudaSetDevice(ctx->device));
    const int64_t V = ctx->V, N = ctx->N;
    cudaStream_t st = ctx->stream;
    Scratch s(ctx->stream);
    uint8_t *d_kr = nullptr, *d_kc = nullptr;
    int64_t *d_rc, *d_cc;
    SCS_TRY(s.get(&d_rc, V));
    SCS_TRY(s.get(&d_cc, N));
    if (keep_rows) { SCS_TRY(s.get(&d_kr, V)); SCS_CUDA(cudaMemcpyAsync(d_kr, keep_rows, (size_t)V, cudaMemcpyHostToDevice, st)); }
    if (keep_cols) { SCS_TRY(s.get(&d_kc, N)); SCS_CUDA(cudaMemcpyAsync(d_kc, keep_cols, (size_t)N, cudaMemcpyHostToDevice, st)); }
    if (row_cnt) {
        k_row_pattern_counts<<<(unsigned)((V + 7) / 8), 256, 0, st>>>(V, ctx->M.lptr, ctx->M.lcode, ctx->M.hptr, ctx->M.hcode, d_kr, d_kc, d_rc);
        ctx->launches++;
        if (ctx->world > 1) SCS_TRY(comm_allreduce_i64(ctx, d_rc, V));
        SCS_CUDA(cudaMemcpyAsync(row_cnt, d_rc, (size_t)V * 8, cudaMemcpyDeviceToHost, st));
    }
    if (col_cnt) {
        k_col_pattern_counts<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(N, V, ctx->E.lptr, ctx->E.lcode, ctx->E.hptr, ctx->E.hcode, d_kr, d_kc, d_cc);
        ctx->launches++;
        SCS_CUDA(cudaMemcpyAsync(col_cnt, d_cc, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    }
    SCS_CUDA(cudaStreamSynchronize(st));
    SCS_CUDA(cudaGetLastError());
    return 0;
}

int scs_doublet_scores(scs_ctx* ctx, int K, const double* P, const uint8_t* mask, double* out) {
    SCS_CHECK(ctx, "null context");
    SCS_CHECK(K >= 1 && K <= SCS_MAX_K, "bad K=%d", K);
    SCS_CHECK(P && mask && out, "null argument");
    SCS_CUDA(cudaSetDevice(ctx->device));
    const int64_t V = ctx->V, N = ctx->N;
    const int KP = round_even(K);
    const int nblk = (int)((N + CELL_BLOCK - 1) / CELL_BLOCK);
    cudaStream_t st = ctx->stream;
    Scratch s(ctx->stream);
    double *d_P, *d_T, *d_L, *d_part, *d_out;
    uint8_t* d_mask;
    SCS_TRY(s.get(&d_P, V * KP));
    SCS_TRY(s.get(&d_T, 2 * V * KP));
    SCS_TRY(s.get(&d_L, N * KP));
    SCS_TRY(s.get(&d_part, (int64_t)nblk * (KP + 1)));
    SCS_TRY(s.get(&d_out, KP + 2));
    SCS_TRY(s.get(&d_mask, N));
    SCS_TRY(upload_pitched(ctx, d_P, P, V, K, KP));
    SCS_CUDA(cudaMemcpyAsync(d_mask, mask, (size_t)N, cudaMemcpyHostToDevice, st));
    const int rpb = finalize_rows_per_block(KP);
    dim3 grid((unsigned)((V + rpb - 1) / rpb), 1);
    TableOut tab;
    tab.T = d_T;
    k_finalize<<<grid, FIN_THREADS, 0, st>>>(V, K, KP, nullptr, 0, nullptr, d_P, nullptr, tab, nullptr, PriorArgs{});
    ctx->launches++;
    SCS_TRY(launch_spmm(ctx, 0, KP, 1, d_T, 2 * V * KP, d_L, N * KP, nullptr));
    SCS_DISPATCH_KP2(KP, (k_masked_colsum<KPc><<<nblk, CELL_BLOCK, 0, st>>>(N, d_L, d_mask, d_part)));
    k_reduce_stats<<<1, 256, 0, st>>>(KP, KP + 1, d_part, nblk, d_out, KP + 2, 0, 1, nullptr);
    ctx->launches += 2;
    if (ctx->world > 1) SCS_TRY(comm_allreduce_f64(ctx, d_out, KP));
    SCS_CUDA(cudaMemcpyAsync(out, d_out, (size_t)K * 8, cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    SCS_CUDA(cudaGetLastError());
    return 0;
}

int scs_refine_scores(scs_ctx* ctx, int K, const double* P, const int32_t* labels, double* out) {
    SCS_CHECK(ctx, "null context");
    SCS_CHECK(K >= 1 && K <= SCS_MAX_K, "bad K=%d", K);
    SCS_CHECK(P && labels && out, "null argument");
    SCS_CUDA(cudaSetDevice(ctx->device));
    const int64_t V = ctx->V, N = ctx->N;
    cudaStream_t st = ctx->stream;
    Scratch s(ctx->stream);
    double *d_P, *d_out;
    int32_t* d_lab;
    SCS_TRY(s.get(&d_P, V * K));
    SCS_TRY(s.get(&d_out, N));
    SCS_TRY(s.get(&d_lab, N));
    SCS_CUDA(cudaMemcpyAsync(d_P, P, (size_t)V * K * 8, cudaMemcpyHostToDevice, st));
    SCS_CUDA(cudaMemcpyAsync(d_lab, labels, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    k_refine_scores<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(N, V, K, ctx->E.lptr, ctx->E.lcode, ctx->E.hptr, ctx->E.hcode, ctx->E.hcnt,
                                                             d_P, d_lab, d_out);
    ctx->launches++;
    SCS_CUDA(cudaMemcpyAsync(out, d_out, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    SCS_CUDA(cudaGetLastError());
    return 0;
}

int scs_cluster_counts(scs_ctx* ctx, int K, const int32_t* labels, int64_t* n_ref, int64_t* n_alt, int8_t* pa) {
    SCS_CHECK(ctx, "null context");
    SCS_CHECK(K >= 1 && K <= SCS_MAX_K, "bad K=%d", K);
    SCS_CHECK(labels, "null labels");
    SCS_CUDA(cudaSetDevice(ctx->device));
    const int64_t V = ctx->V, N = ctx->N;
    cudaStream_t st = ctx->stream;
    Scratch s(ctx->stream);
    int64_t* d_cnt;   // [2][V][K]: ref then alt, contiguous for one all-reduce
    int32_t* d_lab;
    int8_t* d_pa;
    SCS_TRY(s.get(&d_cnt, 2 * V * K));
    SCS_TRY(s.get(&d_lab, N));
    SCS_TRY(s.get(&d_pa, V * K));
    SCS_CUDA(cudaMemcpyAsync(d_lab, labels, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    k_cluster_counts<<<(unsigned)((V + 7) / 8), 256, 0, st>>>(V, K, ctx->M.lptr, ctx->M.lcode, ctx->M.hptr, ctx->M.hcode, ctx->M.hcnt,
                                                              d_lab, d_cnt, d_cnt + V * K);
    ctx->launches++;
    if (ctx->world > 1) SCS_TRY(comm_allreduce_i64(ctx, d_cnt, 2 * V * K));
    k_pa_codes<<<(unsigned)((V * K + 255) / 256), 256, 0, st>>>(V * K, d_cnt, d_cnt + V * K, d_pa);
    ctx->launches++;
    if (n_ref) SCS_CUDA(cudaMemcpyAsync(n_ref, d_cnt, (size_t)V * K * 8, cudaMemcpyDeviceToHost, st));
    if (n_alt) SCS_CUDA(cudaMemcpyAsync(n_alt, d_cnt + V * K, (size_t)V * K * 8, cudaMemcpyDeviceToHost, st));
    if (pa) SCS_CUDA(cudaMemcpyAsync(pa, d_pa, (size_t)V * K, cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    SCS_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
