// Context lifetime, errors and instrumentation of the C-ABI (include/scsplit_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace scs {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

__global__ void k_kalt_from_sums(int64_t V, const int64_t* sum_alt, const int64_t* sum_ref, double* kalt) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    int64_t na = sum_alt[v] + 1, nr = sum_ref[v] + 1;
    kalt[v] = (double)na / (double)(nr + na);
}

int recompute_kalt(scs_ctx* ctx) {
    k_kalt_from_sums<<<(unsigned)((ctx->V + 255) / 256), 256, 0, ctx->stream>>>(ctx->V, ctx->sum_alt, ctx->sum_ref, ctx->kalt);
    ctx->launches++;
    SCS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace scs

using namespace scs;

extern "C" {

int scs_abi_version(void) { return SCS_ABI_VERSION; }

const char* scs_last_error(void) { return g_error.c_str(); }

int scs_device_count(int* count) {
    SCS_CHECK(count, "null argument");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return 1;
    }
    return 0;
}

int scs_create(scs_ctx** out, int device, int64_t V, int64_t N, const int64_t* ref_indptr, const int32_t* ref_indices,
               const void* ref_data, const int64_t* alt_indptr, const int32_t* alt_indices, const void* alt_data, int data_bytes) {
    SCS_CHECK(out, "null output");
    *out = nullptr;
    SCS_CHECK(ref_indptr && alt_indptr, "null indptr");
    int ndev = 0;
    SCS_TRY(scs_device_count(&ndev));
    SCS_CHECK(ndev > 0, "no CUDA device: scsplit_b200 has no CPU path");
    SCS_CHECK(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
    SCS_CUDA(cudaSetDevice(device));
    scs_ctx* ctx = new scs_ctx();
    ctx->device = device;
    ctx->V = V;
    ctx->N = N;
    int sms = 0;   // (cudaGetDeviceProperties costs milliseconds; one attribute query does not)
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess) ctx->sm_count = sms;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = ~0ull;   // keep freed blocks cached: contexts are created once per restart batch in `scSplit run`
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        delete ctx;
        return 1;
    }
    int rc = build_layout(ctx, ref_indptr, ref_indices, ref_data, alt_indptr, alt_indices, alt_data, data_bytes);
    if (rc) {
        std::string keep = g_error;
        scs_destroy(ctx);
        g_error = keep;
        return rc;
    }
    *out = ctx;
    return 0;
}

int scs_destroy(scs_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    em_free(ctx);
    comm_destroy(ctx);
    free_stream(ctx, ctx->E);
    free_stream(ctx, ctx->M);
    dev_free(ctx, ctx->m_order);
    dev_free(ctx, ctx->m_bp);
    dev_free(ctx, ctx->m_order_chunks);
    free_tiled(ctx, ctx->TE);
    free_tiled(ctx, ctx->TM);
    free_tiled(ctx, ctx->QE);
    free_lockstep(ctx, ctx->LK);
    dev_free(ctx, ctx->qpart);
    dev_free(ctx, ctx->part);
    dev_free(ctx, ctx->sum_alt);
    dev_free(ctx, ctx->sum_ref);
    dev_free(ctx, ctx->kalt);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& p : ctx->ev_pending) { cudaEventDestroy(p.second.first); cudaEventDestroy(p.second.second); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (auto e : ctx->ev_iter) if (e) cudaEventDestroy(e);
    if (ctx->h_poll) cudaFreeHost(ctx->h_poll);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

int scs_layout_info(scs_ctx* ctx, int64_t out[6]) {
    SCS_CHECK(ctx && out, "null argument");
    out[0] = ctx->nnz_ref;
    out[1] = ctx->nnz_alt;
    out[2] = ctx->nnz_union;
    out[3] = ctx->E.n_light;
    out[4] = ctx->E.n_heavy;
    out[5] = ctx->device_bytes;
    return 0;
}

int scs_kalt(scs_ctx* ctx, double* out) {
    SCS_CHECK(ctx && out, "null argument");
    SCS_CUDA(cudaSetDevice(ctx->device));
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    SCS_CUDA(cudaMemcpyAsync(out, ctx->kalt, (size_t)ctx->V * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int scs_launch_count(scs_ctx* ctx, int64_t* out) {
    SCS_CHECK(ctx && out, "null argument");
    *out = ctx->launches;
    return 0;
}

int scs_set_timing(scs_ctx* ctx, int enable) {
    SCS_CHECK(ctx, "null context");
    SCS_CHECK(enable >= 0, "timing period must be >= 0");
    ctx->timing = enable;
    ctx->phase_seen[0] = ctx->phase_seen[1] = 0;
    ctx->sample_on = true;
    return 0;
}

int scs_set_variant(scs_ctx* ctx, int variant) {
    SCS_CHECK(ctx, "null context");
    SCS_CHECK(variant >= 0 && variant <= 3, "variant must be 0, 1, 2 or 3");
    ctx->variant = variant;
    return 0;
}

int scs_kernel_times(scs_ctx* ctx, double out[8], int reset) {
    SCS_CHECK(ctx && out, "null argument");
    SCS_CUDA(cudaSetDevice(ctx->device));
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->comm_stream) SCS_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    for (auto& p : ctx->ev_pending) {
        float ms = 0.f;
        SCS_CUDA(cudaEventElapsedTime(&ms, p.second.first, p.second.second));
        ctx->t_sum[p.first] += (double)ms * 1000.0;
        ctx->t_cnt[p.first] += 1;
        ctx->ev_pool.push_back(p.second.first);
        ctx->ev_pool.push_back(p.second.second);
    }
    ctx->ev_pending.clear();
    // the M-step figure is the sum of its gated launches (tiled pair + sparse kernel; the inactive one costs ~2 us)
    out[0] = ctx->t_cnt[0] ? ctx->t_sum[0] / (double)ctx->t_cnt[0] : 0.0;
    out[1] = ctx->m_samples ? (ctx->t_sum[1] + ctx->t_sum[2]) / (double)ctx->m_samples : 0.0;
    out[2] = (double)ctx->t_cnt[0];
    out[3] = (double)ctx->m_samples;
    // the all-reduce of a step is issued in pieces: their durations are summed per sampled M-step
    out[4] = ctx->m_samples ? ctx->t_sum[3] / (double)ctx->m_samples : 0.0;
    out[5] = (double)ctx->t_cnt[3];
    out[6] = ctx->t_cnt[4] ? ctx->t_sum[4] / (double)ctx->t_cnt[4] : 0.0;
    out[7] = (double)ctx->t_cnt[4];
    if (reset) {
        for (int i = 0; i < 5; ++i) { ctx->t_sum[i] = 0.0; ctx->t_cnt[i] = 0; }
        ctx->m_samples = 0;
    }
    return 0;
}

}  // extern "C"
