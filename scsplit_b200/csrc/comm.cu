// Multi-GPU plumbing: NCCL is loaded with dlopen so the single-GPU library has no link-time dependency on it.
// Cells are sharded across ranks; the only data-path exchange is one sum all-reduce of
// the packed M-step statistics per iteration, issued on the context's stream between the M-step SpMM and
// k_finalize.  The unique id travels through the host (torch.distributed broadcast in scsplit_b200/dist.py).
#include <dlfcn.h>

#include "common.cuh"

namespace scs {

// minimal NCCL surface (stable since NCCL 2.x); types mirrored to avoid a header dependency at build time
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclInt64_ = 4, ncclFloat64_ = 8 };   // ncclDataType_t values
enum { ncclSum_ = 0 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.lib) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    SCS_CHECK(g_nccl.lib, "cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(g_nccl.lib, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(g_nccl.lib, "ncclGroupEnd");
    SCS_CHECK(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.GroupStart && g_nccl.GroupEnd,
              "libnccl lacks required symbols");
    return 0;
}

#define SCS_NCCL(expr)                                                                                   \
    do {                                                                                                 \
        int _r = (expr);                                                                                 \
        if (_r != 0) {                                                                                   \
            set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");  \
            return 1;                                                                                    \
        }                                                                                                \
    } while (0)

int comm_allreduce_f64(scs_ctx* ctx, double* buf, int64_t count) {
    if (ctx->world <= 1) return 0;
    SCS_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64_, ncclSum_, (ncclComm_t)ctx->comm, ctx->stream));
    return 0;
}

// Sum over ranks of `nseg` segments send[i * stride .. + count) -> recv[i * stride .. + count) (i = 0 .. nseg-1, segments
// whose skip[i] is set are left out), issued as one NCCL group on `stream`.  Out of place: what a rank sends is never
// overwritten, so a segment nobody recomputes (a restart that has stopped) keeps giving the same sum.
int comm_allreduce_segments(scs_ctx* ctx, const double* send, double* recv, int64_t stride, int64_t count, int nseg, const int32_t* skip,
                            cudaStream_t stream) {
    if (ctx->world <= 1) return 0;
    SCS_NCCL(g_nccl.GroupStart());
    int rc = 0;
    for (int i = 0; i < nseg && !rc; ++i) {
        if (skip && skip[i]) continue;
        rc = g_nccl.AllReduce(send + (int64_t)i * stride, recv + (int64_t)i * stride, (size_t)count, ncclFloat64_, ncclSum_, (ncclComm_t)ctx->comm,
                              stream);
    }
    const int rc2 = g_nccl.GroupEnd();
    SCS_NCCL(rc);
    SCS_NCCL(rc2);
    return 0;
}

int comm_allreduce_i64(scs_ctx* ctx, int64_t* buf, int64_t count) {
    if (ctx->world <= 1) return 0;
    SCS_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclInt64_, ncclSum_, (ncclComm_t)ctx->comm, ctx->stream));
    return 0;
}

int comm_destroy(scs_ctx* ctx) {
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    ctx->comm_stream = nullptr;
    for (auto& e : ctx->ev_comm) { if (e) cudaEventDestroy(e); e = nullptr; }
    ctx->comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
    return 0;
}

int recompute_kalt(scs_ctx* ctx);   // api.cu

}  // namespace scs

using namespace scs;

extern "C" {

int scs_comm_unique_id(uint8_t id[128]) {
    SCS_CHECK(id, "null argument");
    SCS_TRY(load_nccl());
    ncclUniqueId uid;
    SCS_NCCL(g_nccl.GetUniqueId(&uid));
    memcpy(id, uid.internal, 128);
    return 0;
}

int scs_comm_init(scs_ctx* ctx, const uint8_t id[128], int rank, int world) {
    SCS_CHECK(ctx && id, "null argument");
    SCS_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    SCS_CHECK(!ctx->comm, "communicator already initialised");
    if (world == 1) return 0;
    SCS_TRY(load_nccl());
    SCS_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    ncclComm_t comm = nullptr;
    SCS_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    {   // the statistics all-reduce runs on its own high-priority stream so that it overlaps the M-step kernels (em.cu: run_m)
        int lo = 0, hi = 0;
        SCS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        SCS_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
        for (auto& e : ctx->ev_comm) SCS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // k_alt is a function of the GLOBAL per-SNV sums (scSplit:101-103): reduce the integer sums once
    SCS_TRY(comm_allreduce_i64(ctx, ctx->sum_alt, ctx->V));
    SCS_TRY(comm_allreduce_i64(ctx, ctx->sum_ref, ctx->V));
    SCS_TRY(recompute_kalt(ctx));
    {   // entries of the whole pool: decisions every rank must take alike (how many pieces the all-reduce is cut into) use this
        int64_t* d = nullptr;
        int64_t h = ctx->M.n_light + ctx->M.n_heavy;
        SCS_CUDA(cudaMallocAsync((void**)&d, 8, ctx->stream));
        SCS_CUDA(cudaMemcpyAsync(d, &h, 8, cudaMemcpyHostToDevice, ctx->stream));
        int rc = comm_allreduce_i64(ctx, d, 1);
        if (!rc) rc = cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess;
        if (!rc) rc = cudaStreamSynchronize(ctx->stream) != cudaSuccess;
        cudaFreeAsync(d, ctx->stream);
        SCS_CHECK(!rc, "all-reduce of the entry count failed");
        ctx->entries_global = h;
    }
    SCS_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"
