// v2 SpMM: the gathered operand (log tables for the E-step, posteriors for the M-step) is staged tile by tile in
// shared memory with TMA bulk copies (cp.async.bulk + mbarrier), the count-1 entries are streamed from a tile-major
// 16-bit copy of the entry stream, and each table row is read by a GROUP of lanes with one conflict-free LDS.128 per
// lane (lane j owns columns 2j, 2j+1).  Per-(tile,row) partial sums go to HBM/L2 and a second kernel adds them in
// tile order together with the count>1 entries, so the summation order is fixed.  DESIGN.md "v2" has the roofline.
#pragma once
#include "common.cuh"

namespace scs {

constexpr int TILE_BYTES = 208 * 1024;       // table tile resident in shared memory (of 227 KB usable per CTA)
constexpr int TILED_THREADS = 1024;

// padding granule of a unit in the tile-major stream (entries); 2 keeps the 32-bit words aligned
inline int unit_pad_for(int /*KP*/) { return 2; }

inline int tile_rows_for(int KP, int64_t xrows) {
    int64_t tr = TILE_BYTES / (KP * 8);
    if (tr > 65534) tr = 65534;              // 16-bit in-tile row offsets; 65535 would collide with the zero row
    if (tr > xrows) tr = xrows;
    return (int)(tr < 1 ? 1 : tr);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

template <int KP>
struct TileCfg {
    static constexpr int CH = KP / 2;                                                       // 16-byte chunks per table row
    static constexpr int GS = CH <= 1 ? 1 : CH <= 2 ? 2 : CH <= 4 ? 4 : CH <= 8 ? 8 : 16;  // lanes per row group
    static constexpr int GPW = 32 / GS;                                                     // row groups per warp
};

// first unit whose tptr is >= target
__device__ __forceinline__ int64_t unit_lower_bound(const int64_t* __restrict__ tptr, int64_t nunits, int64_t target) {
    int64_t lo = 0, hi = nunits;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (tptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// grid = (CTAs, R).  CTA j owns the units [u_j, u_{j+1}) that split the tile-major entry stream into equal shares.
template <int KP>
__global__ void __launch_bounds__(TILED_THREADS, 1)
k_spmm_tiled(int64_t nrows, int ntiles, int TR, int64_t xrows, const int64_t* __restrict__ tptr,
             const uint16_t* __restrict__ tcode, const double* __restrict__ X, int64_t x_stride, double* __restrict__ part,
             int64_t part_stride, const int32_t* __restrict__ done) {
    using C = TileCfg<KP>;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tile = reinterpret_cast<double*>(smem);                           // [(TR + 1) * KP], row TR is all zeros
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (((size_t)(TR + 1) * KP * 8 + 127) & ~(size_t)127));
    __shared__ int64_t s_range[2];

    const int r = blockIdx.y;
    if (done && done[r]) return;
    X += (int64_t)r * x_stride;
    part += (int64_t)r * part_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int li = lane & (C::GS - 1);          // lane inside its row group: owns columns 2*li, 2*li+1
    const int g = lane / C::GS;                 // row group inside the warp
    const int64_t nunits = (int64_t)ntiles * nrows;
    // 32-bit shared address of this lane's 16-byte column pair inside table row 0; lanes beyond the last chunk
    // re-read the last chunk (same address inside the quarter-warp -> broadcast, no extra wavefront), results unused
    const uint32_t lane_base = smem_u32(smem) + (uint32_t)(li < C::CH ? li : C::CH - 1) * 16u;

    if (tid == 0) {
        const int64_t total = tptr[nunits];
        const int64_t G = gridDim.x, j = blockIdx.x;
        s_range[0] = (j == 0) ? 0 : unit_lower_bound(tptr, nunits, (total / G) * j + ((total % G) * j) / G);
        s_range[1] = (j == G - 1) ? nunits : unit_lower_bound(tptr, nunits, (total / G) * (j + 1) + ((total % G) * (j + 1)) / G);
        mbar_init(bar, 1);
    }
    for (int k = tid; k < KP; k += blockDim.x) tile[(size_t)TR * KP + k] = 0.0;
    __syncthreads();
    const int64_t u0 = s_range[0], u1 = s_range[1];
    if (u0 >= u1) return;
    uint32_t phase = 0;

    for (int64_t t = u0 / nrows; t * nrows < u1; ++t) {
        const int64_t ua = max(u0, t * nrows), ub = min(u1, (t + 1) * nrows);
        // ---- stage tile t of the operand: rows [t*TR, t*TR + rows_t) are contiguous in global memory
        __syncthreads();                                  // every warp is done reading the previous tile
        if (tid == 0) {
            const int64_t rows_t = min((int64_t)TR, xrows - t * TR);
            const uint32_t bytes = (uint32_t)(rows_t * KP * 8);
            const char* src = reinterpret_cast<const char*>(X + t * TR * KP);
            fence_proxy_async();                          // order the generic-proxy reads above before the async writes
            mbar_expect_tx(bar, bytes);
            for (uint32_t off = 0; off < bytes; off += 32768u)
                bulk_g2s(smem + off, src + off, min(32768u, bytes - off), bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;

        // ---- the CTA's units of this tile are dealt to its row groups in equal contiguous runs; every group walks
        // its run on its own (flush the finished unit, start the next), so the groups of a warp stay busy together
        // although their units differ in length.  The next unit's bounds are prefetched one unit ahead.
        const int64_t ngroups = (int64_t)nwarps * C::GPW, G = (int64_t)warp * C::GPW + g, n = ub - ua;
        int64_t u = ua + (n / ngroups) * G + min(G, n % ngroups);            // first unit of this group
        const int64_t uend = u + n / ngroups + (G < n % ngroups ? 1 : 0);
        int64_t ucur = 0;
        int64_t e_next = (u < uend) ? __ldg(tptr + u) : 0;
        const uint32_t* words = nullptr;
        int len = 0, pos = 0;
        bool have = false;
        double2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
        const uint32_t zero_pair = (uint32_t)TR | ((uint32_t)TR << 16);
        while (true) {
            if (pos >= len) {                         // this group's unit is finished (or it has none yet)
                if (have && li < C::CH)
                    *reinterpret_cast<double2*>(part + ucur * KP + 2 * li) = make_double2(acc0.x + acc1.x, acc0.y + acc1.y);
                have = u < uend;
                len = 0;
                pos = 0;
                if (have) {
                    const int64_t s = e_next;         // tptr[u]
                    e_next = __ldg(tptr + u + 1);
                    len = (int)(e_next - s);          // a multiple of the padding granule (zero-row entries fill it up)
                    words = reinterpret_cast<const uint32_t*>(tcode + s);
                    ucur = u++;
                    acc0 = make_double2(0.0, 0.0);
                    acc1 = make_double2(0.0, 0.0);
                }
            }
            if (!__any_sync(0xffffffffu, have)) break;
            // one 32-bit word = two 16-bit in-tile row offsets per lane; past the unit's end -> the zero row
            const uint32_t mine = (pos + 2 * li < len) ? __ldg(words + (pos >> 1) + li) : zero_pair;
            pos += 2 * C::GS;
            constexpr int UW = C::GS < 4 ? C::GS : 4;
#pragma unroll
            for (int j0 = 0; j0 < C::GS; j0 += UW) {
                double2 v[2 * UW];
#pragma unroll
                for (int jj = 0; jj < UW; ++jj) {     // 2*UW LDS.128 in flight per lane before the adds
                    const uint32_t w = __shfl_sync(0xffffffffu, mine, j0 + jj, C::GS);
                    v[2 * jj] = lds_f64x2(lane_base + (w & 0xffffu) * (KP * 8));
                    v[2 * jj + 1] = lds_f64x2(lane_base + (w >> 16) * (KP * 8));
                }
#pragma unroll
                for (int jj = 0; jj < UW; ++jj) {
                    acc0.x += v[2 * jj].x; acc0.y += v[2 * jj].y;
                    acc1.x += v[2 * jj + 1].x; acc1.y += v[2 * jj + 1].y;
                }
            }
        }
    }
}

// Y[row, :] = sum over tiles (in tile order) of the partials + the row's count>1 entries (gathered from global memory)
template <int KP>
__global__ void __launch_bounds__(256) k_spmm_combine(int64_t nrows, int ntiles, const double* __restrict__ part, int64_t part_stride,
                                                      const int64_t* __restrict__ hptr, const uint32_t* __restrict__ hcode,
                                                      const int32_t* __restrict__ hcnt, const double* __restrict__ X, int64_t x_stride,
                                                      double* __restrict__ Y, int64_t y_stride, const int32_t* __restrict__ done) {
    constexpr int CH = KP / 2;
    const int r = blockIdx.y;
    if (done && done[r]) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * CH) return;
    const int64_t row = i / CH;
    const int j = (int)(i - row * CH);
    part += (int64_t)r * part_stride;
    X += (int64_t)r * x_stride;
    double2 acc = make_double2(0.0, 0.0);
    for (int t = 0; t < ntiles; ++t) {
        const double2 p = *reinterpret_cast<const double2*>(part + ((int64_t)t * nrows + row) * KP + 2 * j);
        acc.x += p.x;
        acc.y += p.y;
    }
    for (int64_t h = hptr[row]; h < hptr[row + 1]; ++h) {
        const double cnt = (double)hcnt[h];
        const double2 x = __ldg(reinterpret_cast<const double2*>(X + (int64_t)(hcode[h] >> 1) * KP + 2 * j));
        acc.x += cnt * x.x;
        acc.y += cnt * x.y;
    }
    *reinterpret_cast<double2*>(Y + (int64_t)r * y_stride + row * KP + 2 * j) = acc;
}

bool tiled_supported(scs_ctx* ctx, int kind, int KP);
int launch_spmm_tiled(scs_ctx* ctx, int kind, int KP, int R, const double* X, int64_t x_stride, double* Y, int64_t y_stride,
                      const int32_t* done);

}  // namespace scs
