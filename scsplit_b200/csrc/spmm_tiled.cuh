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

// lanes per row group: the next power of two >= the 16-byte chunks of a table row -- except for 9 chunks (K = 17, the
// "16 samples + doublets" shape), where 8 lanes read the first 8 chunks (one full 128-byte wavefront) and the ninth chunk
// of an entry is read by the lane that holds the entry's index ("side" chunk, TileCfg::SIDE)
constexpr int tiled_group_size(int KP, bool packed = false) {
    const int ch = KP / 2;
    if (packed && ch == 5) return 5;
    return ch <= 1 ? 1 : ch <= 2 ? 2 : ch <= 4 ? 4 : ch <= 9 ? 8 : 16;
}
// "Packed" groups (K = 9: five 16-byte chunks per row): five lanes per group, six groups per warp, so one LDS.128 covers
// six rows in 30 lanes.  The six rows fit four conflict-free wavefronts only if their bank groups interleave, which is the
// case when the rows of neighbouring groups are consecutive mod 8 -- the tile-major stream of such a kernel is therefore
// ordered at build time (layout.cu: k_order_units): the entry at position p of group g's run gets, wherever the unit has
// one, a row with (row mod 8) == (g + p) mod 8.
constexpr bool tiled_packed_for(int KP) { return KP / 2 == 5; }
inline int groups_per_warp_for(int KP, bool packed = false) { return 32 / tiled_group_size(KP, packed); }

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

template <int KP, bool PACKED = false>
struct TileCfg {
    static constexpr int CH = KP / 2;                                                       // 16-byte chunks per table row
    static constexpr int GS = tiled_group_size(KP, PACKED);                                 // lanes per row group
    static constexpr int GPW = 32 / GS;                                                     // row groups per warp
    static constexpr bool SIDE = CH == GS + 1;                                              // one chunk beyond the group
};

// grid = (CTAs, R).  CTA j runs its segments of the static work plan (TiledStream::seg_ptr / seg_tile / gb).
//
// Inside a segment (a run of units of ONE tile) every row group owns a contiguous run of units whose entries are
// contiguous in the tile-major stream, and consumes them as one continuous stream of 2*GS-entry batches: there is no
// per-unit padding of the batches and the four (GPW) groups of a warp never wait for each other's unit ends.  Entries
// before the current unit's end go to `cur`, entries after it to `nxt` (predicated DADDs); when the stream passes a unit
// end the unit is flushed and nxt becomes cur.  A batch never crosses a second unit end (it is cut there), so two
// accumulator sets suffice.  Unit bounds are prefetched two units ahead.
template <int KP, bool PACKED>
__global__ void __launch_bounds__(TILED_THREADS, 1)
k_spmm_tiled(int64_t nrows, int TR, int64_t xrows, const int64_t* __restrict__ tptr, const uint16_t* __restrict__ tcode,
             const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_tile, const int64_t* __restrict__ gb,
             const double* __restrict__ X, int64_t x_stride, double* __restrict__ part, int64_t part_stride,
             const int32_t* __restrict__ done, const Gate gate) {
    using C = TileCfg<KP, PACKED>;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tile = reinterpret_cast<double*>(smem);                           // [(TR + 1) * KP], row TR is all zeros
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (((size_t)(TR + 1) * KP * 8 + 127) & ~(size_t)127));

    const int r = blockIdx.y;
    if (done && done[r]) return;
    if (gate.dense && gate_takes_sparse(gate, r)) return;   // the sparse M-step kernel handles this restart
    X += (int64_t)r * x_stride;
    part += (int64_t)r * part_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // lanes past the last whole group (GS = 5: lanes 30, 31) shadow the last lane of the last group -- same shared address,
    // so no extra wavefront -- and own no units
    const bool idle = lane >= C::GS * C::GPW;
    const int li = idle ? C::GS - 1 : lane % C::GS;   // lane inside its row group: owns columns 2*li, 2*li+1
    const int g = idle ? C::GPW - 1 : lane / C::GS;   // row group inside the warp
    const int ngroups = (int)(blockDim.x >> 5) * C::GPW;
    const int G = warp * C::GPW + g;
    // 32-bit shared address of this lane's 16-byte column pair inside table row 0; lanes beyond the last chunk
    // re-read the last chunk (same address inside the quarter-warp -> broadcast, no extra wavefront), results unused
    const uint32_t lane_base = smem_u32(smem) + (uint32_t)(li < C::CH ? li : C::CH - 1) * 16u;
    const uint32_t zero_pair = (uint32_t)TR | ((uint32_t)TR << 16);
    constexpr int BE = 2 * C::GS;               // entries per batch of one group

    if (tid == 0) mbar_init(bar, 1);
    for (int k = tid; k < KP; k += blockDim.x) tile[(size_t)TR * KP + k] = 0.0;
    __syncthreads();
    uint32_t phase = 0;
    int loaded_tile = -1;

    const int seg_end = seg_ptr[blockIdx.x + 1];
    for (int seg = seg_ptr[blockIdx.x]; seg < seg_end; ++seg) {
        const int t = seg_tile[seg];
        if (t != loaded_tile) {
            // ---- stage tile t of the operand: rows [t*TR, t*TR + rows_t) are contiguous in global memory
            __syncthreads();                              // every warp is done reading the previous tile
            if (tid == 0) {
                const int64_t rows_t = min((int64_t)TR, xrows - (int64_t)t * TR);
                const uint32_t bytes = (uint32_t)(rows_t * KP * 8);
                const char* src = reinterpret_cast<const char*>(X + (int64_t)t * TR * KP);
                fence_proxy_async();                      // order the generic-proxy reads above before the async writes
                mbar_expect_tx(bar, bytes);
                for (uint32_t off = 0; off < bytes; off += 32768u)
                    bulk_g2s(smem + off, src + off, min(32768u, bytes - off), bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            loaded_tile = t;
        }
        const int64_t* gbs = gb + (int64_t)seg * (ngroups + 1);
        int64_t u = __ldg(gbs + G);                       // current unit
        const int64_t uend = idle ? u : __ldg(gbs + G + 1);
        const bool act = u < uend;                        // (a group without units still takes part in the warp's shuffles)
        int64_t e = act ? __ldg(tptr + u) : 0;            // stream position (entries); even
        const int64_t eend = act ? __ldg(tptr + uend) : 0;
        // ends of units u, u+1, u+2 (clamped to the end of this group's run)
        int64_t b1 = act ? __ldg(tptr + u + 1) : 0;
        int64_t b2 = (u + 2 <= uend) ? __ldg(tptr + u + 2) : eend;
        int64_t b3 = (u + 3 <= uend) ? __ldg(tptr + u + 3) : eend;
        double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);   // current unit (even / odd entries)
        double2 n0 = make_double2(0.0, 0.0), n1 = make_double2(0.0, 0.0);   // next unit
        double2 sc = make_double2(0.0, 0.0), sn = make_double2(0.0, 0.0);   // side chunk of this lane's own entries (SIDE)
        // all GPW groups of the warp iterate together; a group that has finished feeds zero rows until the others end
        // this lane's index word (two entries) of the NEXT batch is always in flight: the stream of a group is contiguous,
        // so the next batch starts at `lim` whatever the unit bounds turn out to be (the stream buffer has slack at its end)
        uint32_t word = act ? __ldcs(reinterpret_cast<const uint32_t*>(tcode + e) + li) : zero_pair;
        while (__any_sync(0xffffffffu, u < uend)) {
            const int64_t lim = min(e + BE, b2);          // never past the end of unit u+1
            const int nvalid = (u < uend) ? (int)(lim - e) : 0;
            const int ncur = (int)min((int64_t)nvalid, max((int64_t)0, b1 - e));   // entries of this batch that belong to unit u
            const uint32_t mine = (2 * li < nvalid) ? word : zero_pair;
            if (u < uend) word = __ldcs(reinterpret_cast<const uint32_t*>(tcode + lim) + li);
            constexpr int UW = C::GS < 4 ? C::GS : C::SIDE ? 2 : C::GS % 4 ? C::GS : 4;   // divides GS (the side variant has fewer registers to spare)
#pragma unroll
            for (int j0 = 0; j0 < C::GS; j0 += UW) {
                double2 v[2 * UW];
#pragma unroll
                for (int jj = 0; jj < UW; ++jj) {         // 2*UW LDS.128 in flight per lane before the adds
                    const uint32_t w = __shfl_sync(0xffffffffu, mine, g * C::GS + j0 + jj);
                    SCS_ASSERT((w & 0xffffu) <= (uint32_t)TR && (w >> 16) <= (uint32_t)TR);
                    v[2 * jj] = lds_f64x2(lane_base + (w & 0xffffu) * (KP * 8));
                    v[2 * jj + 1] = lds_f64x2(lane_base + (w >> 16) * (KP * 8));
                }
#pragma unroll
                for (int jj = 0; jj < UW; ++jj) {
                    const int p0 = 2 * (j0 + jj);
                    if (p0 < ncur) { c0.x += v[2 * jj].x; c0.y += v[2 * jj].y; } else { n0.x += v[2 * jj].x; n0.y += v[2 * jj].y; }
                    if (p0 + 1 < ncur) { c1.x += v[2 * jj + 1].x; c1.y += v[2 * jj + 1].y; } else { n1.x += v[2 * jj + 1].x; n1.y += v[2 * jj + 1].y; }
                }
            }
            if constexpr (C::SIDE) {
                // chunk GS of the two entries whose indices this lane holds: 8 different rows per quarter-warp (a few
                // wavefronts per 32 entries instead of one nearly empty wavefront per entry)
                const uint32_t side_base = smem_u32(smem) + (uint32_t)C::GS * 16u;
                const double2 s0 = lds_f64x2(side_base + (mine & 0xffffu) * (KP * 8));
                const double2 s1 = lds_f64x2(side_base + (mine >> 16) * (KP * 8));
                if (2 * li < ncur) { sc.x += s0.x; sc.y += s0.y; } else { sn.x += s0.x; sn.y += s0.y; }
                if (2 * li + 1 < ncur) { sc.x += s1.x; sc.y += s1.y; } else { sn.x += s1.x; sn.y += s1.y; }
            }
            e = (u < uend) ? lim : e;
            // flush every unit the stream has passed (an empty unit is passed immediately and stores zeros)
            while (u < uend && e >= b1) {
                SCS_ASSERT(u >= 0 && (u + 1) * KP <= part_stride);
                if (li < C::CH)
                    *reinterpret_cast<double2*>(part + u * KP + 2 * li) = make_double2(c0.x + c1.x, c0.y + c1.y);
                if constexpr (C::SIDE) {                  // the side chunk's sum is spread over the lanes of the group
                    const uint32_t gmask = ((1u << C::GS) - 1u) << (g * C::GS);
                    double2 s = sc;
#pragma unroll
                    for (int o = C::GS / 2; o; o >>= 1) {
                        s.x += __shfl_xor_sync(gmask, s.x, o);
                        s.y += __shfl_xor_sync(gmask, s.y, o);
                    }
                    if (li == 0) *reinterpret_cast<double2*>(part + u * KP + 2 * C::GS) = s;
                    sc = sn;
                    sn = make_double2(0.0, 0.0);
                }
                c0 = n0; c1 = n1;
                n0 = make_double2(0.0, 0.0); n1 = make_double2(0.0, 0.0);
                ++u;
                b1 = b2; b2 = b3;
                b3 = (u + 3 <= uend) ? __ldg(tptr + u + 3) : eend;
            }
        }
    }
}

// Y[row, :] = sum over tiles (in tile order) of the partials + the row's count>1 entries (gathered from global memory)
template <int KP>
__global__ void __launch_bounds__(256) k_spmm_combine(int64_t nrows, int ntiles, const double* __restrict__ part, int64_t part_stride,
                                                      const int64_t* __restrict__ hptr, const uint32_t* __restrict__ hcode,
                                                      const int32_t* __restrict__ hcnt, const double* __restrict__ X, int64_t x_stride,
                                                      double* __restrict__ Y, int64_t y_stride, const int32_t* __restrict__ done,
                                                      const Gate gate) {
    constexpr int CH = KP / 2;
    const int r = blockIdx.y;
    if (done && done[r]) return;
    if (gate.dense && gate_takes_sparse(gate, r)) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * CH) return;
    const int64_t row = i / CH;
    const int j = (int)(i - row * CH);
    part += (int64_t)r * part_stride;
    X += (int64_t)r * x_stride;
    double2 acc = make_double2(0.0, 0.0);
    // the partials of a row are ntiles strided 16-byte loads: eight are requested before the (ordered) adds, which is what
    // keeps enough bytes in flight at ~700 threads per SM
    const double* prow = part + row * KP + 2 * j;
    const int64_t tstride = nrows * KP;
    int t = 0;
    constexpr int CB = 8;   // (12 and 23 in flight measured no better)
    for (; t + CB <= ntiles; t += CB) {
        double2 p[CB];
#pragma unroll
        for (int u = 0; u < CB; ++u) p[u] = __ldcs(reinterpret_cast<const double2*>(prow + (int64_t)(t + u) * tstride));
#pragma unroll
        for (int u = 0; u < CB; ++u) { acc.x += p[u].x; acc.y += p[u].y; }
    }
    for (; t < ntiles; ++t) {
        const double2 p = __ldcs(reinterpret_cast<const double2*>(prow + (int64_t)t * tstride));
        acc.x += p.x;
        acc.y += p.y;
    }
    // count>1 entries: a few dozen per row, each an index load followed by a dependent table gather -- eight are in flight
    // before the (ordered) adds, otherwise this short loop is the longest latency chain of the kernel
    const int64_t hb = hptr[row], he = hptr[row + 1];
    int64_t h = hb;
    for (; h + 8 <= he; h += 8) {
        double cnt[8];
        double2 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            cnt[u] = (double)__ldg(hcnt + h + u);
            x[u] = __ldg(reinterpret_cast<const double2*>(X + (int64_t)(__ldg(hcode + h + u) >> 1) * KP + 2 * j));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += cnt[u] * x[u].x; acc.y += cnt[u] * x[u].y; }
    }
    for (; h < he; ++h) {
        const double cnt = (double)__ldg(hcnt + h);
        const double2 x = __ldg(reinterpret_cast<const double2*>(X + (int64_t)(__ldg(hcode + h) >> 1) * KP + 2 * j));
        acc.x += cnt * x.x;
        acc.y += cnt * x.y;
    }
    *reinterpret_cast<double2*>(Y + (int64_t)r * y_stride + row * KP + 2 * j) = acc;
}

bool tiled_supported(scs_ctx* ctx, int kind, int KP);
int launch_spmm_tiled(scs_ctx* ctx, int kind, int KP, int R, const double* X, int64_t x_stride, double* Y, int64_t y_stride,
                      const int32_t* done, Gate gate);

}  // namespace scs
