// E-step sums in exact fixed point, one LANE per entry, 32 units in lock step (DESIGN.md "E-step, lock-step lanes").
//
//   L[c,k] = sum over the entries of cell c of cnt * T[g,k]          (scSplit:175-178; T = log2 P, log2(1-P))
//
// Same tables and the same integer arithmetic as estep_q.cuh (55-bit fixed point, 64-byte rows for K <= 9, 128-byte rows for
// K <= 17, sums exact in any order); what changes is who reads a row.  There a group of 4 (8) lanes shares an entry, so the
// entry's code has to reach every lane of the group (a shuffle: shared-memory pipe wavefronts), a warp walks 8 (4) runs of
// units of different lengths, and every unit is flushed by a divergent branch.  Here ONE lane reads the whole row of its own
// entry -- C = 4 (8) 16-byte chunks, one LDS.128 each, lane l starting at chunk l mod C so that the lanes of a quarter warp
// hit different banks -- and keeps all of the row's sums in registers.  The 32 lanes of a warp work on 32 different units
// (cell, table tile) of the SAME padded length, side by side ("task"): nothing is broadcast, nothing diverges, a task ends
// for all lanes at once.  The build (layout.cu: build_lockstep) sorts the units of a tile by length, lays the codes of a task
// out lane-interleaved (stream row = 32 words = one coalesced 128-byte load = two codes per lane), and, for 64-byte rows,
// pairs the units of lanes l and l + 4 so that their rows always differ in parity (the two halves of a 128-byte bank line):
// see k_lk_emit.  Because the sums are integers, a task may be cut anywhere: a CTA owns an equal-cost range of stream rows,
// its warps equal shares of it, and whoever finishes a piece of a unit adds its sums to the cell's 64-bit totals with
// RED.ADD.64 (integer atomics: order-free, hence still deterministic).  There are no per-(tile, cell) partial sums in memory
// and no combine kernel: k_bayes turns the totals into L = -sum * 2^-F, adds the rare entries with counts above Q_REP_MAX in
// fp64, and leaves the totals at zero for the next E-step.
#pragma once
#include "common.cuh"
#include "estep_q.cuh"

namespace scs {

constexpr int LK_THREADS = 512;      // 16 warps: a lane keeps 3 * C 64-bit sums and up to 8 * C loaded words in registers
constexpr int LK_RING = 8;           // stream rows a lane has requested ahead of the one it is adding

// 16-bit code of an entry: (byte offset of the lane's first chunk inside the tile) >> 2, i.e. row * ROWB / 4 with the
// lane's chunk rotation (l mod C) already XORed in by the build; chunk step j reads at ((code << 2) ^ (j << 4))
template <int C> struct LkCode {
    static constexpr int ROWB = C * 16;
    __host__ __device__ static uint32_t code(uint32_t row, int lane) { return (row * (uint32_t)(ROWB / 4)) ^ ((uint32_t)(lane & (C - 1)) << 2); }
};

struct LkStreamView {
    const uint32_t* words;       // [nrows][32]
    const int64_t* task_row;     // [ntasks + 1]
    const int32_t* task_cell;    // [ntasks][32]
    const int32_t* tile_task;    // [ntiles + 1]
    const int64_t* cta_row;      // [gridDim.x + 1]
    int ntiles;
    int TR;
    int64_t xrows;
};

template <int C>
__global__ void __launch_bounds__(LK_THREADS, 1)
k_estep_lk(const LkStreamView s, const uint64_t* __restrict__ Xq, int64_t xq_stride, unsigned long long* __restrict__ Lq, int64_t lq_stride,
           int K, int KP, const int32_t* __restrict__ done, const int32_t* __restrict__ qbad) {
    constexpr int ROWB = C * 16;
    extern __shared__ __align__(128) unsigned char smem[];
    const int TR = s.TR;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (((size_t)(TR + 2) * ROWB + 127) & ~(size_t)127));

    const int r = blockIdx.y;
    if (done && done[r]) return;
    if (qbad[r]) return;                               // this restart's table does not fit: k_bayes gathers its rows in fp64
    Xq += (int64_t)r * xq_stride;
    Lq += (int64_t)r * lq_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (int)(blockDim.x >> 5);
    const uint32_t smem_base = smem_u32(smem);
    const int rot = lane & (C - 1);

    const int64_t xa = __ldg(s.cta_row + blockIdx.x), xb = __ldg(s.cta_row + blockIdx.x + 1);
    if (xa >= xb) return;
    if (tid == 0) mbar_init(bar, 1);
    for (int k = tid; k < 2 * ROWB / 8; k += blockDim.x) reinterpret_cast<uint64_t*>(smem + (size_t)TR * ROWB)[k] = 0ull;
    __syncthreads();
    uint32_t phase = 0;

    int t = 0;
    for (int64_t x0 = xa; x0 < xb;) {
        while (t + 1 < s.ntiles && __ldg(s.task_row + __ldg(s.tile_task + t + 1)) <= x0) ++t;   // tile of stream row x0
        const int kt0 = __ldg(s.tile_task + t), kt1 = __ldg(s.tile_task + t + 1);
        const int64_t seg_end = min(xb, __ldg(s.task_row + kt1));
        __syncthreads();                                  // every warp is done reading the previous tile
        if (tid == 0) {
            const int64_t rows_t = min((int64_t)TR, s.xrows - (int64_t)t * TR);
            const uint32_t bytes = (uint32_t)(rows_t * ROWB);
            const char* src = reinterpret_cast<const char*>(Xq) + (int64_t)t * TR * ROWB;
            fence_proxy_async();
            mbar_expect_tx(bar, bytes);
            for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(smem + off, src + off, min(32768u, bytes - off), bar);
        }
        // this warp's share of the segment, and the task its first row belongs to (while the tile is on its way)
        const int64_t len = seg_end - x0;
        const int64_t wa = x0 + len * warp / nwarps, wb = x0 + len * (warp + 1) / nwarps;
        int k = kt0;
        if (wa < wb) {
            int lo = kt0, hi = kt1 - 1;                   // last task whose first row is <= wa
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (__ldg(s.task_row + mid) <= wa) lo = mid; else hi = mid - 1;
            }
            k = lo;
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        x0 = seg_end;
        if (wa >= wb) continue;

        int64_t tend = min(__ldg(s.task_row + k + 1), wb);          // where the lanes' current units end (or this warp's share does)
        int cell = __ldg(s.task_cell + (int64_t)k * 32 + lane);
        // the next task's end and cells are fetched a task ahead
        int64_t tend_n = __ldg(s.task_row + min(k + 2, kt1));
        int cell_n = (k + 1 < kt1) ? __ldg(s.task_cell + (int64_t)(k + 1) * 32 + lane) : -1;

        uint64_t a[C], b[C], pp[C];
#pragma unroll
        for (int j = 0; j < C; ++j) a[j] = b[j] = pp[j] = 0ull;
        auto flush = [&]() {
            uint64_t x = 0ull;
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const int c = j ^ rot;                     // the chunk this lane read at step j
                const uint64_t pa = pp[j] & 0xffffffull, pb = pp[j] >> 24;
                const uint64_t va = a[j] - (pa << 56), vb = b[j] - (pb << 56);   // the byte pair rode on top (mod 2^64)
                if (c < 4) x += (pa << (16 * c)) + (pb << (16 * c + 8));
                if (cell >= 0) {
                    SCS_ASSERT((int64_t)cell * KP + 2 * c + 1 < lq_stride || 2 * c >= K);
                    if (2 * c < K && va) atomicAdd(Lq + (int64_t)cell * KP + 2 * c, (unsigned long long)va);
                    if (2 * c + 1 < K && vb) atomicAdd(Lq + (int64_t)cell * KP + 2 * c + 1, (unsigned long long)vb);
                }
                a[j] = b[j] = pp[j] = 0ull;
            }
            // every lane holds all chunks of its rows, so the whole extra value is here
            if (cell >= 0 && 2 * C < K && x) atomicAdd(Lq + (int64_t)cell * KP + 2 * C, (unsigned long long)x);
        };

        const uint32_t* wp = s.words + wa * 32 + lane;
        uint32_t ring[LK_RING];
#pragma unroll
        for (int i = 0; i < LK_RING; ++i) ring[i] = (wa + i < wb) ? __ldcs(wp + (int64_t)i * 32) : 0u;
        for (int64_t x = wa; x < wb; x += LK_RING, wp += LK_RING * 32) {
#pragma unroll
            for (int i = 0; i < LK_RING; ++i) {
                if (x + i < wb) {                          // (warp-uniform)
                    const uint32_t w = ring[i];
                    if (x + i + LK_RING < wb) ring[i] = __ldcs(wp + (int64_t)(i + LK_RING) * 32);
                    const uint32_t a0 = smem_base + ((w & 0xffffu) << 2), a1 = smem_base + ((w >> 14) & 0x3fffcu);
                    SCS_ASSERT(((w & 0xffffu) << 2) < (uint32_t)(TR + 2) * ROWB && ((w >> 14) & 0x3fffcu) < (uint32_t)(TR + 2) * ROWB);
                    uint4 v0[C], v1[C];
#pragma unroll
                    for (int j = 0; j < C; ++j) v0[j] = lds_u32x4(a0 ^ (uint32_t)(j << 4));
#pragma unroll
                    for (int j = 0; j < C; ++j) v1[j] = lds_u32x4(a1 ^ (uint32_t)(j << 4));
#pragma unroll
                    for (int j = 0; j < C; ++j) {
                        a[j] += ((uint64_t)v0[j].x | ((uint64_t)v0[j].y << 32)) + ((uint64_t)v1[j].x | ((uint64_t)v1[j].y << 32));
                        b[j] += ((uint64_t)v0[j].z | ((uint64_t)v0[j].w << 32)) + ((uint64_t)v1[j].z | ((uint64_t)v1[j].w << 32));
                        pp[j] += (uint64_t)q_pieces(v0[j].y, v0[j].w) + (uint64_t)q_pieces(v1[j].y, v1[j].w);
                    }
                    if (x + i + 1 == tend) {               // (warp-uniform) the 32 units end here, or this warp's share does
                        flush();
                        ++k;
                        cell = cell_n;
                        tend = min(tend_n, wb);
                        if (x + i + 1 < wb) {
                            tend_n = __ldg(s.task_row + min(k + 2, kt1));
                            cell_n = (k + 1 < kt1) ? __ldg(s.task_cell + (int64_t)(k + 1) * 32 + lane) : -1;
                        }
                    }
                }
            }
        }
    }
}

}  // namespace scs
