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
// A table of more than 9 values per row (K = 10..17) is kept as two tables of 64-byte rows and read in two passes over the same
// stream (K and vbase: the values this pass adds and where they go in a cell's totals): two wavefronts per entry like one
// 128-byte row, but with tiles of twice the rows, hence half the units to flush, and the cheaper 64-byte kernel.
#pragma once
#include "common.cuh"
#include "estep_q.cuh"

namespace scs {

#ifndef SCS_LK_THREADS
#define SCS_LK_THREADS 512
#endif
#ifndef SCS_LK_RING
#define SCS_LK_RING 8
#endif
// threads per CTA: a lane keeps 3 * C 64-bit sums, the ring and up to 8 loaded 16-byte chunks in registers -- 16 warps of 128
// registers for 64-byte rows, 12 warps of 168 for 128-byte rows (at 128 the ring of the wide kernel spilled while in flight)
#ifndef SCS_LK_THREADS8
#define SCS_LK_THREADS8 384
#endif
__host__ __device__ constexpr int lk_threads(int C) { return C <= 4 ? SCS_LK_THREADS : SCS_LK_THREADS8; }
constexpr int LK_RING = SCS_LK_RING;         // stream rows a lane has requested ahead of the one it is adding

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
    const int32_t* cta_tile;     // [gridDim.x] tile of the first row of each range
    int ntiles;
    int TR;
    int64_t xrows;
};

#ifdef SCS_LK_TRACE     // (diagnostic build: per-CTA times of one launch on stdout)
__device__ int lk_trace_launch = 0;
__device__ __forceinline__ unsigned long long lk_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

template <int C>
__global__ void __launch_bounds__(lk_threads(C), 1)
k_estep_lk(const LkStreamView s, const uint64_t* __restrict__ Xq, int64_t xq_stride, unsigned long long* __restrict__ Lq, int64_t lq_stride,
           int K, int KP, int vbase, const int32_t* __restrict__ done, const int32_t* __restrict__ qbad) {
    // rows a lane keeps requested ahead = rows per unrolled block: 8 for 64-byte rows; 4 for 128-byte rows, whose block of 8
    // (1200 instructions) ran out of the instruction cache: `no_instructions` was the top stall of the first capture
    constexpr int ROWB = C * 16, RING = C <= 4 ? LK_RING : LK_RING / 2;
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

#ifdef SCS_LK_TRACE
    const unsigned long long tr_t0 = lk_now();
    unsigned long long tr_t1 = 0;
    int tr_segs = 0;
#endif
    const int64_t xa = __ldg(s.cta_row + blockIdx.x), xb = __ldg(s.cta_row + blockIdx.x + 1);
    int t = __ldg(s.cta_tile + blockIdx.x);              // (walking the tiles from 0 cost the last CTAs 16 dependent round trips)
    if (xa >= xb) return;
    auto load_tile = [&](int tile) {                     // (thread 0) TMA bulk copies of one table tile, completing on `bar`
        const int64_t rows_t = min((int64_t)TR, s.xrows - (int64_t)tile * TR);
        const uint32_t bytes = (uint32_t)(rows_t * ROWB);
        const char* src = reinterpret_cast<const char*>(Xq) + (int64_t)tile * TR * ROWB;
        fence_proxy_async();
        mbar_expect_tx(bar, bytes);
        for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(smem + off, src + off, min(32768u, bytes - off), bar);
    };
    // The first tile is requested before anything else is looked up: a CTA cannot add a row before its 226 KB have arrived
    // (launch + look-ups + loads alone take 13 us of the kernel's 82 at config 3), so the plan look-ups, the binary search for
    // the warp's first task, its cells and its first stream words all go on while the tile is on its way.
    // (Tried: clusters of two CTAs whose ranges start in the same tile loading it once, each CTA one half, multicast into both.
    // No gain, 83.6 against 83.0 us: the L2 already serves neighbouring CTAs' requests for the same lines from one read.)
    if (tid == 0) {
        mbar_init(bar, 1);
        load_tile(t);
    }
    for (int k = tid; k < 2 * ROWB / 8; k += blockDim.x) reinterpret_cast<uint64_t*>(smem + (size_t)TR * ROWB)[k] = 0ull;
    __syncthreads();
    uint32_t phase = 0;
    bool first = true;

    for (int64_t x0 = xa; x0 < xb;) {
        while (t + 1 < s.ntiles && __ldg(s.task_row + __ldg(s.tile_task + t + 1)) <= x0) ++t;   // tile of stream row x0
        const int kt0 = __ldg(s.tile_task + t), kt1 = __ldg(s.tile_task + t + 1);
        const int64_t seg_end = min(xb, __ldg(s.task_row + kt1));
        if (!first) {
            __syncthreads();                              // every warp is done reading the previous tile
            if (tid == 0) load_tile(t);
        }
        first = false;
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
        // stream rows are counted from the start of this warp's share
        const int nrow = (int)(wb - wa);
        int stop = 0, cell = -1, stop_n = 0, cell_n = -1;
        const uint32_t* wp = s.words + wa * 32 + lane;
        uint32_t ring[RING];
        if (wa < wb) {
            stop = (int)(min(__ldg(s.task_row + k + 1), wb) - wa);    // where the lanes' current units end (or the share does)
            cell = __ldg(s.task_cell + (int64_t)k * 32 + lane);
            // the next task's end and cells are fetched a task ahead
            stop_n = (int)(min(__ldg(s.task_row + min(k + 2, kt1)), wb) - wa);
            cell_n = (k + 1 < kt1) ? __ldg(s.task_cell + (int64_t)(k + 1) * 32 + lane) : -1;
#pragma unroll
            for (int i = 0; i < RING; ++i) ring[i] = __ldcs(wp + i * 32);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        x0 = seg_end;
#ifdef SCS_LK_TRACE
        if (!tr_segs++) tr_t1 = lk_now();
#endif
#ifdef SCS_LK_LOAD_ONLY    // (diagnostic build: launch, plan look-ups and the table tile loads, no rows; the results are wrong)
        continue;
#endif
        if (wa >= wb) continue;

        uint64_t a[C], b[C], pp[C];
#pragma unroll
        for (int j = 0; j < C; ++j) a[j] = b[j] = pp[j] = 0ull;
        // Add the lanes' sums to the totals of their cells.  A RED of 32 lanes to 32 different cells is 32 passes of the memory
        // pipe (the same pipe the table reads need: nine such REDs per unit cost 18 us per E-step at config 3), so the sums are
        // first transposed inside each group of 8 lanes (butterfly of shuffles: lane i ends up with slot i of each of the 8
        // cells): then one RED covers 8 consecutive words of 4 cells, 4..8 lines instead of 32.  Slot 2j (+1) of a lane is the
        // first (second) value of the chunk it read at step j, i.e. chunk j ^ rot of the row.
        auto flush = [&]() {
#ifdef SCS_LK_NO_RED     // (diagnostic build: what the kernel costs without its atomics; the results are wrong)
            const bool live = cell == -2;
#else
            const bool live = cell >= 0;
#endif
            uint64_t x = 0ull, R[2 * C];
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const int c = j ^ rot;                     // the chunk this lane read at step j
                const uint64_t pa = pp[j] & 0xffffffull, pb = pp[j] >> 24;
                R[2 * j] = a[j] - (pa << 56);              // the byte pair rode on top (mod 2^64)
                R[2 * j + 1] = b[j] - (pb << 56);
                if (c < 4) x += (pa << (16 * c)) + (pb << (16 * c + 8));
                a[j] = b[j] = pp[j] = 0ull;
            }
            // every lane holds all chunks of its rows, so the whole extra value is here
            if (live && 2 * C < K && x) atomicAdd(Lq + (int64_t)cell * KP + vbase + 2 * C, (unsigned long long)x);
            const int oi = lane & 7;
#pragma unroll
            for (int h = 0; h < 2 * C / 8; ++h) {
#pragma unroll
                for (int d = 4; d; d >>= 1) {
                    const bool up = (oi & d) != 0;
#pragma unroll
                    for (int sl = 0; sl < 8; ++sl) {
                        if (sl & d) continue;
                        uint64_t& lo = R[8 * h + sl];
                        uint64_t& hi = R[8 * h + (sl | d)];
                        const uint64_t send = up ? lo : hi;
                        const uint64_t recv = __shfl_xor_sync(0xffffffffu, send, d);
                        if (up) lo = recv; else hi = recv;
                    }
                }
            }
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {               // octet lane c8's cell; this lane holds its slots oi (+ 8 h)
                const int cell_c = __shfl_sync(0xffffffffu, cell, (lane & ~7) | c8);
                const int rot_c = c8 & (C - 1);
#pragma unroll
                for (int h = 0; h < 2 * C / 8; ++h) {
                    const int slot = oi + 8 * h;
                    const int v = 2 * ((slot >> 1) ^ rot_c) + (slot & 1);
                    const uint64_t val = R[8 * h + c8];
#ifdef SCS_LK_NO_RED
                    if (cell_c == -2 && v < K && val) atomicAdd(Lq + (int64_t)cell_c * KP + vbase + v, (unsigned long long)val);
#else
                    SCS_ASSERT(cell_c < 0 || (int64_t)cell_c * KP + vbase + v < lq_stride || v >= K);
                    if (cell_c >= 0 && v < K && val) atomicAdd(Lq + (int64_t)cell_c * KP + vbase + v, (unsigned long long)val);
#endif
                }
            }
        };

        // the two codes of one stream word: 2 * C LDS.128, integer adds
        auto add_codes = [&](uint32_t a0, uint32_t a1) {
            if constexpr (C <= 4) {                      // both rows requested before the adds: 8 LDS.128 in flight, three-input adds
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
            } else {                                     // a 128-byte row is 8 LDS.128 by itself: one row at a time (registers)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t ah = h ? a1 : a0;
                    uint4 v[C];
#pragma unroll
                    for (int j = 0; j < C; ++j) v[j] = lds_u32x4(ah ^ (uint32_t)(j << 4));
#pragma unroll
                    for (int j = 0; j < C; ++j) {
                        a[j] += (uint64_t)v[j].x | ((uint64_t)v[j].y << 32);
                        b[j] += (uint64_t)v[j].z | ((uint64_t)v[j].w << 32);
                        pp[j] += (uint64_t)q_pieces(v[j].y, v[j].w);
                    }
                }
            }
        };
        auto addr0 = [&](uint32_t w) { return smem_base + ((w & 0xffffu) << 2); };
        auto addr1 = [&](uint32_t w) { return smem_base + ((w >> 14) & 0x3fffcu); };
        auto add_row = [&](uint32_t w) {
            SCS_ASSERT(((w & 0xffffu) << 2) < (uint32_t)(TR + 2) * ROWB && ((w >> 14) & 0x3fffcu) < (uint32_t)(TR + 2) * ROWB);
            add_codes(addr0(w), addr1(w));
        };
        // A lane keeps the words of RING rows in registers, row y (counted from the share's start) in ring[y % RING], each
        // requested RING rows before it is used.  The requests are unconditional -- the stream ends in 2 * LK_RING rows of
        // slack -- and land in fixed registers: nothing is ever moved while in flight.  (A predicated request made the compiler
        // load into a temporary and MOVE it into the ring at once: the kernel then sat in long-scoreboard stalls, 123 us.)
        auto boundary = [&](int x) {                     // the 32 units end here, or this warp's share does
            flush();
            ++k;
            cell = cell_n;
            stop = stop_n;
            if (x < nrow) {
                stop_n = (int)(min(__ldg(s.task_row + min(k + 2, kt1)), wb) - wa);
                cell_n = (k + 1 < kt1) ? __ldg(s.task_cell + (int64_t)(k + 1) * 32 + lane) : -1;
            }
        };
        int x = 0;                                       // next row to add
        for (int xb = 0; xb < nrow; xb += RING, wp += RING * 32) {
            if (stop >= xb + RING) {                  // the block lies inside the current piece
#pragma unroll
                for (int i = 0; i < RING; ++i) {
                    // the word's two addresses are taken out first and named as inputs of the request, so that the request can
                    // land in the register the word leaves (otherwise the ring is copied at the loop's end -- while in flight)
                    const uint32_t a0 = addr0(ring[i]), a1 = addr1(ring[i]);
                    SCS_ASSERT(a0 - smem_base < (uint32_t)(TR + 2) * ROWB && a1 - smem_base < (uint32_t)(TR + 2) * ROWB);
                    asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(ring[i]) : "l"(wp + (i + RING) * 32), "r"(a0), "r"(a1));
                    add_codes(a0, a1);
                }
                x = xb + RING;
                if (x == stop) boundary(x);
            } else {                                     // pieces end inside the block: rows [x, hi) at a time
                const int bend = min(xb + RING, nrow);
                while (x < bend) {
                    const int hi = min(stop, bend);
#pragma unroll
                    for (int i = 0; i < RING; ++i)
                        if (xb + i >= x && xb + i < hi) add_row(ring[i]);      // (warp-uniform)
                    x = hi;
                    if (x == stop) boundary(x);
                }
#pragma unroll
                for (int i = 0; i < RING; ++i) ring[i] = __ldcs(wp + (i + RING) * 32);
            }
        }
    }
#ifdef SCS_LK_TRACE
    __syncthreads();
    if (tid == 0) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        const int launch = atomicAdd(&lk_trace_launch, 0);
        if (launch >= 30 * (int)gridDim.x && launch < 31 * (int)gridDim.x)
            printf("lktrace cta %d sm %u rows %lld segs %d t0 %llu tile %llu end %llu\n", (int)blockIdx.x, smid, (long long)(xb - xa), tr_segs,
                   tr_t0, tr_t1 - tr_t0, lk_now() - tr_t0);
        atomicAdd(&lk_trace_launch, 1);
    }
#endif
}

}  // namespace scs
