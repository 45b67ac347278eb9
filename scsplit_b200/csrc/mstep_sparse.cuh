// M-step fast path for saturated posteriors (DESIGN.md "sparse M-step").
//
// When nearly every posterior row packs into one 32-bit word (state, 1 - weight; see sparsify_group in em_kernels.cuh) the
// M-step sums S[g, k] = sum_{entries of pseudo-row g} W[cell, k] need 4 bytes of posterior per entry instead of a KP*8-byte
// row.  The packed words of ALL cells are staged in shared memory with TMA bulk copies (4 bytes x N), a group of 16 lanes
// walks one pseudo-row of the SNV-major stream (2 rows per warp, 16-byte code vectors), and every lane owns K fp64
// accumulators in shared memory laid out [state][lane] -- the bank is the lane, so the data-dependent read-modify-write
// is conflict-free.  A fixed reduce-scatter butterfly over the 16 lanes finishes a row.  Cells that are not saturated are
// added as full rows from global memory.  Rows are jobs drawn from a ticket counter in descending-length order.  The
// kernel runs only when the dense-path cell count published by k_bayes is below the gate threshold; otherwise the tiled
// SpMM (gated the other way) does the M-step.  Same bits on every run.
#pragma once
#include "em_kernels.cuh"
#include "spmm_tiled.cuh"

namespace scs {

constexpr int SPARSE_THREADS = 1024;                // most threads per CTA; fewer when the accumulators would not fit
constexpr int SPARSE_GS = 16;                      // lanes per pseudo-row

// packed words of all cells + one SPARSE_NOP word, rounded to the 128-byte alignment of the mbarrier behind it
__host__ __device__ inline size_t sparse_tab_bytes(int64_t h_stride) { return (((size_t)h_stride * 4 + 4) + 127) & ~(size_t)127; }
inline size_t sparse_smem_bytes(int64_t h_stride, int KP, int threads) {
    const size_t acc = (size_t)threads * KP * 8;                         // [warp][state][lane] fp64
    return acc + sparse_tab_bytes(h_stride) + 16;
}
// threads per CTA of the sparse M-step for N cells and KP columns: as many warps as fit next to the N packed words (every
// lane owns KP fp64 slots), at least 8 warps; 0 = the pool does not fit in one piece
inline int sparse_threads_for(int64_t N, int KP) {
    for (int t = SPARSE_THREADS; t >= 256; t -= 128)
        if (sparse_smem_bytes((N + 3) / 4 * 4, KP, t) <= (size_t)TILE_BYTES + 1024) return t;
    return 0;
}
// How the sparse M-step covers the cells: in one piece when their packed words fit shared memory next to the accumulators,
// else in `nblocks` blocks of `block_cells` cells (a multiple of 4), one launch per block, each adding the block's share of
// every pseudo-row to S (k_mstep_sparse, CellBlock).  threads == 0: no configuration fits (KP too wide).
struct SparsePlan {
    int threads = 0;
    int nblocks = 0;
    int64_t block_cells = 0;
};
inline SparsePlan sparse_plan_for(int64_t N, int KP) {
    SparsePlan p;
    p.threads = sparse_threads_for(N, KP);
    if (p.threads) { p.nblocks = 1; p.block_cells = (N + 3) / 4 * 4; return p; }
    // blocks: the most threads that still leave a block of 24k cells or more (every block walks all pseudo-rows again, so few
    // large blocks beat many small ones; 24k is what one of 8 GPUs holds of BASELINE config 4), else of 8k or more
    for (int pass = 0; pass < 2; ++pass)
    for (int t = SPARSE_THREADS; t >= 256; t -= 128) {
        const int64_t room = (int64_t)TILE_BYTES + 1024 - 16 - 132 - (int64_t)t * KP * 8;
        const int64_t cap = room > 0 ? room / 4 / 4 * 4 : 0;
        if (cap >= (pass == 0 ? 24576 : 8192)) {
            p.threads = t;
            p.nblocks = (int)((N + cap - 1) / cap);
            p.block_cells = ((N + p.nblocks - 1) / p.nblocks + 3) / 4 * 4;     // equal blocks, each a multiple of 4 cells
            return p;
        }
    }
    return p;
}
// One block of cells of a blocked sparse M-step launch: cells [lo, lo + n) (n a multiple of 4; the last block may reach past
// N into the padding of hq), the entries of pseudo-row g inside the block are lcode[lo_ptr[g] .. hi_ptr[g]) (nullptr: the
// whole row), accumulate = add to S instead of overwriting it (every block but the first).
struct CellBlock {
    int64_t lo = 0, n = 0;
    const int64_t* lo_ptr = nullptr;
    const int64_t* hi_ptr = nullptr;
    int accumulate = 0;
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {   // read-only table: free to be scheduled early
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// Reduce-scatter butterfly over the SPARSE_GS lanes of a row group: every round a lane hands one half of its values to
// its partner and adds the partner's copy of the half it keeps, so the 4 rounds move CNT, CNT/2, CNT/4, ... values instead
// of CNT each (22 instead of 80 shuffles for K = 9).  On return out[0..nval) are finished sums of states base, base+1, ...
// (odd halvings pad with zeros, nval counts the real ones); the tree (hence the rounding) of every state is fixed.
template <int CNT, int O>
struct GroupReduce {
    static constexpr int HALF = (CNT + 1) / 2;
    static constexpr int LEFT = GroupReduce<HALF, O / 2>::LEFT;
    __device__ __forceinline__ static void run(double (&v)[CNT], double (&out)[LEFT], int li, int& base, int& nval) {
        const bool hi = (li & O) != 0;
        double w[HALF];
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
            const double lo_v = v[i], hi_v = (i + HALF < CNT) ? v[i + HALF] : 0.0;
            const double keep = hi ? hi_v : lo_v, send = hi ? lo_v : hi_v;
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
        }
        base += hi ? HALF : 0;
        nval = hi ? max(nval - HALF, 0) : min(nval, HALF);
        GroupReduce<HALF, O / 2>::run(w, out, li, base, nval);
    }
};
template <int CNT>
struct GroupReduce<CNT, 0> {
    static constexpr int LEFT = CNT;
    __device__ __forceinline__ static void run(double (&v)[CNT], double (&out)[CNT], int, int&, int&) {
#pragma unroll
        for (int i = 0; i < CNT; ++i) out[i] = v[i];
    }
};

// acc_base: shared address of this lane's slot for state 0; slots of consecutive states are 256 bytes apart
__device__ __forceinline__ void sparse_rmw(uint32_t acc_base, uint32_t q, double cnt) {
    const double w = 1.0 - (double)__uint_as_float(q & ~31u);            // the exact posterior weight
    const uint32_t a = acc_base + (q & 31u) * 256u;
    sts_f64(a, lds_f64(a) + cnt * w);
}
template <int KP>
__device__ __forceinline__ void dense_add(uint32_t acc_base, int K, int64_t cell, const double* __restrict__ W, double cnt) {
    const double* wr = W + cell * KP;                                    // unsaturated cell: full row from global memory
    for (int k = 0; k < K; ++k) {
        const uint32_t a = acc_base + (uint32_t)k * 256u;
        sts_f64(a, lds_f64(a) + cnt * __ldg(wr + k));
    }
}

// grid = (CTAs, R).  Row pairs are handed out through sched[r][0] (an integer ticket counter), longest rows first, so the
// SMs finish together whatever the row-length profile is; the last CTA to leave resets the tickets for the next launch.
// Which warp sums a row never changes the order inside the row, so the result is the same bits on every run.
template <int KP>
__global__ void __launch_bounds__(SPARSE_THREADS, 1)
k_mstep_sparse(int64_t nrows, int64_t njobs, int64_t N, int K, const int32_t* __restrict__ order, const int64_t* __restrict__ lptr,
               const uint32_t* __restrict__ lcode,
               const int64_t* __restrict__ hptr, const uint32_t* __restrict__ hcode, const int32_t* __restrict__ hcnt,
               const uint32_t* __restrict__ hq, int64_t h_stride, const double* __restrict__ W, double* __restrict__ S,
               int64_t s_stride, const Gate gate, const int32_t* __restrict__ done, int32_t* __restrict__ sched, const CellBlock blk) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int r = blockIdx.y;
    if (done && done[r]) return;
    if (gate.dense && !gate_takes_sparse(gate, r)) return;   // too many unsaturated cells: the tiled SpMM runs instead
    const size_t ACC_BYTES = (size_t)blockDim.x * KP * 8;
    unsigned char* s_tab = smem + ACC_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_tab + sparse_tab_bytes(blk.n));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_proxy_async();
        const uint32_t bytes = (uint32_t)(blk.n * 4);
        mbar_expect_tx(bar, bytes);
        const char* src = reinterpret_cast<const char*>(hq + (int64_t)r * h_stride + blk.lo);
        for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(s_tab + off, src + off, min(32768u, bytes - off), bar);
        // word blk.n of the table (past the copied part): where the codes of lanes beyond a row's end point
        reinterpret_cast<uint32_t*>(s_tab)[blk.n] = SPARSE_NOP;
    }
    // this lane's accumulator slots: [warp][state][lane], zeroed once; every row leaves them zero again
    uint32_t acc_base = smem_u32(smem) + (uint32_t)warp * (KP * 256u) + (uint32_t)lane * 8u;
    uint32_t tab = smem_u32(s_tab) - (uint32_t)(blk.lo * 4);         // indexed by the GLOBAL cell number (wraps below the table; never dereferenced there)
    asm volatile("" : "+r"(acc_base), "+r"(tab));                    // keep both in registers (no re-derivation per entry)
#pragma unroll
    for (int k = 0; k < KP; ++k) sts_f64(acc_base + k * 256u, 0.0);
    __syncthreads();
    mbar_wait(bar, 0);
    W += (int64_t)r * N * KP;
    S += (int64_t)r * s_stride;
    sched += 2 * r;
    const int li = lane & (SPARSE_GS - 1), g = lane / SPARSE_GS;
    constexpr int GPW = 32 / SPARSE_GS;
#ifndef SCS_SPARSE_UV
#define SCS_SPARSE_UV 2
#endif
    constexpr int UV = SCS_SPARSE_UV;
    const int npairs = (int)((njobs + GPW - 1) / GPW);    // `order` lists the njobs pseudo-rows of this launch
    const uint32_t nop_code = (uint32_t)(blk.lo + blk.n) << 1;       // code whose table word is SPARSE_NOP
    // A row (pair) is a short job (a few hundred entries), so the chain ticket -> row bounds -> first codes is software-
    // pipelined across rows: the ticket is drawn two rows ahead, the bounds are loaded one row ahead, and the first code
    // vectors of the next row are requested before the current row's reduction tree runs.
    struct RowJob {
        int64_t row;      // pseudo-row of this lane's group (>= nrows: none)
        int64_t lb;       // first entry
        int n;            // entries
    };
    auto draw = [&]() -> int { return lane == 0 ? atomicAdd(sched, 1) : 0; };           // raw: only lane 0 holds the ticket
    auto job_of = [&](int ticket) -> RowJob {
        // `order` lists the pseudo-rows by descending length: the GPW rows of a warp are neighbours in that list (equal
        // lengths, so the groups of a warp finish together) and the longest rows are drawn first
        RowJob j;
        const int64_t idx = (int64_t)ticket * GPW + g;
        j.row = (ticket < npairs && idx < njobs) ? (int64_t)__ldg(order + idx) : nrows;
        const bool act = j.row < nrows;
        j.lb = act ? __ldg((blk.lo_ptr ? blk.lo_ptr : lptr) + j.row) : 0;
        j.n = act ? (int)(__ldg(blk.hi_ptr ? blk.hi_ptr + j.row : lptr + j.row + 1) - j.lb) : 0;
        return j;
    };
    // The codes are read as 16-byte vectors from the aligned address below the row start (`a` elements early), UV vectors
    // per lane per step, and the vectors of the NEXT step are fetched before this step's accumulator updates.  The step
    // itself is branch-free: elements outside the row and unsaturated cells both add 0.0 (SPARSE_NOP / SPARSE_DENSE decode
    // to weight 0); unsaturated cells are only flagged there and added from their full rows afterwards.
    auto load_vec = [&](const RowJob& j, int v) -> uint4 {             // raw: may hold elements outside the row
        const int a = (int)(j.lb & 3), m = a + j.n;
        uint4 c = make_uint4(nop_code, nop_code, nop_code, nop_code);
        if (4 * v < m) c = __ldcs(reinterpret_cast<const uint4*>(lcode + (j.lb - a)) + v);
        return c;
    };
    // blank the elements of vector v outside [a, m): applied when the vector is USED, so that nothing waits on the load
    auto trim_vec = [&](uint4& c, int v, int a, int m) {
        const int k0 = 4 * v;
        if (k0 < a || k0 + 4 > m) {                                    // first / last vector of the row
            if (k0 < a || k0 >= m) c.x = nop_code;
            if (k0 + 1 < a || k0 + 1 >= m) c.y = nop_code;
            if (k0 + 2 < a || k0 + 2 >= m) c.z = nop_code;
            if (k0 + 3 < a || k0 + 3 >= m) c.w = nop_code;
        }
    };
    int t_cur = __shfl_sync(0xffffffffu, draw(), 0);
    int t_nxt = __shfl_sync(0xffffffffu, draw(), 0);
    RowJob job = job_of(t_cur);
    uint4 cur[UV], nxt[UV];
#pragma unroll
    for (int j = 0; j < UV; ++j) cur[j] = load_vec(job, li + j * SPARSE_GS);
    while (t_cur < npairs) {
        const int raw_after = draw();                       // ticket of the row after next (in flight during this row)
        const RowJob job_nxt = job_of(t_nxt);               // bounds of the next row (in flight during this row)
        const int a = (int)(job.lb & 3), m = a + job.n, nv = (m + 3) >> 2;
        // unsaturated cells met by this lane: bit (EB * steps_back + j) for element j of a step (EB = 4 * UV bits per step,
        // newest step in the low bits); `lost` collects what a row longer than 64 / EB steps per lane shifts out
        constexpr int EB = 4 * UV;
        uint64_t dmask = 0;
        uint32_t lost = 0;
        int steps = 0;
        for (int v = li; v < nv; v += UV * SPARSE_GS) {
#pragma unroll
            for (int j = 0; j < UV; ++j) nxt[j] = load_vec(job, v + (UV + j) * SPARSE_GS);
            uint32_t q[4 * UV];
#pragma unroll
            for (int j = 0; j < UV; ++j) trim_vec(cur[j], v + j * SPARSE_GS, a, m);
#pragma unroll
            for (int j = 0; j < UV; ++j) {  // every code (cell << 1 | dup) addresses the staged block: lo <= cell <= lo + n (the NOP word)
                const uint32_t c_lo = (uint32_t)blk.lo, c_hi = (uint32_t)(blk.lo + blk.n);
                (void)c_lo; (void)c_hi;                                  // (only the bounds-checked build reads them)
                SCS_ASSERT((cur[j].x >> 1) >= c_lo && (cur[j].x >> 1) <= c_hi && (cur[j].y >> 1) >= c_lo && (cur[j].y >> 1) <= c_hi &&
                           (cur[j].z >> 1) >= c_lo && (cur[j].z >> 1) <= c_hi && (cur[j].w >> 1) >= c_lo && (cur[j].w >> 1) <= c_hi);
            }
#pragma unroll
            for (int j = 0; j < UV; ++j) {
                q[4 * j + 0] = lds_u32(tab + ((cur[j].x & ~1u) << 1));
                q[4 * j + 1] = lds_u32(tab + ((cur[j].y & ~1u) << 1));
                q[4 * j + 2] = lds_u32(tab + ((cur[j].z & ~1u) << 1));
                q[4 * j + 3] = lds_u32(tab + ((cur[j].w & ~1u) << 1));
            }
            uint32_t m8 = 0;
#pragma unroll
            for (int j = 0; j < 4 * UV; ++j) {
                m8 |= (q[j] == SPARSE_DENSE) ? (1u << j) : 0u;
                sparse_rmw(acc_base, q[j], 1.0);
            }
            lost |= (uint32_t)(dmask >> (64 - EB));
            dmask = (dmask << EB) | m8;
            ++steps;
#pragma unroll
            for (int j = 0; j < UV; ++j) cur[j] = nxt[j];
        }
        // first codes of the next row: requested now, used after this row's tail work
#pragma unroll
        for (int j = 0; j < UV; ++j) cur[j] = load_vec(job_nxt, li + j * SPARSE_GS);
        const uint32_t* code = lcode + job.lb;
        if (lost) {                                                    // very long row: walk this lane's own elements again
            for (int v = li; v < nv; v += SPARSE_GS)
                for (int k = max(4 * v, a); k < min(4 * v + 4, m); ++k) {
                    const uint32_t c = __ldg(code + (k - a)) >> 1;
                    if (lds_u32(tab + c * 4u) == SPARSE_DENSE) dense_add<KP>(acc_base, K, (int64_t)c, W, 1.0);
                }
        } else {
            while (dmask) {                                            // rare: exactly the flagged elements, full rows
                const int b = __ffsll((long long)dmask) - 1;
                dmask &= dmask - 1;
                const int s_idx = steps - 1 - b / EB, j = b % EB;
                const int k = 4 * (li + (s_idx * UV + (j >> 2)) * SPARSE_GS) + (j & 3);
                dense_add<KP>(acc_base, K, (int64_t)(__ldg(code + (k - a)) >> 1), W, 1.0);
            }
        }
        const bool act = job.row < nrows;
        const int64_t hb = act ? hptr[job.row] : 0, he = act ? hptr[job.row + 1] : 0;
        for (int64_t h = hb + li; h < he; h += SPARSE_GS) {
            const uint32_t c = __ldg(hcode + h) >> 1;
            if ((int64_t)c < blk.lo || (int64_t)c >= blk.lo + blk.n) continue;      // (a count>1 entry of another block of cells)
            const uint32_t q = lds_u32(tab + c * 4u);
            const double cnt = (double)__ldg(hcnt + h);
            if (q != SPARSE_DENSE) sparse_rmw(acc_base, q, cnt);
            else dense_add<KP>(acc_base, K, (int64_t)c, W, cnt);
        }
        __syncwarp();
        // fixed reduce-scatter tree over the lanes of the row group; slots are reset on the way
        {
            using GR = GroupReduce<KP, SPARSE_GS / 2>;
            double v[KP], out[GR::LEFT];
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                v[k] = lds_f64(acc_base + k * 256u);
                sts_f64(acc_base + k * 256u, 0.0);
            }
            int base = 0, nval = KP;
            GR::run(v, out, li, base, nval);
            if (act) {
                SCS_ASSERT(job.row >= 0 && job.row < nrows);
                double* y = S + job.row * KP;
#pragma unroll
                for (int i = 0; i < GR::LEFT; ++i)
                    if (i < nval) y[base + i] = blk.accumulate ? y[base + i] + out[i] : out[i];   // (blocks run as successive launches)
            }
        }
        job = job_nxt;
        t_cur = t_nxt;
        t_nxt = __shfl_sync(0xffffffffu, raw_after, 0);
    }
    __syncthreads();
    if (tid == 0) {
        if (atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) {         // every CTA has drawn its last ticket: reset for the next launch
            sched[0] = 0;
            sched[1] = 0;
        }
    }
}

}  // namespace scs
