// M-step fast path for saturated posteriors (DESIGN.md "sparse M-step").
//
// When nearly every posterior row packs into one 32-bit word (state, 1 - weight; see sparsify_row in em_kernels.cuh) the
// M-step sums S[g, k] = sum_{entries of pseudo-row g} W[cell, k] need 4 bytes of posterior per entry instead of a KP*8-byte
// row.  The packed words of ALL cells are staged in shared memory with TMA bulk copies (4 bytes x N), a group of 8 lanes
// walks one pseudo-row of the SNV-major stream (2 rows per warp), and every lane owns K fp64 accumulators in shared
// memory laid out [state][lane] -- the bank is the lane, so the data-dependent read-modify-write is conflict-free.  A
// fixed xor-shuffle tree over the 8 lanes finishes a row.  Cells that are not saturated are gathered as full rows from
// global memory.  The kernel runs only when the dense-path cell count published by k_reduce_stats is below the gate
// threshold; otherwise the tiled SpMM (gated the other way) does the M-step.  Same bits on every run.
#pragma once
#include "em_kernels.cuh"
#include "spmm_tiled.cuh"

namespace scs {

constexpr int SPARSE_THREADS = 1024;
constexpr int SPARSE_GS = 16;                      // lanes per pseudo-row (>= KP/2: the row's column pairs are stored by lanes 0..KP/2-1)
static_assert(SPARSE_GS * 2 >= SCS_MAX_K, "a row group must cover all column pairs");

inline size_t sparse_smem_bytes(int64_t h_stride, int KP) {
    const size_t acc = (size_t)SPARSE_THREADS * KP * 8;                  // [warp][state][lane] fp64
    const size_t tab = (((size_t)h_stride * 4) + 127) & ~(size_t)127;
    return acc + tab + 16;
}
inline bool sparse_fits(int64_t N, int KP) { return sparse_smem_bytes((N + 3) / 4 * 4, KP) <= (size_t)TILE_BYTES + 1024; }

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

// acc_base: shared address of this lane's slot for state 0; slots of consecutive states are 256 bytes apart
template <int KP>
__device__ __forceinline__ void sparse_add(uint32_t acc_base, int K, uint32_t q, int64_t cell, const double* __restrict__ W, double cnt) {
    if (q != SPARSE_DENSE) {
        const double w = 1.0 - (double)__uint_as_float(q & ~31u);        // the exact posterior weight
        const uint32_t a = acc_base + (q & 31u) * 256u;
        sts_f64(a, lds_f64(a) + cnt * w);
    } else {                                                             // unsaturated cell: full row from global memory
        const double* wr = W + cell * KP;
        for (int k = 0; k < K; ++k) {
            const uint32_t a = acc_base + (uint32_t)k * 256u;
            sts_f64(a, lds_f64(a) + cnt * __ldg(wr + k));
        }
    }
}

template <int KP>
__global__ void __launch_bounds__(SPARSE_THREADS, 1)
k_mstep_sparse(int64_t nrows, int64_t N, int K, const int64_t* __restrict__ lptr, const uint32_t* __restrict__ lcode,
               const int64_t* __restrict__ hptr, const uint32_t* __restrict__ hcode, const int32_t* __restrict__ hcnt,
               const uint32_t* __restrict__ hq, int64_t h_stride, const double* __restrict__ W, double* __restrict__ S,
               int64_t s_stride, const double* __restrict__ gate, int64_t gate_stride, double gate_thr,
               const int32_t* __restrict__ done) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int r = blockIdx.y;
    if (done && done[r]) return;
    if (gate && gate[(int64_t)r * gate_stride] > gate_thr) return;   // too many unsaturated cells: the tiled SpMM runs instead
    constexpr size_t ACC_BYTES = (size_t)SPARSE_THREADS * KP * 8;
    unsigned char* s_tab = smem + ACC_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_tab + ((((size_t)h_stride * 4) + 127) & ~(size_t)127));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_proxy_async();
        const uint32_t bytes = (uint32_t)(h_stride * 4);
        mbar_expect_tx(bar, bytes);
        const char* src = reinterpret_cast<const char*>(hq + (int64_t)r * h_stride);
        for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(s_tab + off, src + off, min(32768u, bytes - off), bar);
    }
    // this lane's accumulator slots: [warp][state][lane], zeroed once; every row leaves them zero again
    const uint32_t acc_base = smem_u32(smem) + (uint32_t)warp * (KP * 256u) + (uint32_t)lane * 8u;
#pragma unroll
    for (int k = 0; k < KP; ++k) sts_f64(acc_base + k * 256u, 0.0);
    __syncthreads();
    mbar_wait(bar, 0);
    const uint32_t tab = smem_u32(s_tab);
    W += (int64_t)r * N * KP;
    S += (int64_t)r * s_stride;
    const int li = lane & (SPARSE_GS - 1), g = lane / SPARSE_GS;
    constexpr int GPW = 32 / SPARSE_GS;
    // GPW consecutive pseudo-rows per warp and step (consecutive rows are the same allele: similar lengths)
    for (int64_t row0 = ((int64_t)blockIdx.x * nwarps + warp) * GPW; row0 < nrows; row0 += (int64_t)gridDim.x * nwarps * GPW) {
        const int64_t row = row0 + g;
        const bool act = row < nrows;
        const int64_t lb = act ? lptr[row] : 0, le = act ? lptr[row + 1] : 0;
        // UNR entries per lane per step; the codes of the NEXT step are fetched before this step's accumulator updates,
        // so the global-load latency overlaps the shared-memory read-modify-writes
        constexpr int UNR = 4;
        uint32_t cur[UNR], nxt[UNR];
        int64_t e = lb + li;
#pragma unroll
        for (int j = 0; j < UNR; ++j) cur[j] = (e + j * SPARSE_GS < le) ? __ldg(lcode + e + j * SPARSE_GS) : SPARSE_DENSE;
        while (e < le) {
            const int64_t en = e + UNR * SPARSE_GS;
#pragma unroll
            for (int j = 0; j < UNR; ++j) nxt[j] = (en + j * SPARSE_GS < le) ? __ldg(lcode + en + j * SPARSE_GS) : SPARSE_DENSE;
            uint32_t q[UNR];
#pragma unroll
            for (int j = 0; j < UNR; ++j) q[j] = (cur[j] != SPARSE_DENSE) ? lds_u32(tab + (cur[j] >> 1) * 4u) : 0u;
#pragma unroll
            for (int j = 0; j < UNR; ++j)
                if (cur[j] != SPARSE_DENSE) sparse_add<KP>(acc_base, K, q[j], (int64_t)(cur[j] >> 1), W, 1.0);
#pragma unroll
            for (int j = 0; j < UNR; ++j) cur[j] = nxt[j];
            e = en;
        }
        const int64_t hb = act ? hptr[row] : 0, he = act ? hptr[row + 1] : 0;
        for (int64_t h = hb + li; h < he; h += SPARSE_GS) {
            const int64_t c0 = (int64_t)(__ldg(hcode + h) >> 1);
            sparse_add<KP>(acc_base, K, lds_u32(tab + (uint32_t)c0 * 4u), c0, W, (double)__ldg(hcnt + h));
        }
        __syncwarp();
        // fixed tree over the lanes of the row group, state by state; slots are reset on the way
        double tot[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            double v = lds_f64(acc_base + k * 256u);
            sts_f64(acc_base + k * 256u, 0.0);
#pragma unroll
            for (int o = SPARSE_GS / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            tot[k] = v;
        }
        if (act) {
            double* y = S + row * KP;
#pragma unroll
            for (int j = 0; j < KP / 2; ++j)
                if (li == j) *reinterpret_cast<double2*>(y + 2 * j) = make_double2(tot[2 * j], tot[2 * j + 1]);
        }
    }
}

}  // namespace scs
