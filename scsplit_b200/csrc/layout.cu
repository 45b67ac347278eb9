// Device layout build: turns the reference's two SNV-major CSR count matrices (scSplit:546-547) into the
// two packed entry streams the EM kernels read (DESIGN.md "Data layout in HBM").
//
// Everything after the host->device copies runs on the GPU and is deterministic: the SNV-major stream keeps
// the input order, the cell-major stream is a STABLE radix sort of it by cell, so every row's entries are in
// ascending (SNV, allele) order on every run and the floating-point summation order of the EM never changes.
// cub (part of the CUDA toolkit) is used for the one-time scan/sort only, never on the per-iteration path.
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstdlib>

#include "common.cuh"
#include "estep_q.cuh"
#include "estep_lk.cuh"

namespace scs {

__device__ __forceinline__ int64_t load_count(const void* data, int64_t i, int bytes) {
    if (bytes == 2) return ((const int16_t*)data)[i];
    if (bytes == 4) return ((const int32_t*)data)[i];
    return ((const int64_t*)data)[i];
}

// one warp per pseudo-row pr = sel*V + v: number of count==1 / count>1 entries and the integer row sum
__global__ void k_count_rows(int64_t V, const int64_t* a_ptr, const void* a_val, const int64_t* r_ptr, const void* r_val,
                             int bytes, int64_t* cntL, int64_t* cntH, int64_t* rowsum, int* err) {
    int64_t pr = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (pr >= 2 * V) return;
    const int sel = pr >= V ? 1 : 0;
    const int64_t v = pr - (int64_t)sel * V;
    const int64_t* ptr = sel ? r_ptr : a_ptr;
    const void* val = sel ? r_val : a_val;
    int64_t b = ptr[v], e = ptr[v + 1];
    int64_t nl = 0, nh = 0, s = 0;
    for (int64_t i = b + lane; i < e; i += 32) {
        int64_t c = load_count(val, i, bytes);
        if (c <= 0 || c > 2147483647LL) atomicExch(err, 1);
        nl += (c == 1);
        nh += (c > 1);
        s += c;
    }
    for (int o = 16; o; o >>= 1) {
        nl += __shfl_xor_sync(0xffffffffu, nl, o);
        nh += __shfl_xor_sync(0xffffffffu, nh, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if (lane == 0) {
        cntL[pr] = nl;
        cntH[pr] = nh;
        rowsum[pr] = s;
    }
}

// one warp per pseudo-row: ordered (ballot) split of its entries into the light and heavy SNV-major streams,
// plus the (cell key, cell-major code) pairs that the stable sort turns into the cell-major stream.
__global__ void k_fill_M(int64_t V, int64_t N, const int64_t* a_ptr, const int32_t* a_idx, const void* a_val,
                         const int64_t* r_ptr, const int32_t* r_idx, const void* r_val, int bytes,
                         const int64_t* lptr, const int64_t* hptr, uint32_t* lcode, uint32_t* hcode, int32_t* hcnt,
                         uint32_t* lkey, uint32_t* lval, uint32_t* hkey, uint64_t* hval,
                         unsigned long long* ndup, int* err) {
    int64_t pr = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (pr >= 2 * V) return;
    const int sel = pr >= V ? 1 : 0;
    const int64_t v = pr - (int64_t)sel * V;
    const int64_t* ptr = sel ? r_ptr : a_ptr;
    const int32_t* idx = sel ? r_idx : a_idx;
    const void* val = sel ? r_val : a_val;
    int64_t b = ptr[v], e = ptr[v + 1];
    int64_t ab = a_ptr[v], ae = a_ptr[v + 1];
    int64_t lpos = lptr[pr], hpos = hptr[pr];
    unsigned long long dups = 0;
    for (int64_t i0 = b; i0 < e; i0 += 32) {
        int64_t i = i0 + lane;
        bool valid = i < e;
        int64_t c = valid ? load_count(val, i, bytes) : 0;
        int32_t cell = valid ? idx[i] : 0;
        if (valid && (cell < 0 || cell >= N)) atomicExch(err, 2);
        if (valid && i > b && idx[i - 1] >= cell) atomicExch(err, 3);  // indices must be sorted and unique
        uint32_t dup = 0;
        if (valid && sel) {  // does the same (SNV, cell) have an alt entry?  binary search in the alt row
            int64_t lo = ab, hi = ae;
            while (lo < hi) {
                int64_t mid = (lo + hi) >> 1;
                if (a_idx[mid] < cell) lo = mid + 1; else hi = mid;
            }
            dup = (lo < ae && a_idx[lo] == cell) ? 1u : 0u;
            dups += dup;
        }
        bool light = valid && c == 1, heavy = valid && c > 1;
        unsigned ml = __ballot_sync(0xffffffffu, light), mh = __ballot_sync(0xffffffffu, heavy);
        unsigned lt = (1u << lane) - 1u;
        uint32_t codeM = ((uint32_t)cell << 1) | dup;
        uint32_t codeE = ((uint32_t)pr << 1) | dup;
        if (light) {
            int64_t p = lpos + __popc(ml & lt);
            lcode[p] = codeM;
            lkey[p] = (uint32_t)cell;
            lval[p] = codeE;
        }
        if (heavy) {
            int64_t p = hpos + __popc(mh & lt);
            hcode[p] = codeM;
            hcnt[p] = (int32_t)c;
            hkey[p] = (uint32_t)cell;
            hval[p] = (uint64_t)codeE | ((uint64_t)(uint32_t)c << 32);
        }
        lpos += __popc(ml);
        hpos += __popc(mh);
    }
    for (int o = 16; o; o >>= 1) dups += __shfl_xor_sync(0xffffffffu, dups, o);
    if (lane == 0 && dups) atomicAdd(ndup, dups);
}

// row pointers from sorted keys: ptr[r] = first position whose key >= r
__global__ void k_ptr_from_sorted(const uint32_t* keys, int64_t n, int64_t nrows, int64_t* ptr) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int64_t kprev = (i == 0) ? -1 : (int64_t)keys[i - 1];
    int64_t k = (i == n) ? nrows : (int64_t)keys[i];
    for (int64_t r = kprev + 1; r <= k; ++r) ptr[r] = i;
}

__global__ void k_split_heavy(const uint64_t* hval, int64_t n, uint32_t* hcode, int32_t* hcnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x = hval[i];
    hcode[i] = (uint32_t)x;
    hcnt[i] = (int32_t)(x >> 32);
}

__global__ void k_row_sums(int64_t V, const int64_t* rowsum, int64_t* sum_alt, int64_t* sum_ref) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    sum_alt[v] = rowsum[v];
    sum_ref[v] = rowsum[V + v];
}

// k_alt, scSplit:100-103: N_alt / (N_ref + N_alt) with N_x = row sum + 1; integer-exact operands
__global__ void k_kalt(int64_t V, const int64_t* sum_alt, const int64_t* sum_ref, double* kalt) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    int64_t na = sum_alt[v] + 1, nr = sum_ref[v] + 1;
    kalt[v] = (double)na / (double)(nr + na);
}

void free_stream(scs_ctx* ctx, Stream& s) {
    dev_free(ctx, s.lptr);
    dev_free(ctx, s.hptr);
    dev_free(ctx, s.lcode);
    dev_free(ctx, s.hcode);
    dev_free(ctx, s.hcnt);
}

static int bits_for(int64_t n) {
    int b = 1;
    while ((1LL << b) < n) ++b;
    return b;
}

template <typename VT>
static int stable_sort_by_cell(scs_ctx* ctx, const uint32_t* kin, uint32_t* kout, const VT* vin, VT* vout, int64_t n,
                               int64_t nrows) {
    if (n == 0) return 0;
    size_t tb = 0;
    int eb = bits_for(nrows);
    SCS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, kin, kout, vin, vout, n, 0, eb, ctx->stream));
    void* tmp = nullptr;
    SCS_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, ctx->stream));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tb, kin, kout, vin, vout, n, 0, eb, ctx->stream);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFreeAsync(tmp, ctx->stream);
    SCS_CUDA(e);
    SCS_CUDA(e2);
    ctx->launches += 4;
    return 0;
}

// ---- job order of the sparse M-step: pseudo-rows by descending length (stable, so ties keep their index order)
static __global__ void k_row_len_keys(int64_t nrows, int nchunks, const int64_t* __restrict__ lptr, uint32_t* __restrict__ key,
                                      int32_t* __restrict__ val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    if (nchunks <= 1) {
        key[i] = 0xffffffffu - (uint32_t)min((int64_t)0xffffffffLL, lptr[i + 1] - lptr[i]);
    } else {          // chunk in the top 4 bits, inverted length below: rows of a chunk stay together, longest first
        const uint32_t chunk = (uint32_t)(i * nchunks / nrows);
        key[i] = (chunk << 28) | (0x0fffffffu - (uint32_t)min((int64_t)0x0fffffffLL, lptr[i + 1] - lptr[i]));
    }
    val[i] = (int32_t)i;
}

int build_row_order(scs_ctx* ctx, const Stream& s, int32_t** order, int nchunks) {
    const int64_t n = s.nrows;
    uint32_t *kin = nullptr, *kout = nullptr;
    int32_t *vin = nullptr, *vout = nullptr;
    void* tmp = nullptr;
    auto body = [&]() -> int {
        SCS_TRY(dev_alloc(ctx, &kin, n));
        SCS_TRY(dev_alloc(ctx, &kout, n));
        SCS_TRY(dev_alloc(ctx, &vin, n));
        SCS_TRY(dev_alloc(ctx, &vout, n));
        k_row_len_keys<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, nchunks, s.lptr, kin, vin);
        size_t tb = 0;
        SCS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, kin, kout, vin, vout, n, 0, 32, ctx->stream));
        SCS_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, ctx->stream));
        SCS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, kin, kout, vin, vout, n, 0, 32, ctx->stream));
        ctx->launches += 5;
        return 0;
    };
    const int rc = body();
    if (tmp) cudaFreeAsync(tmp, ctx->stream);
    dev_free(ctx, kin); dev_free(ctx, kout); dev_free(ctx, vin);
    if (rc) { dev_free(ctx, vout); return rc; }
    *order = vout;
    return 0;
}

static int inclusive_scan_into_ptr(scs_ctx* ctx, const int64_t* cnt, int64_t* ptr, int64_t n) {
    SCS_CUDA(cudaMemsetAsync(ptr, 0, sizeof(int64_t), ctx->stream));
    size_t tb = 0;
    SCS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, cnt, ptr + 1, n, ctx->stream));
    void* tmp = nullptr;
    SCS_CUDA(cudaMallocAsync(&tmp, tb ? tb : 1, ctx->stream));
    cudaError_t e = cub::DeviceScan::InclusiveSum(tmp, tb, cnt, ptr + 1, n, ctx->stream);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFreeAsync(tmp, ctx->stream);
    SCS_CUDA(e);
    SCS_CUDA(e2);
    ctx->launches += 2;
    return 0;
}

// ---- tile-major copy of the count-1 entries (v2 SpMM) -----------------------------------------------------

__device__ __forceinline__ int64_t row_lower_bound(const uint32_t* __restrict__ code, int64_t b, int64_t e, int64_t key) {
    while (b < e) {   // first position whose gather row (code >> 1) is >= key
        int64_t mid = (b + e) >> 1;
        if ((int64_t)(code[mid] >> 1) < key) b = mid + 1; else e = mid;
    }
    return b;
}

// bp[b][row] = first entry of `row` whose index (code >> 1) is >= b * block (b = 0 .. nblocks); bp[0] = lptr[row], bp[nblocks] = end
static __global__ void k_block_ptrs(int64_t nrows, int nblocks, int64_t block, const int64_t* __restrict__ lptr, const uint32_t* __restrict__ lcode,
                                    int64_t* __restrict__ bp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * (nblocks + 1)) return;
    const int64_t b = i / nrows, row = i - b * nrows;
    const int64_t lo = lptr[row], hi = lptr[row + 1];
    bp[i] = b == 0 ? lo : b == nblocks ? hi : row_lower_bound(lcode, lo, hi, b * block);
}
int build_block_ptrs(scs_ctx* ctx, const Stream& s, int nblocks, int64_t block, int64_t** bp) {
    *bp = nullptr;
    SCS_TRY(dev_alloc(ctx, bp, s.nrows * (nblocks + 1)));
    const int64_t n = s.nrows * (nblocks + 1);
    k_block_ptrs<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(s.nrows, nblocks, block, s.lptr, s.lcode, *bp);
    ctx->launches++;
    SCS_CUDA(cudaGetLastError());
    return 0;
}

// one thread per unit u = tile * nrows + row: entries of `row` inside `tile`, padded to a multiple of `pad`
__global__ void k_tile_counts(int64_t nrows, int ntiles, int TR, int pad, const int64_t* __restrict__ lptr,
                              const uint32_t* __restrict__ lcode, int64_t* __restrict__ cnt, int32_t* __restrict__ ulo) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nrows * ntiles) return;
    const int64_t t = u / nrows, row = u - t * nrows;
    const int64_t b = lptr[row], e = lptr[row + 1];
    const int64_t lo = row_lower_bound(lcode, b, e, t * TR), hi = row_lower_bound(lcode, b, e, (t + 1) * (int64_t)TR);
    cnt[u] = ((hi - lo) + pad - 1) / pad * pad;
    ulo[u] = (int32_t)(lo - b);     // where the unit's entries start inside the row (a row holds < 2^31 entries)
}

// one warp per row: scatter its entries to their units as 16-bit in-tile offsets; pad slots point at the zero row
__global__ void k_tile_scatter(int64_t nrows, int ntiles, int TR, const int64_t* __restrict__ lptr, const uint32_t* __restrict__ lcode,
                               const int64_t* __restrict__ tptr, const int32_t* __restrict__ ulo, uint16_t* __restrict__ tcode) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int64_t b = lptr[row], e = lptr[row + 1];
    for (int64_t i = b + lane; i < e; i += 32) {
        const int64_t g = (int64_t)(lcode[i] >> 1);
        const int64_t t = g / TR;
        const int64_t u = t * nrows + row;
        tcode[tptr[u] + (i - b - ulo[u])] = (uint16_t)(g - t * TR);
    }
    for (int64_t t = lane; t < ntiles; t += 32) {   // pad slots: from the unit's true count up to its padded length
        const int64_t u = t * nrows + row, base = tptr[u], plen = tptr[u + 1] - base;
        const int64_t cnt = (t + 1 < ntiles ? (int64_t)ulo[u + nrows] : e - b) - ulo[u];
        for (int64_t k = cnt; k < plen; ++k) tcode[base + k] = (uint16_t)TR;
    }
}

// cost of the units [lo, u): entries plus a fixed charge per unit (flush + bounds), in entry units
constexpr int64_t UNIT_COST = 8;
// `tile_cost` (entries) is charged for every tile boundary inside the range -- a CTA whose range runs into the next tile loads a
// second table tile (a few microseconds with every CTA pulling its tile from L2 at once) and should get fewer entries for it;
// `nrows` units make a tile.  tile_cost = 0 inside a segment (one tile by construction).
__device__ __forceinline__ int64_t plan_cost(const int64_t* __restrict__ tptr, int64_t lo, int64_t u, int64_t nrows = 1, int64_t tile_cost = 0,
                                             int64_t unit_cost = UNIT_COST) {
    return (tptr[u] - tptr[lo]) + unit_cost * (u - lo) + (tile_cost ? tile_cost * ((u > lo ? (u - 1) / nrows : lo / nrows) - lo / nrows) : 0);
}
__device__ __forceinline__ int64_t plan_cut(const int64_t* __restrict__ tptr, int64_t lo, int64_t hi, int64_t part, int64_t nparts,
                                            int64_t nrows = 1, int64_t tile_cost = 0, int64_t unit_cost = UNIT_COST) {
    if (part <= 0) return lo;
    if (part >= nparts) return hi;
    const int64_t total = plan_cost(tptr, lo, hi, nrows, tile_cost, unit_cost);
    const int64_t target = (total / nparts) * part + ((total % nparts) * part) / nparts;
    int64_t a = lo, b = hi;
    while (a < b) {          // first unit u with cost(lo, u) >= target
        const int64_t mid = (a + b) >> 1;
        if (plan_cost(tptr, lo, mid, nrows, tile_cost, unit_cost) < target) a = mid + 1; else b = mid;
    }
    return a;
}
// cu[j] = first unit of CTA j (equal-cost split of all units)
__global__ void k_plan_ctas(const int64_t* __restrict__ tptr, int64_t nunits, int ncta, int64_t* __restrict__ cu, int64_t nrows, int64_t tile_cost,
                            int64_t unit_cost = UNIT_COST) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > ncta) return;
    cu[j] = plan_cut(tptr, 0, nunits, j, ncta, nrows, tile_cost, unit_cost);
}
// Tile segments of the CTA ranges: CTA j runs the pieces of [cu[j], cu[j+1]) between tile boundaries (a tile owns nrows
// consecutive units).  At most ncta + ntiles pieces; the unused tail of the arrays is left as empty segments.  One block:
// thread j counts the pieces of CTA j, a shared-memory scan places them, thread j writes them.
__global__ void __launch_bounds__(1024) k_plan_segments(const int64_t* __restrict__ cu, int ncta, int64_t nrows, int nseg_max,
                                                        int32_t* __restrict__ seg_ptr, int32_t* __restrict__ seg_tile,
                                                        int64_t* __restrict__ seg_lo, int64_t* __restrict__ seg_hi) {
    __shared__ int s_off[1025];
    const int j = threadIdx.x;
    int64_t lo_u = 0, hi_u = 0;
    int cnt = 0;
    if (j < ncta) {
        lo_u = cu[j];
        hi_u = cu[j + 1];
        if (hi_u > lo_u) cnt = (int)((hi_u - 1) / nrows - lo_u / nrows) + 1;   // every tile touched holds at least one unit
    }
    s_off[j + 1] = cnt;
    if (j == 0) s_off[0] = 0;
    __syncthreads();
    if (j == 0)
        for (int i = 1; i <= ncta; ++i) s_off[i] = min(s_off[i] + s_off[i - 1], nseg_max);
    __syncthreads();
    if (j < ncta) {
        int n = s_off[j];
        seg_ptr[j] = n;
        for (int64_t tt = lo_u / nrows; tt * nrows < hi_u && n < s_off[j + 1]; ++tt, ++n) {
            seg_tile[n] = (int32_t)tt;
            seg_lo[n] = max(lo_u, tt * nrows);
            seg_hi[n] = min(hi_u, (tt + 1) * nrows);
        }
    }
    if (j == 0) seg_ptr[ncta] = s_off[ncta];
    for (int n = s_off[ncta] + j; n < nseg_max; n += blockDim.x) { seg_tile[n] = 0; seg_lo[n] = 0; seg_hi[n] = 0; }
}
// gb[seg][G] = first unit of row group G inside segment [lo, hi)
__global__ void k_plan_groups(const int64_t* __restrict__ tptr, const int64_t* __restrict__ seg_lo, const int64_t* __restrict__ seg_hi,
                              int ngroups, int64_t* __restrict__ gb, int64_t unit_cost = UNIT_COST) {
    const int seg = blockIdx.x;
    for (int G = threadIdx.x; G <= ngroups; G += blockDim.x)
        gb[(int64_t)seg * (ngroups + 1) + G] = plan_cut(tptr, seg_lo[seg], seg_hi[seg], G, ngroups, 1, 0, unit_cost);
}

// ---- build-time order of the entries inside the units of a packed tile-major stream (spmm_tiled.cuh, "packed" groups).
// One warp per (segment, row group): it walks the group's run of units, and inside every unit it puts, wherever the unit
// has one, an entry whose in-tile row r satisfies (r & 7) == (g + p) & 7 at stream position p of the run (g = the group's
// index inside its warp); the entries that find no slot of their class fill the slots that found no entry.  Only the order
// inside a unit changes, so sums are reassociated, never altered; the result is a pure function of the input.
// phase[u] = class wanted at the first slot of unit u: (g + position of the unit inside its group's run) & 7
__global__ void k_unit_phase(const int64_t* __restrict__ tptr, const int64_t* __restrict__ gb, int ngroups, int gpw, int nseg,
                             uint8_t* __restrict__ phase) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = (int)(w / ngroups), G = (int)(w % ngroups);
    if (seg >= nseg) return;
    const int64_t u0 = gb[(int64_t)seg * (ngroups + 1) + G], u1 = gb[(int64_t)seg * (ngroups + 1) + G + 1];
    if (u0 >= u1) return;
    const int g = G % gpw;
    const int64_t run0 = tptr[u0];
    for (int64_t u = u0; u < u1; ++u) phase[u] = (uint8_t)((g + (tptr[u] - run0)) & 7);
}
// A warp takes 32 consecutive units (contiguous in the stream), stages them in shared memory with coalesced loads, and each
// lane orders ONE unit serially -- the per-class counters live in the eight bytes of a 64-bit register -- before the warp
// writes the range back.  (A warp per unit spends its time in shuffles and waits: 0.25 ms at config 3 against 0.05 ms.)
constexpr int ORDER_WARPS = 4, ORDER_CAP = 2816;     // entries a warp can stage (32 units of ~66 entries are ~2100)
__global__ void __launch_bounds__(ORDER_WARPS * 32) k_order_units(const int64_t* __restrict__ tptr, uint16_t* __restrict__ tcode,
                                                                 const uint8_t* __restrict__ phase, int64_t nunits, int upw) {
    __shared__ __align__(16) uint16_t s_in[ORDER_WARPS][ORDER_CAP], s_out[ORDER_WARPS][ORDER_CAP];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int64_t U0 = ((int64_t)blockIdx.x * ORDER_WARPS + wl) * upw;   // upw (<= 32) units per warp: what usually fits the staging
    if (U0 >= nunits) return;
    const int64_t Uend = min(U0 + upw, nunits);
    const int64_t Ul = min(U0 + lane, Uend - 1);                      // lanes past the last unit shadow it and do nothing
    const int64_t e0 = tptr[U0], e1 = tptr[Uend];
    const int len = (int)min(e1 - e0, (int64_t)ORDER_CAP + 1);
    if (len < 2 || len > ORDER_CAP) return;                           // (over-long ranges keep their natural order)
    const int ua = (int)(tptr[Ul] - e0), n = (U0 + lane < Uend) ? (int)(tptr[Ul + 1] - tptr[Ul]) : 0;
    const int ph = phase[Ul];                                         // slot p of this unit wants class (ph + p) & 7
    uint16_t* in = s_in[wl];
    uint16_t* out = s_out[wl];
    // units are padded to an even number of entries and start at even positions: move 32-bit words
    for (int i = lane; i < len / 2; i += 32)
        reinterpret_cast<uint32_t*>(in)[i] = reinterpret_cast<const uint32_t*>(tcode + e0)[i];
    __syncwarp();
    if (n >= 2 && n <= 255) {
        uint64_t cntE = 0;                                            // byte r: entries of class r
#pragma unroll 8
        for (int i = 0; i < n; ++i) cntE += 1ull << (8 * (in[ua + i] & 7));
        uint64_t cntS = 0, preLE = 0, preLS = 0;                      // slots by class; exclusive prefixes of the surpluses
        int le = 0, ls = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int first = (r - ph) & 7;
            const int sl = first < n ? (n - first + 7) >> 3 : 0, en = (int)((cntE >> (8 * r)) & 255);
            cntS |= (uint64_t)sl << (8 * r);
            preLE |= (uint64_t)le << (8 * r);
            preLS |= (uint64_t)ls << (8 * r);
            le += max(en - sl, 0);
            ls += max(sl - en, 0);
        }
        uint64_t run = 0;                                             // byte r: entries of class r placed so far
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const uint16_t v = in[ua + i];
            const int cls = v & 7, sh = 8 * cls;
            const int rank = (int)((run >> sh) & 255), sl = (int)((cntS >> sh) & 255);
            run += 1ull << sh;
            int p;
            if (rank < sl) {
                p = ((cls - ph) & 7) + 8 * rank;
            } else {                                                  // surplus entry: the idx-th surplus slot, classes ascending
                const int idx = (int)((preLE >> sh) & 255) + rank - sl;
                int r2 = 0;
#pragma unroll
                for (int r = 1; r < 8; ++r) r2 = (int)((preLS >> (8 * r)) & 255) <= idx ? r : r2;
                // (the last class whose prefix is <= idx holds the idx-th surplus slot: the next prefix is larger)
                p = ((r2 - ph) & 7) + 8 * ((int)((cntE >> (8 * r2)) & 255) + idx - (int)((preLS >> (8 * r2)) & 255));
            }
            out[ua + p] = v;
        }
    } else {
        for (int i = 0; i < n; ++i) out[ua + i] = in[ua + i];
    }
    __syncwarp();
    for (int i = lane; i < len / 2; i += 32)
        reinterpret_cast<uint32_t*>(tcode + e0)[i] = reinterpret_cast<const uint32_t*>(out)[i];
}

void free_tiled(scs_ctx* ctx, TiledStream& t) {
    dev_free(ctx, t.tptr);
    dev_free(ctx, t.tcode);
    dev_free(ctx, t.seg_ptr);
    dev_free(ctx, t.seg_tile);
    dev_free(ctx, t.gb);
    t = TiledStream();
}

int build_tiled(scs_ctx* ctx, const Stream& s, TiledStream& t, int KP, int TR, int pad, int ncta, int ngroups, int order_gpw) {
    free_tiled(ctx, t);
    SCS_CHECK(ncta >= 1 && ncta <= 1024, "work plan supports 1..1024 CTAs (got %d)", ncta);
    t.KP = KP;
    t.TR = TR;
    t.ntiles = (int)((s.xrows + TR - 1) / TR);
    t.nunits = (int64_t)t.ntiles * s.nrows;
    cudaStream_t st = ctx->stream;
    SCS_TRY(dev_alloc(ctx, &t.tptr, t.nunits + 1));
    SCS_CUDA(cudaMemsetAsync(t.tptr, 0, sizeof(int64_t), st));
    Scratch tmps(st);
    int32_t* d_ulo = nullptr;
    SCS_TRY(tmps.get(&d_ulo, t.nunits));
    k_tile_counts<<<(unsigned)((t.nunits + 255) / 256), 256, 0, st>>>(s.nrows, t.ntiles, TR, pad, s.lptr, s.lcode, t.tptr + 1, d_ulo);
    ctx->launches++;
    size_t tb = 0;
    SCS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, t.tptr + 1, t.tptr + 1, t.nunits, st));
    void* tmp = nullptr;
    SCS_TRY(tmps.get_bytes(&tmp, tb));
    SCS_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb, t.tptr + 1, t.tptr + 1, t.nunits, st));
    ctx->launches += 2;
    // Nothing below waits for the device: the stream buffer is sized by its upper bound (every unit is padded by fewer than
    // `pad` entries) and the work plan is computed where the unit pointers are.
    const int64_t total_max = s.n_light + t.nunits * (int64_t)(pad > 0 ? pad - 1 : 0);
    SCS_TRY(dev_alloc(ctx, &t.tcode, total_max + 64));   // slack: the kernel prefetches one batch past a group's run
    k_tile_scatter<<<(unsigned)((s.nrows + 7) / 8), 256, 0, st>>>(s.nrows, t.ntiles, TR, s.lptr, s.lcode, t.tptr, d_ulo, t.tcode);
    ctx->launches++;
    // ---- static work plan: CTA ranges, their tile segments (<= ncta + ntiles of them), group cuts
    t.ncta = ncta;
    t.ngroups = ngroups;
    t.nseg = ncta + t.ntiles;                            // capacity; seg_ptr[ncta] holds the number in use
    int64_t *d_cu = nullptr, *d_lo = nullptr, *d_hi = nullptr;
    SCS_TRY(tmps.get(&d_cu, ncta + 1));
    SCS_TRY(tmps.get(&d_lo, t.nseg));
    SCS_TRY(tmps.get(&d_hi, t.nseg));
    SCS_TRY(dev_alloc(ctx, &t.seg_ptr, ncta + 1));
    SCS_TRY(dev_alloc(ctx, &t.seg_tile, t.nseg));
    SCS_TRY(dev_alloc(ctx, &t.gb, (int64_t)t.nseg * (ngroups + 1)));
    const char* tc_env = std::getenv("SCS_TILE_COST");   // (see build_qtiled; the same charge per extra tile, in entries)
    const int64_t tile_cost = tc_env ? atoll(tc_env) : (int64_t)TR * KP * 8 / 8;
    k_plan_ctas<<<(ncta + 1 + 127) / 128, 128, 0, st>>>(t.tptr, t.nunits, ncta, d_cu, s.nrows, tile_cost);
    k_plan_segments<<<1, 1024, 0, st>>>(d_cu, ncta, s.nrows, t.nseg, t.seg_ptr, t.seg_tile, d_lo, d_hi);
    k_plan_groups<<<t.nseg, 256, 0, st>>>(t.tptr, d_lo, d_hi, ngroups, t.gb);
    ctx->launches += 3;
    if (order_gpw > 0) {                                 // packed row groups: order the entries inside the units (see k_order_units)
        uint8_t* d_phase = nullptr;
        SCS_TRY(tmps.get(&d_phase, t.nunits));
        SCS_CUDA(cudaMemsetAsync(d_phase, 0, (size_t)t.nunits, st));
        const int64_t runs = (int64_t)t.nseg * ngroups;
        k_unit_phase<<<(unsigned)((runs + 255) / 256), 256, 0, st>>>(t.tptr, t.gb, ngroups, order_gpw, t.nseg, d_phase);
        // units per warp: two thirds of the staging buffer at the average (padded) unit length, so that ordinary ranges fit
        const int64_t avg = std::max<int64_t>(2, (s.n_light + t.nunits) / std::max<int64_t>(t.nunits, 1));
        const int upw = (int)std::min<int64_t>(32, std::max<int64_t>(1, (int64_t)ORDER_CAP * 2 / 3 / avg));
        const int64_t owarps = (t.nunits + upw - 1) / upw;
        k_order_units<<<(unsigned)((owarps + ORDER_WARPS - 1) / ORDER_WARPS), ORDER_WARPS * 32, 0, st>>>(t.tptr, t.tcode, d_phase, t.nunits, upw);
        ctx->launches += 2;
    }
    SCS_CUDA(cudaGetLastError());
    return 0;
}

// ---- tile-major stream of the fixed-point E-step (estep_q.cuh) --------------------------------------------------
// unit u = tile * nrows + row holds one 16-bit code (the in-tile row) per read of the row's entries inside the tile --
// an entry of count c <= Q_REP_MAX is written c times, larger counts are left to the combine kernel -- padded to a
// multiple of Q_UNIT_PAD codes.  Integer sums: the order inside a unit is free.  The codes are packed into the kernel's 32-bit words
// of two (QCode) as the last step.

__device__ __forceinline__ int64_t q_codes_of(int64_t c) { return c <= Q_REP_MAX ? c : 0; }

// one thread per unit: padded code count, where the unit's light / heavy entries start inside the row
__global__ void k_qunit_counts(int64_t nrows, int ntiles, int TR, const int64_t* __restrict__ lptr, const uint32_t* __restrict__ lcode,
                               const int64_t* __restrict__ hptr, const uint32_t* __restrict__ hcode, const int32_t* __restrict__ hcnt,
                               int64_t* __restrict__ cnt, int32_t* __restrict__ ulo, int32_t* __restrict__ uho, int pad = Q_UNIT_PAD) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nrows * ntiles) return;
    const int64_t t = u / nrows, row = u - t * nrows;
    const int64_t b = lptr[row], e = lptr[row + 1], hb = hptr[row], he = hptr[row + 1];
    const int64_t lo = row_lower_bound(lcode, b, e, t * TR), hi = row_lower_bound(lcode, b, e, (t + 1) * (int64_t)TR);
    const int64_t hlo = row_lower_bound(hcode, hb, he, t * TR), hhi = row_lower_bound(hcode, hb, he, (t + 1) * (int64_t)TR);
    int64_t n = hi - lo;
    for (int64_t h = hlo; h < hhi; ++h) n += q_codes_of(hcnt[h]);
    cnt[u] = (n + pad - 1) / pad * pad;
    ulo[u] = (int32_t)(lo - b);
    uho[u] = (int32_t)(hlo - hb);
}
// one thread per row: total reads of the row (a cell's depth bounds its whole sum)
__global__ void k_qrow_depth(int64_t nrows, const int64_t* __restrict__ lptr, const int64_t* __restrict__ hptr, const int32_t* __restrict__ hcnt,
                             unsigned long long* __restrict__ maxdepth) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    int64_t d = lptr[row + 1] - lptr[row];
    for (int64_t h = hptr[row]; h < hptr[row + 1]; ++h) d += hcnt[h];
    if ((unsigned long long)d > *maxdepth) atomicMax(maxdepth, (unsigned long long)d);
}
// one warp per row: light entries by the warp, then lane = tile for the repeated codes of the heavy entries and the padding
__global__ void k_qtile_scatter(int64_t nrows, int ntiles, int TR, const int64_t* __restrict__ lptr, const uint32_t* __restrict__ lcode,
                                const int64_t* __restrict__ hptr, const uint32_t* __restrict__ hcode, const int32_t* __restrict__ hcnt,
                                const int64_t* __restrict__ tptr, const int32_t* __restrict__ ulo, const int32_t* __restrict__ uho,
                                uint16_t* __restrict__ tcode) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int64_t b = lptr[row], e = lptr[row + 1], hb = hptr[row], he = hptr[row + 1];
    for (int64_t i = b + lane; i < e; i += 32) {
        const int64_t g = (int64_t)(lcode[i] >> 1);
        const int64_t t = g / TR;
        const int64_t u = t * nrows + row;
        SCS_ASSERT(t < ntiles && i - b - ulo[u] >= 0 && tptr[u] + (i - b - ulo[u]) < tptr[u + 1]);
        tcode[tptr[u] + (i - b - ulo[u])] = (uint16_t)(g - t * TR);
    }
    for (int64_t t = lane; t < ntiles; t += 32) {
        const int64_t u = t * nrows + row, end = tptr[u + 1];
        const bool last = t + 1 >= ntiles;
        int64_t pos = tptr[u] + ((last ? e - b : (int64_t)ulo[u + nrows]) - ulo[u]);
        const int64_t h1 = hb + (last ? he - hb : (int64_t)uho[u + nrows]);
        for (int64_t h = hb + uho[u]; h < h1; ++h) {
            const uint16_t rowt = (uint16_t)((int64_t)(hcode[h] >> 1) - t * TR);
            for (int64_t c = q_codes_of(hcnt[h]); c > 0; --c) tcode[pos++] = rowt;
        }
        SCS_ASSERT(pos <= end);
        for (; pos < end; ++pos) tcode[pos] = (uint16_t)TR;
    }
}
// Order inside the units for 64-byte rows: two row groups share a quarter-warp wavefront, conflict-free when their rows lie
// in different halves of the 128 bytes, i.e. differ in parity.  Position p of a unit whose phase is ph gets, wherever the
// unit has one, a row of parity (ph + p) & 1 -- neighbouring groups have opposite phases.  The unit's padding codes are
// wildcards: there is a zero row of either parity (TR, TR + 1), so they fill slots whose parity has run out of entries
// before any entry has to take a slot of the wrong parity.  One lane orders one unit, staged in shared memory like
// k_order_units.
__global__ void __launch_bounds__(ORDER_WARPS * 32) k_qorder_units(const int64_t* __restrict__ tptr, uint16_t* __restrict__ tcode,
                                                                  const uint8_t* __restrict__ phase, int64_t nunits, int upw, int TR) {
    __shared__ __align__(16) uint16_t s_in[ORDER_WARPS][ORDER_CAP], s_out[ORDER_WARPS][ORDER_CAP];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int64_t U0 = ((int64_t)blockIdx.x * ORDER_WARPS + wl) * upw;
    if (U0 >= nunits) return;
    const int64_t Uend = min(U0 + upw, nunits);
    const int64_t Ul = min(U0 + lane, Uend - 1);
    const int64_t e0 = tptr[U0], e1 = tptr[Uend];
    const int len = (int)min(e1 - e0, (int64_t)ORDER_CAP + 1);
    if (len < 2 || len > ORDER_CAP) return;                           // (over-long ranges keep their natural order)
    const int ua = (int)(tptr[Ul] - e0), n = (U0 + lane < Uend) ? (int)(tptr[Ul + 1] - tptr[Ul]) : 0;
    const int ph = phase[Ul] & 1;
    uint16_t* in = s_in[wl];
    uint16_t* out = s_out[wl];
    for (int i = lane; i < len / 2; i += 32)
        reinterpret_cast<uint32_t*>(in)[i] = reinterpret_cast<const uint32_t*>(tcode + e0)[i];
    __syncwarp();
    int npad = 0;                                                     // padding codes sit at the end of the unit
    while (npad < n && in[ua + n - 1 - npad] == (uint16_t)TR) ++npad;
    const int ne = n - npad;                                          // real entries: in[ua .. ua + ne)
    const uint16_t zrow[2] = {(uint16_t)(TR + (TR & 1)), (uint16_t)(TR + ((TR & 1) ^ 1))};   // zero row of parity 0 / 1
    int i0 = 0, i1 = 0;                                               // next unread entry of parity 0 / 1
    for (int p = 0; p < n; ++p) {
        while (i0 < ne && (in[ua + i0] & 1u) != 0u) ++i0;
        while (i1 < ne && (in[ua + i1] & 1u) != 1u) ++i1;
        const int want = (ph + p) & 1;
        int& iw = want ? i1 : i0;
        int& io = want ? i0 : i1;
        uint16_t c;
        if (iw < ne) c = in[ua + iw++];
        else if (npad > 0) { c = zrow[want]; --npad; }
        else c = in[ua + io++];
        out[ua + p] = c;
    }
    __syncwarp();
    for (int i = lane; i < len / 2; i += 32)
        reinterpret_cast<uint32_t*>(tcode + e0)[i] = reinterpret_cast<const uint32_t*>(out)[i];
}
// codes (pairs of 16-bit rows) -> stream words, in place
template <int LANES>
__global__ void k_qpack_words(int64_t nwords, uint32_t* __restrict__ w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    const uint32_t x = w[i];
    w[i] = QCode<LANES>::word(x & 0xffffu, x >> 16);
}

int build_qtiled(scs_ctx* ctx, const Stream& s, TiledStream& t, int lanes, int TR, int ncta, int ngroups, int* F, int* ok) {
    free_tiled(ctx, t);
    *ok = 0;
    SCS_CHECK(ncta >= 1 && ncta <= 1024, "work plan supports 1..1024 CTAs (got %d)", ncta);
    cudaStream_t st = ctx->stream;
    const int ntiles = (int)((s.xrows + TR - 1) / TR);
    const int64_t nunits = (int64_t)ntiles * s.nrows;
    Scratch tmps(st);
    int32_t *d_ulo = nullptr, *d_uho = nullptr;
    unsigned long long* d_max = nullptr;
    int64_t* tptr = nullptr;
    SCS_TRY(tmps.get(&d_ulo, nunits));
    SCS_TRY(tmps.get(&d_uho, nunits));
    SCS_TRY(tmps.get(&d_max, 1));
    SCS_CUDA(cudaMemsetAsync(d_max, 0, 8, st));
    SCS_TRY(dev_alloc(ctx, &tptr, nunits + 1));
    t.tptr = tptr;                                       // (free_tiled releases it on every path from here)
    SCS_CUDA(cudaMemsetAsync(tptr, 0, sizeof(int64_t), st));
    k_qunit_counts<<<(unsigned)((nunits + 255) / 256), 256, 0, st>>>(s.nrows, ntiles, TR, s.lptr, s.lcode, s.hptr, s.hcode, s.hcnt, tptr + 1,
                                                                      d_ulo, d_uho);
    k_qrow_depth<<<(unsigned)((s.nrows + 255) / 256), 256, 0, st>>>(s.nrows, s.lptr, s.hptr, s.hcnt, d_max);
    ctx->launches += 2;
    unsigned long long h_max = 0;
    SCS_CUDA(cudaMemcpyAsync(&h_max, d_max, 8, cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    *F = h_max < (1ull << 40) ? q_frac_bits((int64_t)h_max) : 0;
    if (*F < Q_MIN_F) {                                  // cells too deep for a useful fraction: fp64 keeps the E-step
        free_tiled(ctx, t);
        return 0;
    }
    size_t tb = 0;
    SCS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, tptr + 1, tptr + 1, nunits, st));
    void* tmp = nullptr;
    SCS_TRY(tmps.get_bytes(&tmp, tb));
    SCS_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb, tptr + 1, tptr + 1, nunits, st));
    ctx->launches += 2;
    int64_t total = 0;
    SCS_CUDA(cudaMemcpyAsync(&total, tptr + nunits, 8, cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    t.KP = lanes;
    t.TR = TR;
    t.ntiles = ntiles;
    t.nunits = nunits;
    SCS_TRY(dev_alloc(ctx, &t.tcode, total + 64));
    k_qtile_scatter<<<(unsigned)((s.nrows + 7) / 8), 256, 0, st>>>(s.nrows, ntiles, TR, s.lptr, s.lcode, s.hptr, s.hcode, s.hcnt, tptr, d_ulo, d_uho,
                                                                    t.tcode);
    ctx->launches++;
    t.ncta = ncta;
    t.ngroups = ngroups;
    t.nseg = ncta + ntiles;
    int64_t *d_cu = nullptr, *d_lo = nullptr, *d_hi = nullptr;
    SCS_TRY(tmps.get(&d_cu, ncta + 1));
    SCS_TRY(tmps.get(&d_lo, t.nseg));
    SCS_TRY(tmps.get(&d_hi, t.nseg));
    SCS_TRY(dev_alloc(ctx, &t.seg_ptr, ncta + 1));
    SCS_TRY(dev_alloc(ctx, &t.seg_tile, t.nseg));
    SCS_TRY(dev_alloc(ctx, &t.gb, (int64_t)t.nseg * (ngroups + 1)));
    // what running into a second tile costs a CTA, in codes: measured at config 3 (E-step sums 118.2 us with 0, 117.5 / 116.7 /
    // 114.5 / 109.6 / 112.4 / 116.1 us with 4k / 12k / 20k / 30k / 50k / 150k): far more than the 5 us of the load itself -- the CTA
    // drains at the segment's end and every group starts cold again.  tile bytes / 8 = 29k for the 227 KB tile.
    const char* tc_env = std::getenv("SCS_Q_TILE_COST");
    const int64_t tile_cost = tc_env ? atoll(tc_env) : (int64_t)(TR + 2) * lanes * 16 / 8;
    const char* uc_env = std::getenv("SCS_Q_UNIT_COST");
    const int64_t unit_cost = uc_env ? atoll(uc_env) : UNIT_COST;
    k_plan_ctas<<<(ncta + 1 + 127) / 128, 128, 0, st>>>(tptr, nunits, ncta, d_cu, s.nrows, tile_cost, unit_cost);
    k_plan_segments<<<1, 1024, 0, st>>>(d_cu, ncta, s.nrows, t.nseg, t.seg_ptr, t.seg_tile, d_lo, d_hi);
    k_plan_groups<<<t.nseg, 256, 0, st>>>(tptr, d_lo, d_hi, ngroups, t.gb, unit_cost);
    ctx->launches += 3;
    if (lanes == 4 && !std::getenv("SCS_Q_NO_ORDER")) {
        uint8_t* d_phase = nullptr;
        SCS_TRY(tmps.get(&d_phase, nunits));
        SCS_CUDA(cudaMemsetAsync(d_phase, 0, (size_t)nunits, st));
        const int64_t runs = (int64_t)t.nseg * ngroups;
        k_unit_phase<<<(unsigned)((runs + 255) / 256), 256, 0, st>>>(tptr, t.gb, ngroups, 32 / lanes, t.nseg, d_phase);
        const int64_t avg = std::max<int64_t>(2, (total + nunits) / std::max<int64_t>(nunits, 1));
        const int upw = (int)std::min<int64_t>(32, std::max<int64_t>(1, (int64_t)ORDER_CAP * 2 / 3 / avg));
        const int64_t owarps = (nunits + upw - 1) / upw;
        k_qorder_units<<<(unsigned)((owarps + ORDER_WARPS - 1) / ORDER_WARPS), ORDER_WARPS * 32, 0, st>>>(tptr, t.tcode, d_phase, nunits, upw, TR);
        ctx->launches += 2;
    }
    if (lanes == 4) k_qpack_words<4><<<(unsigned)((total / 2 + 255) / 256), 256, 0, st>>>(total / 2, reinterpret_cast<uint32_t*>(t.tcode));
    else k_qpack_words<8><<<(unsigned)((total / 2 + 255) / 256), 256, 0, st>>>(total / 2, reinterpret_cast<uint32_t*>(t.tcode));
    ctx->launches++;
    SCS_CUDA(cudaGetLastError());
    *ok = 1;
    return 0;
}

// SCS_TRACE=1: host-side phase times of the layout build on stderr (each phase ends with a stream synchronise)
struct PhaseTrace {
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseTrace(cudaStream_t s) : on(std::getenv("SCS_TRACE") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[scs trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ---- lock-step stream of the fixed-point E-step (estep_lk.cuh) ----------------------------------------------------
// The units (cell, table tile) of a tile are sorted by length and dealt 32 at a time to tasks; a task's 32 units are read by
// the 32 lanes of a warp side by side, so they are padded to one length.  For 64-byte rows the lanes l and l + 4 of a quarter
// warp read the same 16-byte chunk position of their rows at every step, which is conflict-free exactly when the two rows lie
// in different halves of a 128-byte bank line, i.e. differ in parity.  Units are therefore PAIRED: A (lanes 0..3 of a
// quarter) lists its even rows first and then its odd rows, B (lanes 4..7) its odd rows first and then its even rows; with
// e / o the even / odd codes of a unit, the pair needs max(eA, oB) + max(oA, eB) positions, the shorter list of either part
// padded by codes of the all-zero row of the right parity (rows TR and TR + 1 of the tile are zero).  That is eA + oA when B
// is A's mirror image, so inside every class of equal length the units are sorted by e - o and the i-th smallest is paired
// with the i-th largest.  The pairs of a tile are then sorted by the positions they need and cut into tasks of 16 pairs.
// The order inside a unit does not matter: the sums are integers.

constexpr int LK_FIELD = 21;                                   // bits of the length / imbalance fields of a sort key
constexpr uint64_t LK_FMASK = (1ull << LK_FIELD) - 1;
constexpr uint64_t LK_NONE = ~0ull;                            // key of an empty unit / of a unit that is the second of its pair

// one thread per unit: key = tile | codes | (even - odd + 2^20), the unit's id, its codes in even rows
__global__ void k_lk_unit_keys(int64_t nunits, int64_t ncells, int lanes, const int64_t* __restrict__ tptr, const uint16_t* __restrict__ tcode,
                               uint64_t* __restrict__ key, uint32_t* __restrict__ val, int32_t* __restrict__ ue) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nunits) return;
    const int64_t b = tptr[u], e = tptr[u + 1], len = e - b;
    val[u] = (uint32_t)u;
    if (len == 0) { key[u] = LK_NONE; ue[u] = 0; return; }
    int64_t ev = 0;
    for (int64_t i = b; i < e; ++i) ev += !(tcode[i] & 1u);
    if (lanes != 4) ev = len;                                  // a 128-byte row spans every bank: parity means nothing
    const int64_t d = 2 * ev - len;
    key[u] = ((uint64_t)(u / ncells) << (2 * LK_FIELD)) | ((uint64_t)len << LK_FIELD) | (uint64_t)((lanes == 4 ? d : 0) + (1 << (LK_FIELD - 1)));
    ue[u] = (int32_t)ev;
}

// first i in [0, n) with (key[i] >> shift) >= want (upper: > want)
__device__ __forceinline__ int64_t lk_bound(const uint64_t* __restrict__ key, int64_t n, uint64_t want, int shift, bool upper) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const uint64_t k = key[mid] >> shift;
        if (upper ? k <= want : k < want) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one thread per unit in sorted order: the i-th unit of its (tile, length) class is paired with the i-th from the class's end;
// the first of a pair gets the pair's key = tile | positions needed (even), the second (and every empty unit) LK_NONE
__global__ void k_lk_pairs(int64_t nunits, int lanes, const uint64_t* __restrict__ key_s, const uint32_t* __restrict__ val_s,
                           const int32_t* __restrict__ ue, const int64_t* __restrict__ tptr, uint64_t* __restrict__ pkey, uint32_t* __restrict__ pval,
                           int32_t* __restrict__ pair_u) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nunits) return;
    pval[i] = (uint32_t)i;
    const uint64_t k = key_s[i];
    pkey[i] = LK_NONE;
    if (k == LK_NONE) return;
    const uint64_t cls = k >> LK_FIELD;
    const int64_t s = lk_bound(key_s, nunits, cls, LK_FIELD, false), e = lk_bound(key_s, nunits, cls, LK_FIELD, true);
    const int64_t j = s + e - 1 - i;
    if (i > j) return;
    const int64_t uA = val_s[i], uB = j != i ? (int64_t)val_s[j] : -1;
    const int64_t lenA = tptr[uA + 1] - tptr[uA], eA = ue[uA], oA = lenA - eA;
    const int64_t lenB = uB >= 0 ? tptr[uB + 1] - tptr[uB] : 0, eB = uB >= 0 ? ue[uB] : 0, oB = lenB - eB;
    int64_t n = lanes == 4 ? max(eA, oB) + max(oA, eB) : max(lenA, lenB);
    n = (n + 1) & ~(int64_t)1;
    pair_u[2 * i] = (int32_t)uA;
    pair_u[2 * i + 1] = (int32_t)uB;
    pkey[i] = ((k >> (2 * LK_FIELD)) << (2 * LK_FIELD)) | ((uint64_t)n << LK_FIELD);
}

// one block: where each tile's pairs start in the sorted pair list, and its first task (16 pairs per task)
__global__ void k_lk_tiles(int ntiles, int64_t n, const uint64_t* __restrict__ pkey_s, int64_t* __restrict__ ps, int32_t* __restrict__ tile_task) {
    for (int t = threadIdx.x; t <= ntiles; t += blockDim.x) ps[t] = lk_bound(pkey_s, n, (uint64_t)t << (2 * LK_FIELD), 0, false);
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t acc = 0;
        for (int t = 0; t < ntiles; ++t) { tile_task[t] = (int32_t)acc; acc += (ps[t + 1] - ps[t] + 15) / 16; }
        tile_task[ntiles] = (int32_t)acc;
    }
}

// tile of task k: the last t with tile_task[t] <= k (empty tiles share their first task with the next tile)
__device__ __forceinline__ int lk_tile_of(const int32_t* __restrict__ tile_task, int ntiles, int64_t k) {
    int lo = 0, hi = ntiles - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tile_task[mid] <= k) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// one thread per task: stream rows (two positions each) = half the positions of its longest (= last) pair
__global__ void k_lk_task_len(int64_t ntasks, int ntiles, const int32_t* __restrict__ tile_task, const int64_t* __restrict__ ps,
                              const uint64_t* __restrict__ pkey_s, int64_t* __restrict__ rows) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ntasks) return;
    const int t = lk_tile_of(tile_task, ntiles, k);
    const int64_t last = min(ps[t] + 16 * (k - tile_task[t]) + 15, ps[t + 1] - 1);
    rows[k] = (int64_t)((pkey_s[last] >> LK_FIELD) & LK_FMASK) / 2;
}

// one thread per (task, lane): the lane's unit and its codes in the pair's order, written lane-interleaved
template <int C>
__global__ void k_lk_emit(int64_t ntasks, int ntiles, int64_t ncells, int TR, const int32_t* __restrict__ tile_task, const int64_t* __restrict__ ps,
                          const uint32_t* __restrict__ pval_s, const int32_t* __restrict__ pair_u, const int32_t* __restrict__ ue,
                          const int64_t* __restrict__ tptr, const uint16_t* __restrict__ tcode, const int64_t* __restrict__ task_row,
                          int32_t* __restrict__ task_cell, uint32_t* __restrict__ words) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k = gid >> 5;
    const int lane = (int)(gid & 31);
    if (k >= ntasks) return;
    const int t = lk_tile_of(tile_task, ntiles, k);
    const int m = (lane >> 2) & 1;                                 // 0: unit A of the pair, 1: unit B
    const int64_t p = ps[t] + 16 * (k - tile_task[t]) + ((lane >> 3) * 4 + (lane & 3));
    int64_t uA = -1, uB = -1;
    if (p < ps[t + 1]) {
        const int64_t i = pval_s[p];
        uA = pair_u[2 * i];
        uB = pair_u[2 * i + 1];
    }
    const int64_t lenA = uA >= 0 ? tptr[uA + 1] - tptr[uA] : 0, eA = uA >= 0 ? ue[uA] : 0, oA = lenA - eA;
    const int64_t lenB = uB >= 0 ? tptr[uB + 1] - tptr[uB] : 0, eB = uB >= 0 ? ue[uB] : 0, oB = lenB - eB;
    const int64_t n1 = C == 4 ? max(eA, oB) : max(lenA, lenB), n2 = C == 4 ? max(oA, eB) : 0;
    const int64_t u = m ? uB : uA;
    task_cell[k * 32 + lane] = u >= 0 ? (int32_t)(u - (int64_t)t * ncells) : -1;
    const int64_t x0 = task_row[k], nk = 2 * (task_row[k + 1] - x0);
    const uint32_t zrow[2] = {(uint32_t)(TR + (TR & 1)), (uint32_t)(TR + ((TR & 1) ^ 1))};     // the zero row of parity 0 / 1
    const int64_t b = u >= 0 ? tptr[u] : 0, e = u >= 0 ? tptr[u + 1] : 0;
    const uint32_t f = (uint32_t)m;                                // parity of the rows this lane lists first
    int64_t pos = 0;
    uint32_t half = 0;
    auto put = [&](uint32_t row) {
        const uint32_t c = LkCode<C>::code(row, lane);
        SCS_ASSERT(c < 65536u && pos < nk);
        if (pos & 1) words[(x0 + (pos >> 1)) * 32 + lane] = half | (c << 16);
        else half = c;
        ++pos;
    };
    for (int64_t i = b; i < e; ++i) {
        const uint32_t row = tcode[i];
        if (C != 4 || (row & 1u) == f) put(row);
    }
    while (pos < n1) put(zrow[f]);
    if (C == 4) {
        for (int64_t i = b; i < e; ++i) {
            const uint32_t row = tcode[i];
            if ((row & 1u) != f) put(row);
        }
        while (pos < n1 + n2) put(zrow[f ^ 1u]);
    }
    while (pos < nk) put(zrow[f]);
}

// Equal-cost ranges of stream rows for the CTAs.  Running into another table tile costs a CTA far more than the load itself
// (it drains, loads 227 KB, and every warp starts cold): a range is charged `charge` rows for every tile boundary inside it.
__global__ void k_lk_plan(int ncta, int ntiles, int64_t nrows, const int32_t* __restrict__ tile_task, const int64_t* __restrict__ task_row,
                          int64_t charge, int64_t* __restrict__ cta_row, int32_t* __restrict__ cta_tile) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > ncta) return;
    auto tile_of = [&](int64_t x) {                                // the last non-empty tile that starts at or before row x
        int t = 0;
        for (int u = 1; u < ntiles; ++u)
            if (tile_task[u] < tile_task[u + 1] && task_row[tile_task[u]] <= x) t = u;
        return t;
    };
    auto cost = [&](int64_t x) {                                   // of the rows [0, x)
        int64_t c = x;
        for (int t = 1; t < ntiles; ++t) {
            const int64_t r = task_row[tile_task[t]];
            if (tile_task[t] < tile_task[t + 1] && r > 0 && r < x) c += charge;
        }
        return c;
    };
    if (j == 0 || j == ncta) { cta_row[j] = j ? nrows : 0; if (j == 0) cta_tile[0] = tile_of(0); return; }
    const int64_t total = cost(nrows), target = (total * j + ncta - 1) / ncta;
    int64_t lo = 0, hi = nrows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cost(mid) < target) lo = mid + 1; else hi = mid;
    }
    // a target inside the jump of a tile boundary: cut AT the boundary instead of one row into the next tile
    if (lo > 0 && cost(lo) - cost(lo - 1) > 1 && cost(lo - 1) < target) --lo;
    cta_row[j] = lo;
    cta_tile[j] = tile_of(lo);
}

// one thread per cell: entries whose count is too large to be written as repeated codes
__global__ void k_lk_xheavy(int64_t ncells, const int64_t* __restrict__ hptr, const int32_t* __restrict__ hcnt, int32_t* __restrict__ xheavy) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    int n = 0;
    for (int64_t h = hptr[c]; h < hptr[c + 1]; ++h) n += hcnt[h] > Q_REP_MAX;
    xheavy[c] = n;
}

void free_lockstep(scs_ctx* ctx, LockstepStream& t) {
    dev_free(ctx, t.words);
    dev_free(ctx, t.task_row);
    dev_free(ctx, t.task_cell);
    dev_free(ctx, t.tile_task);
    dev_free(ctx, t.cta_row);
    dev_free(ctx, t.cta_tile);
    dev_free(ctx, t.xheavy);
    t = LockstepStream();
}

template <typename KT, typename VT>
static int lk_sort_pairs(cudaStream_t st, Scratch& tmps, const KT* kin, KT* kout, const VT* vin, VT* vout, int64_t n, int b0, int b1) {
    size_t tb = 0;
    SCS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, kin, kout, vin, vout, n, b0, b1, st));
    void* tmp = nullptr;
    SCS_TRY(tmps.get_bytes(&tmp, tb));
    SCS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, kin, kout, vin, vout, n, b0, b1, st));
    return 0;
}

int build_lockstep(scs_ctx* ctx, const Stream& s, LockstepStream& t, int lanes, int TR, int ncta, int* F, int* ok) {
    free_lockstep(ctx, t);
    *ok = 0;
    SCS_CHECK(ncta >= 1 && ncta <= 4096, "work plan supports 1..4096 CTAs (got %d)", ncta);
    cudaStream_t st = ctx->stream;
    const int ntiles = (int)((s.xrows + TR - 1) / TR);
    const int64_t nunits = (int64_t)ntiles * s.nrows;
    SCS_CHECK(nunits < (1LL << 31) && ntiles < (1 << (LK_FIELD - 1)), "pool too large for the lock-step stream (%lld units)", (long long)nunits);
    Scratch tmps(st);
    PhaseTrace trace(st);
    int32_t *d_ulo = nullptr, *d_uho = nullptr;
    int64_t *tptr = nullptr, *d_hv = nullptr;
    SCS_TRY(tmps.get(&d_ulo, nunits));
    SCS_TRY(tmps.get(&d_uho, nunits));
    SCS_TRY(tmps.get(&d_hv, 2));
    SCS_TRY(tmps.get(&tptr, nunits + 1));
    SCS_CUDA(cudaMemsetAsync(d_hv, 0, 16, st));
    SCS_CUDA(cudaMemsetAsync(tptr, 0, sizeof(int64_t), st));
    const unsigned ugrid = (unsigned)((nunits + 255) / 256);
    k_qunit_counts<<<ugrid, 256, 0, st>>>(s.nrows, ntiles, TR, s.lptr, s.lcode, s.hptr, s.hcode, s.hcnt, tptr + 1, d_ulo, d_uho, 1);
    k_qrow_depth<<<(unsigned)((s.nrows + 255) / 256), 256, 0, st>>>(s.nrows, s.lptr, s.hptr, s.hcnt, reinterpret_cast<unsigned long long*>(d_hv));
    {
        size_t tb = 0;
        SCS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, tptr + 1, tptr + 1, nunits, st));
        void* tmp = nullptr;
        SCS_TRY(tmps.get_bytes(&tmp, tb));
        SCS_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb, tptr + 1, tptr + 1, nunits, st));
    }
    ctx->launches += 4;
    int64_t h_max = 0, total = 0;
    SCS_CUDA(cudaMemcpyAsync(&h_max, d_hv, 8, cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaMemcpyAsync(&total, tptr + nunits, 8, cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    *F = (uint64_t)h_max < (1ull << 40) ? q_frac_bits(h_max) : 0;
    if (*F < Q_MIN_F) return 0;                          // cells too deep for a useful fraction: fp64 keeps the E-step
    trace.mark("lk: unit counts + scan");
    uint16_t* tcode = nullptr;
    SCS_TRY(tmps.get(&tcode, total + 2));
    k_qtile_scatter<<<(unsigned)((s.nrows + 7) / 8), 256, 0, st>>>(s.nrows, ntiles, TR, s.lptr, s.lcode, s.hptr, s.hcode, s.hcnt, tptr, d_ulo, d_uho, tcode);
    // units -> sorted by (tile, length, even - odd) -> pairs -> sorted by (tile, positions) -> tasks
    uint64_t *key = nullptr, *key_s = nullptr, *pkey = nullptr, *pkey_s = nullptr;
    uint32_t *val = nullptr, *val_s = nullptr, *pval = nullptr, *pval_s = nullptr;
    int32_t *ue = nullptr, *pair_u = nullptr;
    int64_t* ps = nullptr;
    SCS_TRY(tmps.get(&key, nunits)); SCS_TRY(tmps.get(&key_s, nunits));
    SCS_TRY(tmps.get(&val, nunits)); SCS_TRY(tmps.get(&val_s, nunits));
    SCS_TRY(tmps.get(&ue, nunits));
    SCS_TRY(tmps.get(&pair_u, 2 * nunits));
    SCS_TRY(tmps.get(&ps, ntiles + 1));
    trace.mark("lk: scatter codes");
    k_lk_unit_keys<<<ugrid, 256, 0, st>>>(nunits, s.nrows, lanes, tptr, tcode, key, val, ue);
    trace.mark("lk: unit keys");
    const int top = 2 * LK_FIELD + bits_for(ntiles + 1);  // (the key of an empty unit is all ones: it sorts behind every tile)
    SCS_TRY(lk_sort_pairs(st, tmps, key, key_s, val, val_s, nunits, 0, top));
    trace.mark("lk: sort units");
    pkey = key; pval = val;                               // (the unsorted unit keys are not needed any more)
    SCS_TRY(tmps.get(&pkey_s, nunits));
    SCS_TRY(tmps.get(&pval_s, nunits));
    k_lk_pairs<<<ugrid, 256, 0, st>>>(nunits, lanes, key_s, val_s, ue, tptr, pkey, pval, pair_u);
    trace.mark("lk: pairs");
    SCS_TRY(lk_sort_pairs(st, tmps, pkey, pkey_s, pval, pval_s, nunits, LK_FIELD, top));
    trace.mark("lk: sort pairs");
    SCS_TRY(dev_alloc(ctx, &t.tile_task, ntiles + 1));
    k_lk_tiles<<<1, 256, 0, st>>>(ntiles, nunits, pkey_s, ps, t.tile_task);
    ctx->launches += 12;
    int32_t h_ntasks = 0;
    SCS_CUDA(cudaMemcpyAsync(&h_ntasks, t.tile_task + ntiles, 4, cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    const int64_t ntasks = h_ntasks;
    SCS_TRY(dev_alloc(ctx, &t.task_row, ntasks + 1));
    SCS_TRY(dev_alloc(ctx, &t.task_cell, (ntasks + 1) * 32));
    SCS_CUDA(cudaMemsetAsync(t.task_row, 0, sizeof(int64_t), st));
    int64_t nrows = 0;
    if (ntasks > 0) {
        k_lk_task_len<<<(unsigned)((ntasks + 255) / 256), 256, 0, st>>>(ntasks, ntiles, t.tile_task, ps, pkey_s, t.task_row + 1);
        size_t tb = 0;
        SCS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, t.task_row + 1, t.task_row + 1, ntasks, st));
        void* tmp = nullptr;
        SCS_TRY(tmps.get_bytes(&tmp, tb));
        SCS_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb, t.task_row + 1, t.task_row + 1, ntasks, st));
        SCS_CUDA(cudaMemcpyAsync(&nrows, t.task_row + ntasks, 8, cudaMemcpyDeviceToHost, st));
        SCS_CUDA(cudaStreamSynchronize(st));
        ctx->launches += 3;
    }
    trace.mark("lk: tasks");
    SCS_TRY(dev_alloc(ctx, &t.words, (nrows + 2 * LK_RING) * 32));     // (the kernel requests rows ahead without looking)
    SCS_CUDA(cudaMemsetAsync(t.words + nrows * 32, 0, (size_t)2 * LK_RING * 32 * 4, st));
    if (ntasks > 0) {
        const unsigned egrid = (unsigned)((ntasks * 32 + 255) / 256);
        if (lanes == 4) k_lk_emit<4><<<egrid, 256, 0, st>>>(ntasks, ntiles, s.nrows, TR, t.tile_task, ps, pval_s, pair_u, ue, tptr, tcode, t.task_row, t.task_cell, t.words);
        else k_lk_emit<8><<<egrid, 256, 0, st>>>(ntasks, ntiles, s.nrows, TR, t.tile_task, ps, pval_s, pair_u, ue, tptr, tcode, t.task_row, t.task_cell, t.words);
        ctx->launches++;
    }
    trace.mark("lk: emit");
    SCS_TRY(dev_alloc(ctx, &t.cta_row, ncta + 1));
    SCS_TRY(dev_alloc(ctx, &t.cta_tile, ncta + 1));
    // (in stream rows of 64 codes; sweep at config 3, E-step sums: 95.2 us with no charge, 83.9 at 453, 82.8 at 1000, 84.8 at 2000, 88.0 at 3000)
    const char* tc_env = std::getenv("SCS_LK_TILE_COST");
    const int64_t charge = tc_env ? atoll(tc_env) : (int64_t)(TR + 2) * lanes * 16 / 256;
    k_lk_plan<<<(ncta + 1 + 127) / 128, 128, 0, st>>>(ncta, ntiles, nrows, t.tile_task, t.task_row, charge, t.cta_row, t.cta_tile);
    SCS_TRY(dev_alloc(ctx, &t.xheavy, s.nrows));
    k_lk_xheavy<<<(unsigned)((s.nrows + 255) / 256), 256, 0, st>>>(s.nrows, s.hptr, s.hcnt, t.xheavy);
    ctx->launches += 2;
    SCS_CUDA(cudaGetLastError());
    trace.mark("lk: plan + xheavy");
    t.lanes = lanes;
    t.TR = TR;
    t.ntiles = ntiles;
    t.ntasks = ntasks;
    t.nrows = nrows;
    t.ncta = ncta;
    *ok = 1;
    return 0;
}

int build_layout(scs_ctx* ctx, const int64_t* ref_indptr, const int32_t* ref_indices, const void* ref_data,
                 const int64_t* alt_indptr, const int32_t* alt_indices, const void* alt_data, int data_bytes) {
    const int64_t V = ctx->V, N = ctx->N;
    SCS_CHECK(data_bytes == 2 || data_bytes == 4 || data_bytes == 8, "data_bytes must be 2, 4 or 8 (got %d)", data_bytes);
    SCS_CHECK(V > 0 && N > 0, "empty matrix: V=%lld N=%lld", (long long)V, (long long)N);
    SCS_CHECK(V < (1LL << 29) && N < (1LL << 30), "V must be below 2^29 and N below 2^30");
    SCS_CHECK(ref_indptr[0] == 0 && alt_indptr[0] == 0, "indptr[0] must be 0");
    const int64_t nR = ref_indptr[V], nA = alt_indptr[V];
    SCS_CHECK(nR >= 0 && nA >= 0, "negative nnz");
    ctx->nnz_ref = nR;
    ctx->nnz_alt = nA;
    cudaStream_t st = ctx->stream;

    // ---- host -> device (raw CSR of both matrices); temporaries: returned to the pool on every path out of this function
    Scratch tmp(st);
    int64_t *a_ptr = nullptr, *r_ptr = nullptr;
    int32_t *a_idx = nullptr, *r_idx = nullptr;
    void *a_val = nullptr, *r_val = nullptr;
    SCS_TRY(tmp.get(&a_ptr, V + 1));
    SCS_TRY(tmp.get(&r_ptr, V + 1));
    SCS_TRY(tmp.get(&a_idx, nA));
    SCS_TRY(tmp.get(&r_idx, nR));
    SCS_TRY(tmp.get_bytes(&a_val, (size_t)nA * data_bytes));
    SCS_TRY(tmp.get_bytes(&r_val, (size_t)nR * data_bytes));
    SCS_CUDA(cudaMemcpyAsync(a_ptr, alt_indptr, (V + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    SCS_CUDA(cudaMemcpyAsync(r_ptr, ref_indptr, (V + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (nA) SCS_CUDA(cudaMemcpyAsync(a_idx, alt_indices, (size_t)nA * 4, cudaMemcpyHostToDevice, st));
    if (nR) SCS_CUDA(cudaMemcpyAsync(r_idx, ref_indices, (size_t)nR * 4, cudaMemcpyHostToDevice, st));
    if (nA) SCS_CUDA(cudaMemcpyAsync(a_val, alt_data, (size_t)nA * data_bytes, cudaMemcpyHostToDevice, st));
    if (nR) SCS_CUDA(cudaMemcpyAsync(r_val, ref_data, (size_t)nR * data_bytes, cudaMemcpyHostToDevice, st));

    PhaseTrace trace(st);
    trace.mark("alloc + H2D");
    int* err = nullptr;
    unsigned long long* ndup = nullptr;
    SCS_TRY(tmp.get(&err, 1));
    SCS_TRY(tmp.get(&ndup, 1));
    SCS_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    SCS_CUDA(cudaMemsetAsync(ndup, 0, sizeof(unsigned long long), st));

    // ---- SNV-major stream (2V pseudo-rows)
    Stream& M = ctx->M;
    M.nrows = 2 * V;
    M.xrows = N;
    int64_t *cntL = nullptr, *cntH = nullptr, *rowsum = nullptr;
    SCS_TRY(tmp.get(&cntL, 2 * V));
    SCS_TRY(tmp.get(&cntH, 2 * V));
    SCS_TRY(tmp.get(&rowsum, 2 * V));
    const int WPB = 8;  // warps per block
    unsigned gridPR = (unsigned)((2 * V + WPB - 1) / WPB);
    k_count_rows<<<gridPR, WPB * 32, 0, st>>>(V, a_ptr, a_val, r_ptr, r_val, data_bytes, cntL, cntH, rowsum, err);
    ctx->launches++;
    trace.mark("count_rows");
    SCS_TRY(dev_alloc(ctx, &M.lptr, 2 * V + 1));
    SCS_TRY(dev_alloc(ctx, &M.hptr, 2 * V + 1));
    SCS_TRY(inclusive_scan_into_ptr(ctx, cntL, M.lptr, 2 * V));
    SCS_TRY(inclusive_scan_into_ptr(ctx, cntH, M.hptr, 2 * V));
    trace.mark("scans");
    int herr = 0;
    SCS_CUDA(cudaMemcpy(&herr, err, sizeof(int), cudaMemcpyDeviceToHost));
    SCS_CHECK(herr == 0, "count matrices must hold strictly positive counts that fit int32 (explicit zeros are not allowed)");
    SCS_CUDA(cudaMemcpy(&M.n_light, M.lptr + 2 * V, sizeof(int64_t), cudaMemcpyDeviceToHost));
    SCS_CUDA(cudaMemcpy(&M.n_heavy, M.hptr + 2 * V, sizeof(int64_t), cudaMemcpyDeviceToHost));
    SCS_CHECK(M.n_light + M.n_heavy == nA + nR, "internal: light+heavy != nnz");
    SCS_TRY(dev_alloc(ctx, &M.lcode, M.n_light + 4));   // + slack: the sparse M-step reads whole 16-byte vectors
    SCS_TRY(dev_alloc(ctx, &M.hcode, M.n_heavy));
    SCS_TRY(dev_alloc(ctx, &M.hcnt, M.n_heavy));

    uint32_t *lkey = nullptr, *lval = nullptr, *hkey = nullptr, *lkey_s = nullptr, *hkey_s = nullptr;
    uint64_t *hval = nullptr, *hval_s = nullptr;
    const int64_t nl = M.n_light ? M.n_light : 1, nh = M.n_heavy ? M.n_heavy : 1;
    SCS_TRY(tmp.get(&lkey, nl));
    SCS_TRY(tmp.get(&lval, nl));
    SCS_TRY(tmp.get(&lkey_s, nl));
    SCS_TRY(tmp.get(&hkey, nh));
    SCS_TRY(tmp.get(&hkey_s, nh));
    SCS_TRY(tmp.get(&hval, nh));
    SCS_TRY(tmp.get(&hval_s, nh));
    k_fill_M<<<gridPR, WPB * 32, 0, st>>>(V, N, a_ptr, a_idx, a_val, r_ptr, r_idx, r_val, data_bytes, M.lptr, M.hptr,
                                          M.lcode, M.hcode, M.hcnt, lkey, lval, hkey, hval, ndup, err);
    ctx->launches++;

    trace.mark("allocs + fill_M");
    // ---- cell-major stream: stable sort of the SNV-major order by cell
    Stream& E = ctx->E;
    E.nrows = N;
    E.xrows = 2 * V;
    E.n_light = M.n_light;
    E.n_heavy = M.n_heavy;
    SCS_TRY(dev_alloc(ctx, &E.lptr, N + 1));
    SCS_TRY(dev_alloc(ctx, &E.hptr, N + 1));
    SCS_TRY(dev_alloc(ctx, &E.lcode, E.n_light));
    SCS_TRY(dev_alloc(ctx, &E.hcode, E.n_heavy));
    SCS_TRY(dev_alloc(ctx, &E.hcnt, E.n_heavy));
    SCS_TRY(stable_sort_by_cell<uint32_t>(ctx, lkey, lkey_s, lval, E.lcode, M.n_light, N));
    SCS_TRY(stable_sort_by_cell<uint64_t>(ctx, hkey, hkey_s, hval, hval_s, M.n_heavy, N));
    trace.mark("allocs + sorts");
    k_ptr_from_sorted<<<(unsigned)((M.n_light + 256) / 256), 256, 0, st>>>(lkey_s, M.n_light, N, E.lptr);
    k_ptr_from_sorted<<<(unsigned)((M.n_heavy + 256) / 256), 256, 0, st>>>(hkey_s, M.n_heavy, N, E.hptr);
    if (M.n_heavy) k_split_heavy<<<(unsigned)((M.n_heavy + 255) / 256), 256, 0, st>>>(hval_s, M.n_heavy, E.hcode, E.hcnt);
    ctx->launches += 3;

    // ---- integer row sums and k_alt (scSplit:100-103)
    SCS_TRY(dev_alloc(ctx, &ctx->sum_alt, V));
    SCS_TRY(dev_alloc(ctx, &ctx->sum_ref, V));
    SCS_TRY(dev_alloc(ctx, &ctx->kalt, V));
    k_row_sums<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(V, rowsum, ctx->sum_alt, ctx->sum_ref);
    k_kalt<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(V, ctx->sum_alt, ctx->sum_ref, ctx->kalt);
    ctx->launches += 2;

    unsigned long long hdup = 0;
    SCS_CUDA(cudaMemcpyAsync(&hdup, ndup, sizeof(hdup), cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    SCS_CUDA(cudaGetLastError());
    trace.mark("ptrs, sums, kalt");
    SCS_CHECK(herr != 2, "column index out of range [0, N)");
    SCS_CHECK(herr != 3, "CSR column indices must be sorted and duplicate-free within each row");
    SCS_CHECK(herr == 0, "invalid count matrix (code %d)", herr);
    ctx->nnz_union = nA + nR - (int64_t)hdup;
    return 0;
}

}  // namespace scs
