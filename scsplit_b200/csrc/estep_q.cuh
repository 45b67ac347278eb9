// E-step sums in exact fixed point (DESIGN.md "E-step, fixed-point tables").
//
//   L[c,k] = sum over the entries of cell c of cnt * T[g,k]          (scSplit:175-178; T = log2 P, log2(1-P))
//
// The fp64 SpMM (spmm_tiled.cuh) is bound by the shared-memory data pipe: an 80-byte fp64 row per entry is one 128-byte
// wavefront per entry.  Here k_finalize also writes the log table as 55-bit fixed point, q = rint(-x * 2^F), packed so that
// a row of K <= 9 values is 64 bytes (two rows per wavefront) and a row of K <= 17 values is 128 bytes; the sums are
// accumulated as INTEGERS, which is exact: no rounding at all between the table and L, and any summation order gives the
// same bits.  F is chosen per context so that a cell's whole sum fits 63 bits (F = 46 at 30k x 20k): every table entry
// carries an absolute error of at most 2^-(F+1), and then nothing is lost any more -- an fp64 accumulation of ~1500 terms
// of magnitude ~1 rounds at 2^-43 per add.  tests/test_gpu_qpath.py measures both paths against extended-precision sums.
// A table that does not fit the format (NaN / -inf of a dying cluster, |x| >= 64) raises the restart's `qbad` flag in
// k_finalize and that restart's rows are gathered from the fp64 table by k_bayes for that iteration.
//
// Row layout (LANES = 4 or 8 lanes of 16 bytes; lane j holds values a = 2j, b = 2j+1 and one byte pair of the "extra"
// value x = 2*LANES when K is odd-full (K = 9 / 17)), as two 64-bit words:
//     q0 = a | pa << 56        q1 = b | pb << 56          (a, b, x < 2^55; pa = byte 2j of x, pb = byte 2j+1; byte 7 = 0)
// The hot loop adds q0 and q1 as plain 64-bit integers (the byte on top wraps around harmlessly: a cell's sum of a stays
// below 2^63, so sum(a) = sum(q0) - (sum(pa) << 56) mod 2^64) and the byte pair into a third sum (PRMT puts pa at bit 0 and
// pb at bit 24; a unit holds < 65536 codes, which keeps the low field below 2^24).  Two 64-bit adds on the integer pipe
// and one IMAD.WIDE on the multiply pipe per 16 bytes -- an IMAD.WIDE per 32-bit word (which would also carry a per-entry
// count) measured multiply-pipe bound at 2.7x the shared-memory floor.  Counts therefore are not multiplied: an entry of
// count c <= Q_REP_MAX is c codes, larger counts (rare in droplet data) are added in fp64 by k_combine_q.
//
// Stream: 32-bit words of two codes, word = rowA << SA | rowB << (SA + 14), SA = log2(row bytes): the first row's byte
// offset is one AND away, the second's a shift and an AND.  Padding codes point at one of the two all-zero rows TR, TR + 1
// (one of each parity: a padding code never costs a bank conflict, see k_qorder_units).
#pragma once
#include "common.cuh"
#include "spmm_tiled.cuh"

namespace scs {

constexpr int Q_REP_MAX = 8;         // largest count written as repeated codes; larger counts go through the fp64 leftover loop
constexpr int Q_INT_BITS = 6;        // |x| < 64
constexpr int Q_MAX_F = 49;          // 6 + 49 = 55 bits per value
constexpr int Q_MIN_F = 40;
constexpr int Q_UNIT_PAD = 8;        // units hold a multiple of this many codes (the kernel looks for a unit's end once per 8 codes)          // pools whose deepest cell would push F below this keep the fp64 E-step

inline int q_lanes_for(int K) { return K <= 9 ? 4 : K <= 17 ? 8 : 0; }
// shared memory of k_estep_q: table tile (TR + 2 rows) | mbarrier
constexpr int Q_SMEM_BYTES = 227 * 1024;
inline int q_tile_rows(int lanes, int64_t xrows) {
    int64_t tr = (Q_SMEM_BYTES - 256) / (lanes * 16) - 2;      // + the two zero rows
    const int64_t cap = lanes == 4 ? 4094 : 2046;    // 12 / 11 bits of in-tile row per code; rows TR and TR + 1 are all zeros
    if (tr > cap) tr = cap;
    if (tr > xrows) tr = xrows;
    if (tr < 1) tr = 1;
    // equal tiles: a short last tile means short units there (3 % of a tile at a 60k-SNV pool with 64-byte rows: units of 8
    // codes, flushed every 4 stream rows -- the CTA that got them ran 60 % longer than the others)
    const int64_t ntiles = (xrows + tr - 1) / tr;
    tr = (xrows + ntiles - 1) / ntiles;
    return (int)(tr < 1 ? 1 : tr);
}
// fraction bits for a pool whose deepest cell holds `depth` reads: depth * 2^(6+F) < 2^63
inline int q_frac_bits(int64_t depth) {
    int b = 0;
    while ((1LL << b) <= depth && b < 62) ++b;       // depth < 2^b
    const int f = 63 - Q_INT_BITS - b;
    return f > Q_MAX_F ? Q_MAX_F : f;
}

// per-unit sums of k_estep_q, word-major: [sum q0 of lanes 0..L-1 | sum q1 of lanes 0..L-1 | byte-pair sums of lanes 0..3]
// (lanes 4..7 of a 128-byte row carry no byte pair): 12 words per unit at 64-byte rows, 20 at 128-byte rows
__host__ __device__ constexpr int q_part_words(int lanes) { return lanes == 4 ? 12 : 20; }

// quantise one table value; *bad is raised when it does not fit the format
__device__ __forceinline__ uint64_t q_quant(double x, double scale, bool* bad) {
    const bool ok = (x <= 0.0) && (x > -(double)(1 << Q_INT_BITS));      // false for NaN, -inf, positive values
    if (!ok) *bad = true;
    return ok ? (uint64_t)__double2ll_rn(-x * scale) : 0ull;
}
// the 16 bytes of lane j of a row: values a, b and bytes 2j, 2j+1 of the extra value
__device__ __forceinline__ uint4 q_pack_lane(uint64_t a, uint64_t b, uint64_t x, int j) {
    const uint32_t pa = j < 4 ? (uint32_t)(x >> (16 * j)) & 0xffu : 0u;
    const uint32_t pb = j < 4 ? (uint32_t)(x >> (16 * j + 8)) & 0xffu : 0u;
    return make_uint4((uint32_t)a, (uint32_t)(a >> 32) | (pa << 24), (uint32_t)b, (uint32_t)(b >> 32) | (pb << 24));
}

__device__ __forceinline__ uint4 lds_u32x4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// byte 3 of a -> bits 0..7, byte 3 of b -> bits 24..31, zeros between (sign replication of byte 2 of a, whose top bit is
// bit 55 of a value below 2^55, i.e. always clear)
__device__ __forceinline__ uint32_t q_pieces(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, 0x7aa3;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

struct QAcc {
    uint64_t a, b, pp;
    __device__ __forceinline__ void zero() { a = b = pp = 0ull; }
    __device__ __forceinline__ void add(const uint4& v) {
        a += (uint64_t)v.x | ((uint64_t)v.y << 32);     // IADD3 + IADD3.X
        b += (uint64_t)v.z | ((uint64_t)v.w << 32);
        pp += (uint64_t)q_pieces(v.y, v.w) * 1u;        // (kept on the multiply pipe: IMAD.WIDE)
    }
};

// stream word of two codes (rows inside the tile); SA = log2(row bytes)
template <int LANES> struct QCode {
    static constexpr int SA = LANES == 4 ? 6 : 7;
    static constexpr uint32_t MASK = (LANES == 4 ? 0xfffu : 0x7ffu) << SA;   // byte offset of a row inside the tile
    __host__ __device__ static uint32_t word(uint32_t rowA, uint32_t rowB) { return (rowA << SA) | (rowB << (SA + 14)); }
};

// grid = (CTAs, R), the static work plan of spmm_tiled.cuh (CTA -> tile segments -> one run of units per row group).
// A group of LANES lanes consumes its run as a continuous stream, one entry per LDS.128; a unit is flushed by a (rare,
// divergent) branch where it ends -- units are padded to whole blocks of 8 codes, so that is looked for once per 8 codes --
// so there is one accumulator set and no per-entry predication.
// part[unit * q_part_words + {0, LANES, 2 * LANES} + lane] = sum q0, sum q1, byte-pair sums (raw; k_combine_q takes them apart).
template <int LANES>
__global__ void __launch_bounds__(TILED_THREADS, 1)
k_estep_q(int TR, int64_t xrows, const int64_t* __restrict__ tptr, const uint32_t* __restrict__ tword,
          const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_tile, const int64_t* __restrict__ gb,
          const uint64_t* __restrict__ Xq, int64_t xq_stride, uint64_t* __restrict__ part, int64_t part_stride,
          const int32_t* __restrict__ done, const int32_t* __restrict__ qbad) {
    using QC = QCode<LANES>;
    constexpr int ROWB = LANES * 16, GPW = 32 / LANES, BE = 8 * LANES;   // BE: codes per batch of one group
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (((size_t)(TR + 2) * ROWB + 127) & ~(size_t)127));

    const int r = blockIdx.y;
    if (done && done[r]) return;
    if (qbad[r]) return;                               // this restart's table does not fit: the fp64 kernel runs instead
    Xq += (int64_t)r * xq_stride;
    part += (int64_t)r * part_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = lane % LANES, g = lane / LANES;
    const int ngroups = (int)(blockDim.x >> 5) * GPW;
    const int G = warp * GPW + g;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t lane_off = (uint32_t)li * 16u;     // ORed into the row's byte offset (rows are ROWB-aligned)
    // padding for positions beyond a group's run: zero rows of the parities this group's positions want (G + p) & 1
    const uint32_t zero_word = QC::word((uint32_t)(TR + ((TR ^ G) & 1)), (uint32_t)(TR + ((TR ^ G ^ 1) & 1)));

    if (tid == 0) mbar_init(bar, 1);
    for (int k = tid; k < 2 * ROWB / 8; k += blockDim.x) reinterpret_cast<uint64_t*>(smem + (size_t)TR * ROWB)[k] = 0ull;
    __syncthreads();
    uint32_t phase = 0;
    int loaded_tile = -1;

    const int seg_end = seg_ptr[blockIdx.x + 1];
    for (int seg = seg_ptr[blockIdx.x]; seg < seg_end; ++seg) {
        const int t = seg_tile[seg];
        if (t != loaded_tile) {
            __syncthreads();                              // every warp is done reading the previous tile
            if (tid == 0) {
                const int64_t rows_t = min((int64_t)TR, xrows - (int64_t)t * TR);
                const uint32_t bytes = (uint32_t)(rows_t * ROWB);
                const char* src = reinterpret_cast<const char*>(Xq) + (int64_t)t * TR * ROWB;
                fence_proxy_async();
                mbar_expect_tx(bar, bytes);
                for (uint32_t off = 0; off < bytes; off += 32768u)
                    bulk_g2s(smem + off, src + off, min(32768u, bytes - off), bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            loaded_tile = t;
        }
        const int64_t* gbs = gb + (int64_t)seg * (ngroups + 1);
        int64_t u = __ldg(gbs + G);
        const int64_t uend = __ldg(gbs + G + 1);
        const bool act = u < uend;
        const int64_t e0 = act ? __ldg(tptr + u) : 0;                            // start of this group's run (even)
        const int run = act ? (int)(__ldg(tptr + uend) - e0) : 0;                // codes in the run (even)
        const uint32_t* words = tword + (e0 >> 1);
        // ends (relative to e0) of units u, u+1, u+2, clamped to the run
        int b1 = act ? (int)(__ldg(tptr + u + 1) - e0) : 0x7fffffff;
        int b2 = (u + 2 <= uend) ? (int)(__ldg(tptr + u + 2) - e0) : run;
        int b3 = (u + 3 <= uend) ? (int)(__ldg(tptr + u + 3) - e0) : run;
        QAcc acc;
        acc.zero();
        auto flush = [&](int upto) {       // flush every unit that ends at or before stream position `upto`
            do {
                uint64_t* dst = part + u * q_part_words(LANES) + li;
                SCS_ASSERT(u >= 0 && (u + 1) * q_part_words(LANES) <= part_stride);
                dst[0] = acc.a;
                dst[LANES] = acc.b;
                if (li < 4) dst[2 * LANES] = acc.pp;
                acc.zero();
                ++u;
                b1 = b2; b2 = b3;
                b3 = (u + 3 <= uend) ? (int)(__ldg(tptr + u + 3) - e0) : run;
            } while (u < uend && b1 <= upto);
            if (u >= uend) b1 = 0x7fffffff;
        };
        if (b1 <= 0) flush(0);                           // empty units at the head of the run
        int pos = 0;
        // every lane holds four stream words (eight codes) of the current batch: one LDG.128, so that a warp's stream loads
        // touch each 32-byte sector once per 32 codes of a group.  (Staging the words in shared memory instead of shuffling
        // them measured no better: a 16-byte broadcast read still costs one wavefront per quarter-warp.)
        const uint4 zero4 = make_uint4(zero_word, zero_word, zero_word, zero_word);
        uint4 word = (8 * li < run) ? __ldcs(reinterpret_cast<const uint4*>(words) + li) : zero4;
        while (__any_sync(0xffffffffu, u < uend)) {
            const uint4 mine = word;
            {   // this lane's words of the next batch (positions beyond the run read as padding; runs are multiples of 8 codes)
                const int nw = (pos + BE) / 8 + li;
                word = (8 * nw < run) ? __ldcs(reinterpret_cast<const uint4*>(words) + nw) : zero4;
            }
            constexpr int UW = 4;                        // stream words (code pairs) whose rows are requested before the adds
#pragma unroll
            for (int j0 = 0; j0 < 4 * LANES; j0 += UW) {  // (block j0 / 4 of the batch is exactly the four words of lane j0 / 4)
                uint4 v[2 * UW];
                const uint32_t mw[4] = {mine.x, mine.y, mine.z, mine.w};
#pragma unroll
                for (int jj = 0; jj < UW; ++jj) {
                    const uint32_t w = __shfl_sync(0xffffffffu, mw[jj], g * LANES + j0 / 4);
                    SCS_ASSERT(((w & QC::MASK) >> QC::SA) <= (uint32_t)TR + 1 && (((w >> 14) & QC::MASK) >> QC::SA) <= (uint32_t)TR + 1);
                    v[2 * jj] = lds_u32x4(smem_base + ((w & QC::MASK) | lane_off));
                    v[2 * jj + 1] = lds_u32x4(smem_base + (((w >> 14) & QC::MASK) | lane_off));
                }
#pragma unroll
                for (int jj = 0; jj < 2 * UW; ++jj) acc.add(v[jj]);
                // units are padded to multiples of Q_UNIT_PAD = 2 * UW codes, so they can only end here
                const int upto = pos + 2 * (j0 + UW);
                if (b1 <= upto) flush(upto);
            }
            pos += BE;
        }
    }
}

// L[c, :] = -(sum over tiles of the unit sums) * 2^-F + the row's entries of count > Q_REP_MAX in fp64.  Four threads per
// (cell, lane) pair, each taking every fourth tile (the integer adds are order-free), folded by shuffles; the few fp64 terms
// are added in stream order.  Warp lane = s * 8 + q: q picks one of 8 consecutive pairs, s the tile class; the word-major
// layout of the unit sums makes every load of 8 pairs one or two fully used 32-byte sectors per class.  (Measured at a
// config-4 shard, 270 MB of unit sums from HBM: 8 classes per warp 124 us, a block of 8 warps with one class each 96 us.)
constexpr int QC_SUB = 4;
template <int LANES>
__global__ void __launch_bounds__(256) k_combine_q(int64_t nrows, int K, int KP, int ntiles, const uint64_t* __restrict__ part,
                                                   int64_t part_stride, double scale_inv, const int64_t* __restrict__ hptr,
                                                   const uint32_t* __restrict__ hcode, const int32_t* __restrict__ hcnt,
                                                   const double* __restrict__ T, int64_t t_stride, double* __restrict__ L,
                                                   int64_t l_stride, const int32_t* __restrict__ done, const int32_t* __restrict__ qbad) {
    constexpr int PW = q_part_words(LANES);
    const int r = blockIdx.y;
    if (done && done[r]) return;
    if (qbad[r]) return;
    const int lane = threadIdx.x & 31, sub = lane >> 3;
    const int64_t i = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + (lane & 7);   // (cell, lane) pair
    const bool live = i < nrows * LANES;
    const int64_t c = live ? i / LANES : 0;
    const int li = (int)(i - (i / LANES) * LANES);
    part += (int64_t)r * part_stride;
    T += (int64_t)r * t_stride;
    uint64_t sa = 0, sb = 0, pa = 0, pb = 0;
    if (live) {
        const uint64_t* p = part + c * PW + li;
        const int64_t tstride = nrows * PW;
        const bool has_pp = li < 4;                      // (lanes 4..7 of a 128-byte row carry no byte pair and store none)
        int t = sub;
        constexpr int CB = 3;
        for (; t + (CB - 1) * QC_SUB < ntiles; t += CB * QC_SUB) {
            uint64_t x[CB][3];
#pragma unroll
            for (int q = 0; q < CB; ++q) {
                x[q][0] = __ldcs(p + (int64_t)(t + q * QC_SUB) * tstride);
                x[q][1] = __ldcs(p + (int64_t)(t + q * QC_SUB) * tstride + LANES);
                x[q][2] = has_pp ? __ldcs(p + (int64_t)(t + q * QC_SUB) * tstride + 2 * LANES) : 0ull;
            }
#pragma unroll
            for (int q = 0; q < CB; ++q) { sa += x[q][0]; sb += x[q][1]; pa += x[q][2] & 0xffffffull; pb += x[q][2] >> 24; }
        }
        for (; t < ntiles; t += QC_SUB) {
            const uint64_t x2 = has_pp ? __ldcs(p + (int64_t)t * tstride + 2 * LANES) : 0ull;
            sa += __ldcs(p + (int64_t)t * tstride);
            sb += __ldcs(p + (int64_t)t * tstride + LANES);
            pa += x2 & 0xffffffull;
            pb += x2 >> 24;
        }
    }
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {                   // fold the four tile classes
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
        pa += __shfl_xor_sync(0xffffffffu, pa, o);
        pb += __shfl_xor_sync(0xffffffffu, pb, o);
    }
    sa -= pa << 56;                                      // the byte pair rode on top of q0 / q1 (mod 2^64; the true sums are < 2^63)
    sb -= pb << 56;
    uint64_t sx = li < 4 ? (pa << (16 * li)) + (pb << (16 * li + 8)) : 0ull;
#pragma unroll
    for (int o = 2; o; o >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o);   // lanes 0..3 of a row are neighbours in the warp
    if (!live || sub) return;
    auto val = [&](uint64_t s, int k) { return (k < K && s) ? -((double)s * scale_inv) : 0.0; };
    double y0 = val(sa, 2 * li), y1 = val(sb, 2 * li + 1), yx = val(sx, 2 * LANES);
    for (int64_t h = hptr[c]; h < hptr[c + 1]; ++h) {    // counts too large to be written as repeated codes
        const int cnt = __ldg(hcnt + h);
        if (cnt <= Q_REP_MAX) continue;
        const double* row = T + (int64_t)(__ldg(hcode + h) >> 1) * KP;
        if (2 * li + 1 < KP) { y0 += (double)cnt * row[2 * li]; y1 += (double)cnt * row[2 * li + 1]; }
        if (li == 0 && 2 * LANES < KP) yx += (double)cnt * row[2 * LANES];
    }
    double* y = L + (int64_t)r * l_stride + c * KP;
    if (2 * li + 1 < KP) *reinterpret_cast<double2*>(y + 2 * li) = make_double2(y0, y1);
    if (li == 0 && 2 * LANES + 1 < KP) *reinterpret_cast<double2*>(y + 2 * LANES) = make_double2(yx, 0.0);
}

}  // namespace scs
