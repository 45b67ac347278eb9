// EM kernels of the scSplit hot path (sm_100a).  See DESIGN.md for the roofline of each.
//
//   k_spmm_rows      v1 SpMM (warp per row, gathers from global; scs_set_variant 1); the production kernels are k_estep_q
//                    (E-step sums in exact fixed point, estep_q.cuh), k_mstep_sparse (mstep_sparse.cuh) and k_spmm_tiled (spmm_tiled.cuh)
//                    Y[row,:] = sum_{count-1 entries} X[g(code),:] + sum_{count>1 entries} cnt * X[g(code),:]
//                    E-step:  rows = cells,          X = log table T[2V][KP],  Y = L[N][KP]     (scSplit:175-178)
//                    M-step:  rows = 2V pseudo-rows, X = W[N][KP],             Y = S[2V][KP]    (scSplit:196)
//   k_bayes          W, per-cell LL, packed sparse posterior word, fixed-order block partials    (scSplit:181-188); also the fp64
//                    stand-in of the fixed-point E-step for a restart whose table does not fit the format
//   k_colsum_partial the same partials for an injected W (step-level M-step entry)
//   k_reduce_stats   block partials -> colsum[K], LL, dense-cell count (step-level path; the EM loop does it in k_bayes)  (scSplit:188,198)
//   k_finalize       S, k_alt -> P, log2 P, log2(1-P) in fp64 and as packed fixed point            (scSplit:196, 176-177)
//   update_prior     colsum -> lP_s; LL history; bitwise stop rule (last block of k_finalize)      (scSplit:198, 156-161)
//   (k_mstep_sparse, the M-step for saturated posteriors, lives in mstep_sparse.cuh)
#pragma once
#include "common.cuh"
#include "estep_q.cuh"

namespace scs {

// threads per block of the thread-per-cell kernel k_masked_colsum (post.cu): small blocks so that even 20k cells give >2
// CTAs per SM; the block count fixes the order of its two-level reduction
constexpr int CELL_BLOCK = 64;

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// ---------------------------------------------------------------------------------------------------------
// v1: one warp per row, lanes stride over the row's entries, gathers straight from global memory (L1/L2).
// ---------------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) k_spmm_rows(int64_t nrows, const int64_t* __restrict__ lptr,
                                                   const uint32_t* __restrict__ lcode, const int64_t* __restrict__ hptr,
                                                   const uint32_t* __restrict__ hcode, const int32_t* __restrict__ hcnt,
                                                   const double* __restrict__ X, int64_t x_stride, double* __restrict__ Y,
                                                   int64_t y_stride, const int32_t* __restrict__ done) {
    const int r = blockIdx.y;
    if (done && done[r]) return;
    const int lane = threadIdx.x & 31;
    X += (int64_t)r * x_stride;
    Y += (int64_t)r * y_stride;
    // (grid-stride over the rows)
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < nrows; row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        double acc[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = 0.0;

        const int64_t lb = lptr[row], le = lptr[row + 1];
        int64_t e = lb + lane;
        for (; e + 32 < le; e += 64) {  // two independent gathers in flight per lane
            const uint32_t c0 = __ldg(lcode + e), c1 = __ldg(lcode + e + 32);
            const double* x0 = X + (int64_t)(c0 >> 1) * KP;
            const double* x1 = X + (int64_t)(c1 >> 1) * KP;
            double2 t0[KP / 2], t1[KP / 2];
#pragma unroll
            for (int j = 0; j < KP / 2; ++j) { t0[j] = ldg2(x0 + 2 * j); t1[j] = ldg2(x1 + 2 * j); }
#pragma unroll
            for (int j = 0; j < KP / 2; ++j) {
                acc[2 * j] += t0[j].x; acc[2 * j + 1] += t0[j].y;
                acc[2 * j] += t1[j].x; acc[2 * j + 1] += t1[j].y;
            }
        }
        if (e < le) {
            const uint32_t c0 = __ldg(lcode + e);
            const double* x0 = X + (int64_t)(c0 >> 1) * KP;
#pragma unroll
            for (int j = 0; j < KP / 2; ++j) { double2 t = ldg2(x0 + 2 * j); acc[2 * j] += t.x; acc[2 * j + 1] += t.y; }
        }
        const int64_t hb = hptr[row], he = hptr[row + 1];
        for (int64_t h = hb + lane; h < he; h += 32) {
            const uint32_t c0 = __ldg(hcode + h);
            const double cnt = (double)__ldg(hcnt + h);
            const double* x0 = X + (int64_t)(c0 >> 1) * KP;
#pragma unroll
            for (int j = 0; j < KP / 2; ++j) {
                double2 t = ldg2(x0 + 2 * j);
                acc[2 * j] += cnt * t.x;       // contracted to one DFMA (count is an exact small integer)
                acc[2 * j + 1] += cnt * t.y;
            }
        }
#pragma unroll
        for (int k = 0; k < KP; ++k) {
#pragma unroll
            for (int o = 16; o; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        }
        double* y = Y + row * KP;
#pragma unroll
        for (int j = 0; j < KP / 2; ++j)
            if (lane == j) *reinterpret_cast<double2*>(y + 2 * j) = make_double2(acc[2 * j], acc[2 * j + 1]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fixed-order block reduction of NV values per thread (256 threads): shuffle tree inside each warp, then the
// 8 warp results are added in warp order.  Same bits for the same inputs, every run.
// ---------------------------------------------------------------------------------------------------------
template <int NV, int MINO = 1>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* smem /*[warps][NV]*/, double* out /*[NV]*/) {
    // MINO > 1: only lanes that are multiples of MINO carry values (the others hold zeros), so the tree stops early
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o >= MINO; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) smem[warp * NV + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += smem[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// Posterior rows saturate: after a few iterations almost every cell has one state with W = 1 - O(1e-13) and all others
// below 1e-40.  Such a row is packed into ONE 32-bit word: the state in the low 5 bits and eps = 1 - W as a float whose low
// 5 mantissa bits are zero.  W near 1 is a multiple of 2^-53 (2^-52 above 1), so eps = +-m * 2^-53 is exact in float whenever
// m < 2^19 (|eps| < 5.8e-11), and 1.0 - (double)eps gives W back bit for bit.  The entries dropped (< 1e-40) are invisible to
// the M-step, where the sums only ever enter (S + k_alt) and (S + 1) with k_alt >= 1e-10: 2^31 such terms stay under half
// an ulp.  Rows that do not fit the format (soft, NaN) are marked SPARSE_DENSE and take the dense path.
constexpr double SPARSE_TINY = 1e-40;
// Both special words decode to "state s, 1 - weight = 1.0f", i.e. weight 0: the sparse M-step can run them through its
// branch-free accumulate (adds 0.0) and only has to notice SPARSE_DENSE.  A packed row never looks like either
// (its 1 - weight lies in [-0.5, 0.5]).
constexpr uint32_t SPARSE_DENSE = 0x3f800000u;       // unsaturated row: take the full posterior row
constexpr uint32_t SPARSE_NOP = 0x3f800001u;         // no entry (a lane past the end of a pseudo-row)
// The packing is decided by the GL lanes of a cell together (lane i holds W[c,i]): the row is sparse iff exactly one
// live lane holds a weight in [0.5, 1.5], every other live lane is below SPARSE_TINY (NaN is neither) and the winner's
// 1 - weight fits the float format.  Every lane of the group returns the word.
template <int GL>
__device__ __forceinline__ uint32_t sparsify_group(double wi, int i, int K) {
    const int lane = threadIdx.x & 31, base = (lane / GL) * GL;        // (lanes beyond the last whole group get an empty mask)
    const uint32_t gmask = (GL == 32 ? 0xffffffffu : ((1u << GL) - 1u) << base);
    const uint32_t live = (K >= 32 ? 0xffffffffu : ((1u << K) - 1u) << base) & gmask;
    const bool big = wi >= 0.5 && wi <= 1.5;
    const uint32_t bigm = __ballot_sync(0xffffffffu, big && i < K) & gmask;
    const uint32_t tinym = __ballot_sync(0xffffffffu, wi < SPARSE_TINY && i < K) & gmask;
    const double eps = 1.0 - wi;                      // exact for the winner (Sterbenz)
    const float ef = (float)eps;
    const uint32_t bits = __float_as_uint(ef);
    const bool fits = ((double)ef == eps) && ((bits & 31u) == 0u);
    const int win = __ffs(bigm) - 1;                  // warp lane of the winner (if any)
    const uint32_t word = __shfl_sync(0xffffffffu, fits ? (bits | (uint32_t)i) : SPARSE_DENSE, win < 0 ? 0 : win);
    const bool ok = __popc(bigm) == 1 && ((bigm | tinym) == live);
    return ok ? word : SPARSE_DENSE;
}

// Block statistics of k_bayes / k_colsum_partial: lane i of every cell group has accumulated column i's mass in `mass`, lane
// 0 of the group the LL terms and the dense-row count.  Fixed order: the groups of a warp, then the warps.
template <int KP, int GL>
__device__ __forceinline__ void block_stats_store(double mass, double ll, double dense, double* smem /*[warps][KP+2]*/,
                                                  double* out /*[KP+2]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, i = lane % GL;
    const double mass0 = mass, ll0 = ll, dense0 = dense;
#pragma unroll
    for (int q = 1; q < 32 / GL; ++q) {              // the groups of the warp, in order, into its first group
        mass += __shfl_down_sync(0xffffffffu, mass0, q * GL);
        ll += __shfl_down_sync(0xffffffffu, ll0, q * GL);
        dense += __shfl_down_sync(0xffffffffu, dense0, q * GL);
    }
    if (lane < GL && i < KP) smem[warp * (KP + 2) + i] = mass;
    if (lane == 0) {
        smem[warp * (KP + 2) + KP] = ll;
        smem[warp * (KP + 2) + KP + 1] = dense;
    }
    __syncthreads();
    if (threadIdx.x < KP + 2) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += smem[w * (KP + 2) + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// partial[blk][NV] -> tail[0..NV): warp w takes values w, w + nwarps, ...; lanes stride over the blocks, then an xor tree.
// L2 loads (__ldcg): the partials were written by other blocks of the same launch when this runs inside k_bayes.
__device__ __forceinline__ void reduce_partials(int KP, int NV, const double* partial, int nblk, double* tail, int with_ll) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int k = warp; k < NV; k += nwarps) {
        double s = 0.0;
        for (int b = lane; b < nblk; b += 32) s += __ldcg(partial + (int64_t)b * NV + k);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && (k != KP || with_ll)) tail[k] = s;
    }
}

// W[c,i] = 1 / sum_j 2^(((L_j + s_j) - L_i) - s_i)   (scSplit:181-185, j ascending, left-to-right)
// ll_c   = log2( sum_k 2^((L_k - max_k L) + s_k) ) + max_k L   with pandas skipna  (scSplit:188)
// partial[r][blk][0..KP) = sum over the block's cells of W[c,k] (NaN skipped, scSplit:198), [KP] = sum of ll_c,
// [KP+1] = number of cells whose posterior row does not pack into a sparse word.
//
// One cell is handled by a group of BAYES_GL lanes (lane i = state i), so the K^2 exp2 of a cell run K-wide and a 20k-cell
// pool still gives every SM sub-partition several warps; sums over states are taken in the reference's order by shuffling.
// 2^x for the Bayes ratios, branch-free.  x = n + f with n = rint(x), |f| <= 1/2; 2^f by the degree-13 Taylor series in
// f ln 2 (remainder < 4.2e-18, below a quarter ulp; Horner rounding keeps the result within ~1 ulp, as CUDA's own exp2
// is); 2^n is applied as two exact power-of-two factors so that results in the denormal range are rounded once and
// exponents from 1024 on overflow to +inf.  The argument is clamped to +-1100 first, where the result is exactly 0 or
// +inf anyway (saturated posteriors put almost every exponent out there); NaN stays NaN.  The coefficients sit in the
// constant bank so the fused multiply-adds read them as operands.
__constant__ double EXP2_C[14] = {1.0,
                                  0x1.62e42fefa39efp-1,  0x1.ebfbdff82c58fp-3,  0x1.c6b08d704a0c0p-5,  0x1.3b2ab6fba4e77p-7,
                                  0x1.5d87fe78a6731p-10, 0x1.430912f86c787p-13, 0x1.ffcbfc588b0c7p-17, 0x1.62c0223a5c824p-20,
                                  0x1.b5253d395e7c4p-24, 0x1.e4cf5158b8ecap-28, 0x1.e8cac7351bb25p-32, 0x1.c3bd650fc2986p-36,
                                  0x1.816193166d0f9p-40};
__device__ __forceinline__ double exp2_bf(double x) {
    const double xc = fmin(fmax(x, -1100.0), 1100.0);
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52: the add rounds xc to an integer in the low word
    const double t = xc + MAGIC;
    const int n = __double2loint(t);
    const double f = xc - (t - MAGIC);
    double p = EXP2_C[13];
#pragma unroll
    for (int k = 12; k >= 0; --k) p = fma(p, f, EXP2_C[k]);
    const int n1 = n >> 1, n2 = n - n1;
    const double r = (p * __hiloint2double((n1 + 1023) << 20, 0)) * __hiloint2double((n2 + 1023) << 20, 0);
    return x != x ? x : r;
}

template <int KP>
struct BayesCfg {
    // lanes per cell.  The kernel is bound by the fp64 pipe (ten 2^x per lane and cell at K = 9; B200 issues 64 fp64 lanes per
    // clock and SM), and a warp instruction costs the same whatever lanes are idle: K = 9 in groups of 16 wasted 7 lanes of
    // 16, in groups of 10 (three cells per warp, lanes 30 and 31 idle) 2 of 32.
    static constexpr int GL = KP <= 8 ? 8 : KP <= 10 ? 10 : KP <= 16 ? 16 : 32;
    static constexpr int CPW = 32 / GL;               // cells per warp
    static constexpr int THREADS = 1024;
    static constexpr int CPB = (THREADS / 32) * CPW;  // cells per block and pass
};
inline int bayes_cells_per_block(int KP) { return 32 * (32 / (KP <= 8 ? 8 : KP <= 10 ? 10 : KP <= 16 ? 16 : 32)); }

// Block b of nblk owns the contiguous cells [N*b/nblk, N*(b+1)/nblk) and walks them CPB at a time (whole warps without
// cells skip the pass), so every block does the same work whatever N is and the grid can be sized to the SM count.
__device__ __forceinline__ int64_t bayes_cell_lo(int64_t N, int b, int nblk) { return N * b / nblk; }

// The E-step sums normally reach k_bayes through L (written by k_combine_q or the fp64 SpMM kernels).  With the fixed-point
// E-step a restart whose table did not fit the format (qbad[r], set by k_finalize) has no sums: k_bayes gathers that
// restart's rows from the fp64 table right here -- every lane walks the cell's entries for its own state -- and writes L for
// the readers of lP_c_s.  Rare (a dying cluster), so it is the slow simple loop, not a kernel of its own standing by.
struct LSource {
    const int32_t* qbad = nullptr;    // [R]; nullptr: L is always in memory
    const int64_t* lptr = nullptr;    // cell-major stream (count-1 / count>1 entries)
    const int64_t* hptr = nullptr;
    const uint32_t* lcode = nullptr;
    const uint32_t* hcode = nullptr;
    const int32_t* hcnt = nullptr;
    const double* T = nullptr;        // [R][2V][KP]
    int64_t t_stride = 0;
    // lock-step fixed-point E-step (estep_lk.cuh): the sums arrive as 64-bit integer totals, L = -total * 2^-F; k_bayes adds the
    // entries whose counts were too large for repeated codes (xheavy[c] of them in cell c) in fp64, writes L and zeroes the totals
    unsigned long long* Lq = nullptr; // [R][N][KP]
    const int32_t* xheavy = nullptr;  // [N]
    double scale_inv = 0.0;           // 2^-F
    int rep_max = 0;                  // counts up to this were written as repeated codes
};

template <int KP>
__device__ __forceinline__ double bayes_row_from_table(const LSource& s, int r, int64_t c, bool cell, int i, int K) {
    const double* T = s.T + (int64_t)r * s.t_stride;
    double acc = 0.0;
    if (cell && i < K) {
        const int64_t lb = s.lptr[c], le = s.lptr[c + 1];
        int64_t e = lb;
        for (; e + 2 <= le; e += 2) {
            const double x0 = __ldg(T + (int64_t)(__ldg(s.lcode + e) >> 1) * KP + i);
            const double x1 = __ldg(T + (int64_t)(__ldg(s.lcode + e + 1) >> 1) * KP + i);
            acc += x0;
            acc += x1;
        }
        for (; e < le; ++e) acc += __ldg(T + (int64_t)(__ldg(s.lcode + e) >> 1) * KP + i);
        for (int64_t h = s.hptr[c]; h < s.hptr[c + 1]; ++h)
            acc += (double)__ldg(s.hcnt + h) * __ldg(T + (int64_t)(__ldg(s.hcode + h) >> 1) * KP + i);
    }
    return acc;
}

template <int KP>
__global__ void __launch_bounds__(1024) k_bayes(int64_t N, int K, const LSource src, double* __restrict__ L, double* __restrict__ W,
                                               const double* __restrict__ lps, double* __restrict__ partial, int nblk,
                                               uint32_t* __restrict__ hq, int64_t h_stride, const int32_t* __restrict__ done,
                                               double* __restrict__ tail /* stats + tail_off */, int64_t stats_stride,
                                               int32_t* __restrict__ count, int32_t* __restrict__ sat, double sat_thr) {
    using C = BayesCfg<KP>;
    const int r = blockIdx.y;
    if (done && done[r]) return;
    __shared__ double s_red[32 * (KP + 2)];
    __shared__ int s_last;
    const int grp = (threadIdx.x & 31) / C::GL, i = (threadIdx.x & 31) - grp * C::GL, gbase = grp * C::GL;
    const double si = i < K ? lps[(int64_t)r * KP + i] : 0.0;
    const int64_t c_lo = bayes_cell_lo(N, blockIdx.x, nblk), c_hi = bayes_cell_lo(N, blockIdx.x + 1, nblk);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    double mass = 0.0, llsum = 0.0, dense = 0.0;
    for (int64_t c0 = c_lo + (threadIdx.x >> 5) * C::CPW; c0 < c_hi; c0 += C::CPB) {   // warp-uniform
        const int64_t c = c0 + grp;
        const bool cell = grp < C::CPW && c < c_hi, live = cell && i < K;       // (lanes beyond the last whole group idle)
        double Li;
        if (src.qbad && src.qbad[r]) {
            Li = bayes_row_from_table<KP>(src, r, c, cell, i, K);
            if (cell && i < KP) L[((int64_t)r * N + c) * KP + i] = Li;
        } else if (src.Lq) {
            Li = 0.0;
            if (cell && i < KP) {
                unsigned long long* q = src.Lq + ((int64_t)r * N + c) * KP + i;
                const unsigned long long sq = __ldcs(q);
                const int nx = __ldg(src.xheavy + c);                    // (both loads are in flight before the store below)
                *q = 0ull;                                               // the next E-step adds to zero again
                if (i < K && sq) Li = -((double)sq * src.scale_inv);
                if (nx && i < K) {                                       // (rare) counts above Q_REP_MAX, in stream order
                    const double* T = src.T + (int64_t)r * src.t_stride;
                    for (int64_t h = src.hptr[c]; h < src.hptr[c + 1]; ++h) {
                        const int cnt = __ldg(src.hcnt + h);
                        if (cnt > src.rep_max) Li += (double)cnt * __ldg(T + (int64_t)(__ldg(src.hcode + h) >> 1) * KP + i);
                    }
                }
                L[((int64_t)r * N + c) * KP + i] = Li;
            }
        } else {
            Li = live ? L[((int64_t)r * N + c) * KP + i] : 0.0;
        }
        const double ti = Li + si;                                   // L_i + s_i
        double denom = 0.0;
        // (unrolled over the padded state count: the K exponentials are independent and run side by side; only the adds are
        // a chain.  With the loop bound K at run time the kernel evaluated them one after the other.)
#pragma unroll
        for (int j = 0; j < KP; ++j) {                               // j ascending, left to right, as the reference adds
            const double tj = __shfl_sync(0xffffffffu, ti, min(gbase + j, 31));
            const double ej = exp2_bf((tj - Li) - si);
            if (j < K) denom += ej;
        }
        const double wi = live ? 1.0 / denom : 0.0;
        if (cell && i < KP) W[((int64_t)r * N + c) * KP + i] = wi;
        // total log-likelihood term: skipna max and skipna sum over the states.  Lanes beyond K carry NaN into the max
        // (fmax drops NaN operands, and an all-NaN row stays NaN) and 0 into the sum; every lane of the group gathers both.
        const double mi = live ? Li : qnan;
        double m = qnan;
#pragma unroll
        for (int j = 0; j < KP; ++j) m = fmax(m, __shfl_sync(0xffffffffu, mi, min(gbase + j, 31)));
        const double xi = exp2_bf((Li - m) + si);
        const double xs = (live && xi == xi) ? xi : 0.0;
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < KP; ++j) sum += __shfl_sync(0xffffffffu, xs, min(gbase + j, 31));     // states ascending
        const uint32_t q = sparsify_group<C::GL>(wi, i, K);
        if (live && wi == wi) mass += wi;                            // column mass, NaN skipped (scSplit:198)
        if (cell && i == 0) {
            const double ll = log2(sum) + m;
            if (ll == ll) llsum += ll;
            hq[(int64_t)r * h_stride + c] = q;
            dense += q == SPARSE_DENSE ? 1.0 : 0.0;                  // number of cells that need the dense M-step path
        }
    }
    block_stats_store<KP, C::GL>(mass, llsum, dense, s_red, partial + ((int64_t)r * nblk + blockIdx.x) * (KP + 2));
    // the last block of the restart to get here folds the block partials into the statistics tail (column mass, LL,
    // dense-row count), in the order k_reduce_stats uses; the block counter is left at zero again
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(count + r, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    reduce_partials(KP, KP + 2, partial + (int64_t)r * nblk * (KP + 2), nblk, tail + (int64_t)r * stats_stride, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        count[r] = 0;
        // latch: from the first iteration whose dense-row count is at or below the saturation threshold the sparse M-step
        // owns this restart for the rest of the run (Gate, common.cuh)
        if (sat && tail[(int64_t)r * stats_stride + KP + 1] <= sat_thr) sat[r] = 1;
    }
}

// column mass and packed form of an arbitrary W (step-level M-step entry): same block mapping and order as k_bayes
template <int KP>
__global__ void __launch_bounds__(1024) k_colsum_partial(int64_t N, int K, const double* __restrict__ W, double* __restrict__ partial,
                                                        int nblk, uint32_t* __restrict__ hq, int64_t h_stride) {
    using C = BayesCfg<KP>;
    const int r = blockIdx.y;
    __shared__ double s_red[32 * (KP + 2)];
    const int grp = (threadIdx.x & 31) / C::GL, i = (threadIdx.x & 31) - grp * C::GL;
    const int64_t c_lo = bayes_cell_lo(N, blockIdx.x, nblk), c_hi = bayes_cell_lo(N, blockIdx.x + 1, nblk);
    double mass = 0.0, dense = 0.0;
    for (int64_t c0 = c_lo + (threadIdx.x >> 5) * C::CPW; c0 < c_hi; c0 += C::CPB) {   // warp-uniform
        const int64_t c = c0 + grp;
        const bool cell = grp < C::CPW && c < c_hi, live = cell && i < K;
        const double wi = live ? W[((int64_t)r * N + c) * KP + i] : 0.0;
        const uint32_t q = sparsify_group<C::GL>(wi, i, K);
        if (live && wi == wi) mass += wi;
        if (cell && i == 0) {
            hq[(int64_t)r * h_stride + c] = q;
            dense += q == SPARSE_DENSE ? 1.0 : 0.0;
        }
    }
    // [KP] (LL) is left untouched by the caller's reduce (with_ll = 0)
    block_stats_store<KP, C::GL>(mass, 0.0, dense, s_red, partial + ((int64_t)r * nblk + blockIdx.x) * (KP + 2));
}

// partial[r][blk][NV] -> tail[r][0..NV): column mass [0,KP), LL [KP] (if with_ll), dense-path cell count [KP+1] (if NV = KP+2);
// one block of 256 threads per restart, fixed order.
static __global__ void __launch_bounds__(256) k_reduce_stats(int KP, int NV, const double* __restrict__ partial, int nblk,
                                                      double* __restrict__ stats, int64_t stats_stride, int64_t tail_off, int with_ll,
                                                      const int32_t* __restrict__ done) {
    const int r = blockIdx.x;
    if (done && done[r]) return;
    reduce_partials(KP, NV, partial + (int64_t)r * nblk * NV, nblk, stats + (int64_t)r * stats_stride + tail_off, with_ll);
}

// numpy's pairwise add.reduce for n < 128 contiguous doubles (what pandas Series.sum() runs for K values)
__device__ __forceinline__ double numpy_sum_small(const double* a, int n) {
    if (n < 8) {
        double s = 0.0;   // numpy starts from a[0]; 0.0 + a[0] is exact, and colsums are never -0.0
        for (int i = 0; i < n; ++i) s += a[i];
        return s;
    }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 += a[i]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
        r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += a[i];
    return res;
}

// What one thread does per restart once the statistics tail is final.
struct PriorArgs {
    int64_t tail_off;     // offset of [colsum KP | LL | dense count] inside a restart's statistics
    double* lps;          // [R][KP]
    double* hist;         // [R][max_iter]
    int max_iter;
    double* ll_prev;      // [R]
    int32_t* iters;       // [R]
    int32_t* done;        // [R]
    int32_t* count;       // [R] blocks of k_finalize that have finished (self-resetting)
    int track, check_conv;
};

// lP_s = log2(colsum) - log2(sum colsum)  (scSplit:198); LL history and the bitwise stop rule (scSplit:156-162)
__device__ __forceinline__ void update_prior(int r, int K, int KP, const double* __restrict__ stats, int64_t stats_stride, const PriorArgs& a) {
    if (a.track && a.check_conv && a.done[r]) return;
    const double* tail = stats + (int64_t)r * stats_stride + a.tail_off;
    const double tot = numpy_sum_small(tail, K);
    const double ltot = log2(tot);
    for (int k = 0; k < K; ++k) a.lps[(int64_t)r * KP + k] = log2(tail[k]) - ltot;
    if (a.track) {
        const double ll = tail[KP];
        const int it = a.iters[r];
        if (it < a.max_iter) a.hist[(int64_t)r * a.max_iter + it] = ll;
        a.iters[r] = it + 1;
        if (a.check_conv && (ll == a.ll_prev[r] || it + 1 >= a.max_iter)) a.done[r] = 1;
        a.ll_prev[r] = ll;
    }
}

// What k_finalize writes besides P: the fp64 log table T and (when the E-step runs in fixed point, estep_q.cuh) its
// quantised copy Tq with the per-restart "does not fit" flag.
struct TableOut {
    double* T = nullptr;          // [R][2V][KP]
    uint64_t* Tq = nullptr;       // [R][2V][2 * qlanes]; split: [2][R][2V][8], table h at Tq + h * split_stride
    int qlanes = 0;
    int split = 0;                // 16-value-wide table kept as two tables of 64-byte rows (EmState::q_split)
    int64_t split_stride = 0;
    double qscale = 0.0;          // 2^F
    int32_t* qbad = nullptr;      // [R] the table of restart r does not fit the format
    int32_t* qbad_next = nullptr; // [R] staging of the flag while the blocks of an EM-loop launch are still writing the table
};

constexpr int FIN_THREADS = 256, FIN_RPT = 2;           // threads per block; SNVs per thread (two independent log2 chains in flight)
inline int finalize_rows_per_block(int KP) { return FIN_RPT * (FIN_THREADS / KP); }

// P = (altW + k_alt) / ((altW + refW) + 1);  T[v] = log2 P, T[V+v] = log2(1 - P)   (scSplit:196, 176-177)
// A block owns FIN_RPT * (FIN_THREADS / KP) whole SNVs (thread = one state of FIN_RPT SNVs), so that the packed fixed-point
// rows -- which mix the states of a row into 16-byte lanes -- can be put together from shared memory, and so that the grid
// of a 30k-SNV table is half a wave instead of a wave and a sliver.  With Pin the probabilities are taken as
// given (EM start, scs_em_set(SCS_P), post-EM scores) instead of computed from the statistics.
// With `prior` (the EM loop) the LAST block of a restart to finish also runs update_prior (block counter in prior.count,
// left at zero again): by then every block has read the stop flag, so a restart that stops NOW still gets this
// iteration's P (scSplit:155-162: E, M, then the test); `done` skips the restarts that had stopped before this iteration.
static __global__ void __launch_bounds__(FIN_THREADS) k_finalize(int64_t V, int K, int KP, const double* __restrict__ stats, int64_t stats_stride,
                                                          const double* __restrict__ kalt, const double* __restrict__ Pin,
                                                          double* __restrict__ P, const TableOut out,
                                                          const int32_t* __restrict__ done, const PriorArgs prior) {
    __shared__ double s_l[2][FIN_RPT * FIN_THREADS];
    const int r = blockIdx.y;
    if (done && done[r]) return;
    const int rph = FIN_THREADS / KP, rpb = FIN_RPT * rph;             // SNVs per pass of the threads / per block
    const int64_t v0 = (int64_t)blockIdx.x * rpb;
    const int rl = threadIdx.x / KP, k = threadIdx.x - rl * KP;
    double lp[FIN_RPT], lq[FIN_RPT], p[FIN_RPT];
#pragma unroll
    for (int h = 0; h < FIN_RPT; ++h) {
        const int64_t v = v0 + h * rph + rl;
        p[h] = 0.0;
        if (rl < rph && v < V && k < K) {
            if (Pin) {
                p[h] = Pin[((int64_t)r * V + v) * KP + k];
            } else {
                const double* S = stats + (int64_t)r * stats_stride;
                const double a = S[v * KP + k], rf = S[(V + v) * KP + k];
                p[h] = (a + kalt[v]) / ((a + rf) + 1.0);
            }
        }
    }
#pragma unroll
    for (int h = 0; h < FIN_RPT; ++h) {
        const int64_t v = v0 + h * rph + rl;
        const bool live = rl < rph && v < V;
        lp[h] = (live && k < K) ? log2(p[h]) : 0.0;
        lq[h] = (live && k < K) ? log2(1.0 - p[h]) : 0.0;
        if (live) {
            if (P) P[((int64_t)r * V + v) * KP + k] = p[h];
            if (out.T) {
                out.T[((int64_t)r * 2 * V + v) * KP + k] = lp[h];
                out.T[((int64_t)r * 2 * V + V + v) * KP + k] = lq[h];
            }
        }
    }
    if (out.Tq) {                                        // (uniform over the grid)
        if (rl < rph) {
#pragma unroll
            for (int h = 0; h < FIN_RPT; ++h) { s_l[0][(h * rph + rl) * KP + k] = lp[h]; s_l[1][(h * rph + rl) * KP + k] = lq[h]; }
        }
        __syncthreads();
        const int QL = out.qlanes;
        bool bad = false;
        for (int o = threadIdx.x; o < rpb * 2 * QL; o += FIN_THREADS) {
            const int row = o / (2 * QL), rem = o - row * 2 * QL, sel = rem / QL, j = rem - sel * QL;
            if (v0 + row >= V) continue;
            const double* l = s_l[sel] + row * KP;
            if (out.split) {                             // lanes 0..3: values 0..7 and the bytes of value 8; lanes 4..7: values 9..16
                const int h = j >> 2, jj = j & 3, ka = 9 * h + 2 * jj, kb = ka + 1;
                const uint64_t a = ka < K ? q_quant(l[ka], out.qscale, &bad) : 0ull;
                const uint64_t b = kb < K ? q_quant(l[kb], out.qscale, &bad) : 0ull;
                const uint64_t x = (h == 0 && 8 < K) ? q_quant(l[8], out.qscale, &bad) : 0ull;
                uint4* dst = reinterpret_cast<uint4*>(out.Tq + h * out.split_stride + (((int64_t)r * 2 + sel) * V + v0 + row) * 8) + jj;
                *dst = q_pack_lane(a, b, x, jj);
                continue;
            }
            const int ka = 2 * j, kb = 2 * j + 1, kx = 2 * QL;
            const uint64_t a = ka < K ? q_quant(l[ka], out.qscale, &bad) : 0ull;
            const uint64_t b = kb < K ? q_quant(l[kb], out.qscale, &bad) : 0ull;
            const uint64_t x = (kx < K && j < 4) ? q_quant(l[kx], out.qscale, &bad) : 0ull;
            uint4* dst = reinterpret_cast<uint4*>(out.Tq + (((int64_t)r * 2 + sel) * V + v0 + row) * (2 * QL)) + j;
            *dst = q_pack_lane(a, b, x, j);
        }
        if (bad) (prior.lps ? out.qbad_next : out.qbad)[r] = 1;
    }
    if (prior.lps) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(prior.count + r, 1) == (int)gridDim.x - 1) {
            prior.count[r] = 0;
            if (out.Tq) {                                // every block has written its part of the table: publish its flag
                out.qbad[r] = *reinterpret_cast<volatile int32_t*>(out.qbad_next + r);
                out.qbad_next[r] = 0;
            }
            update_prior(r, K, KP, stats, stats_stride, prior);
        }
    }
}

}  // namespace scs
