// EM kernels of the scSplit hot path (sm_100a).  See DESIGN.md for the roofline of each.
//
//   k_spmm_rows      v1 SpMM (warp per row, gathers from global); the production SpMM is k_spmm_tiled (spmm_tiled.cuh)
//                    Y[row,:] = sum_{count-1 entries} X[g(code),:] + sum_{count>1 entries} cnt * X[g(code),:]
//                    E-step:  rows = cells,          X = log table T[2V][KP],  Y = L[N][KP]     (scSplit:175-178)
//                    M-step:  rows = 2V pseudo-rows, X = W[N][KP],             Y = S[2V][KP]    (scSplit:196)
//   k_bayes          W, per-cell LL, packed sparse posterior word, fixed-order block partials    (scSplit:181-188)
//   k_colsum_partial the same partials for an injected W (step-level M-step entry)
//   k_reduce_stats   block partials -> colsum[K], LL, dense-cell count                           (scSplit:188,198)
//   k_finalize       S, k_alt -> P, log2 P, log2(1-P)                                             (scSplit:196, 176-177)
//   update_prior     colsum -> lP_s; LL history; bitwise stop rule (last block of k_finalize)      (scSplit:198, 156-161)
//   (k_mstep_sparse, the M-step for saturated posteriors, lives in mstep_sparse.cuh)
#pragma once
#include "common.cuh"

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
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    X += (int64_t)r * x_stride;
    Y += (int64_t)r * y_stride;
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
template <int KP>
__device__ __forceinline__ uint32_t sparsify_row(const double (&w)[KP], int K) {
    int im = 0;
    double wm = w[0];
#pragma unroll
    for (int k = 1; k < KP; ++k)
        if (k < K && w[k] > wm) { wm = w[k]; im = k; }
    // the reference's own rounding puts the winning posterior at 1 +- O(1e-13), on either side of 1; NaN rows are never sparse
    bool ok = (wm == wm) && wm >= 0.5 && wm <= 1.5;
#pragma unroll
    for (int k = 0; k < KP; ++k)
        if (k < K && k != im) ok = ok && (w[k] < SPARSE_TINY);
    const double eps = 1.0 - wm;                      // exact (Sterbenz)
    const float ef = (float)eps;
    const uint32_t bits = __float_as_uint(ef);
    ok = ok && ((double)ef == eps) && ((bits & 31u) == 0u);
    return ok ? (bits | (uint32_t)im) : SPARSE_DENSE;
}

// W[c,i] = 1 / sum_j 2^(((L_j + s_j) - L_i) - s_i)   (scSplit:181-185, j ascending, left-to-right)
// ll_c   = log2( sum_k 2^((L_k - max_k L) + s_k) ) + max_k L   with pandas skipna  (scSplit:188)
// partial[r][blk][0..KP) = sum over the block's cells of W[c,k] (NaN skipped, scSplit:198), [KP] = sum of ll_c,
// [KP+1] = number of cells whose posterior row does not pack into a sparse word.
//
// One cell is handled by a group of BAYES_GL lanes (lane i = state i), so the K^2 exp2 of a cell run K-wide and a 20k-cell
// pool still gives every SM sub-partition several warps; sums over states are taken in the reference's order by shuffling.
// 2^x for the Bayes ratios.  Saturated posteriors make almost every exponent either hugely negative or hugely positive;
// exp2 returns exactly 0 below -1075 and exactly +inf from 1024 on, so those lanes skip the (long) double-precision exp2
// sequence without changing a bit of the result.  NaN exponents fall through to exp2 and stay NaN.
// The winning state's own term has an exponent of rounding-noise size (|x| ~ 1e-13): there 2^x = 1 + y + y^2/2 with
// y = x ln 2 is exact to well below half an ulp (the next term is < 2e-20 for |x| < 2^-20).
__device__ __forceinline__ double exp2_sat(double x) {
    if (x < -1100.0) return 0.0;
    if (x > 1100.0) return __longlong_as_double(0x7ff0000000000000LL);
    if (fabs(x) < 9.5367431640625e-07) {
        const double y = x * 0.6931471805599453094;
        return fma(y, fma(y, 0.5, 1.0), 1.0);
    }
    return exp2(x);
}

template <int KP>
struct BayesCfg {
    static constexpr int GL = KP <= 16 ? 16 : 32;     // lanes per cell
    static constexpr int THREADS = 1024;
    static constexpr int CPB = THREADS / GL;          // cells per block
};
inline int bayes_cells_per_block(int KP) { return KP <= 16 ? 64 : 32; }

// Block b of nblk owns the contiguous cells [N*b/nblk, N*(b+1)/nblk) and walks them CPB at a time (whole warps without
// cells skip the pass), so every block does the same work whatever N is and the grid can be sized to the SM count.
__device__ __forceinline__ int64_t bayes_cell_lo(int64_t N, int b, int nblk) { return N * b / nblk; }

template <int KP>
__global__ void __launch_bounds__(1024) k_bayes(int64_t N, int K, const double* __restrict__ L, double* __restrict__ W,
                                               const double* __restrict__ lps, double* __restrict__ partial, int nblk,
                                               uint32_t* __restrict__ hq, int64_t h_stride, const int32_t* __restrict__ done) {
    using C = BayesCfg<KP>;
    const int r = blockIdx.y;
    if (done && done[r]) return;
    __shared__ double s_red[32 * (KP + 2)];
    const int i = threadIdx.x & (C::GL - 1);
    const double si = i < K ? lps[(int64_t)r * KP + i] : 0.0;
    const int64_t c_lo = bayes_cell_lo(N, blockIdx.x, nblk), c_hi = bayes_cell_lo(N, blockIdx.x + 1, nblk);
    double v[KP + 2];
#pragma unroll
    for (int k = 0; k <= KP + 1; ++k) v[k] = 0.0;
    for (int64_t c0 = c_lo + (threadIdx.x >> 5) * (32 / C::GL); c0 < c_hi; c0 += C::CPB) {   // warp-uniform
        const int64_t c = c0 + (threadIdx.x & 31) / C::GL;
        const bool live = c < c_hi && i < K;
        const double Li = live ? L[((int64_t)r * N + c) * KP + i] : 0.0;
        const double ti = Li + si;                                   // L_i + s_i
        double denom = 0.0;
        for (int j = 0; j < K; ++j) {
            const double tj = __shfl_sync(0xffffffffu, ti, j, C::GL);
            denom += exp2_sat((tj - Li) - si);
        }
        const double wi = live ? 1.0 / denom : 0.0;
        if (c < c_hi && i < KP) W[((int64_t)r * N + c) * KP + i] = wi;
        // total log-likelihood term (skipna max, skipna sum), states taken in ascending order
        double m = __shfl_sync(0xffffffffu, Li, 0, C::GL);
        for (int k = 1; k < K; ++k) m = fmax(m, __shfl_sync(0xffffffffu, Li, k, C::GL));   // fmax ignores NaN operands
        const double xi = exp2_sat((Li - m) + si);
        double sum = 0.0;
        for (int k = 0; k < K; ++k) {
            const double xk = __shfl_sync(0xffffffffu, xi, k, C::GL);
            if (xk == xk) sum += xk;
        }
        double w[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) w[k] = __shfl_sync(0xffffffffu, wi, k, C::GL);
        if (c < c_hi && i == 0) {
            const double ll = log2(sum) + m;
#pragma unroll
            for (int k = 0; k < KP; ++k) v[k] += (w[k] == w[k]) ? w[k] : 0.0;
            v[KP] += (ll == ll) ? ll : 0.0;
            const uint32_t q = sparsify_row<KP>(w, K);
            hq[(int64_t)r * h_stride + c] = q;
            v[KP + 1] += q == SPARSE_DENSE ? 1.0 : 0.0;      // number of cells that need the dense M-step path
        }
    }
    block_reduce_store<KP + 2, C::GL>(v, s_red, partial + ((int64_t)r * nblk + blockIdx.x) * (KP + 2));
}

// column mass and packed form of an arbitrary W (step-level M-step entry): same block mapping and order as k_bayes
template <int KP>
__global__ void __launch_bounds__(1024) k_colsum_partial(int64_t N, int K, const double* __restrict__ W, double* __restrict__ partial,
                                                        int nblk, uint32_t* __restrict__ hq, int64_t h_stride) {
    using C = BayesCfg<KP>;
    const int r = blockIdx.y;
    __shared__ double s_red[32 * (KP + 2)];
    const int i = threadIdx.x & (C::GL - 1);
    const int64_t c_lo = bayes_cell_lo(N, blockIdx.x, nblk), c_hi = bayes_cell_lo(N, blockIdx.x + 1, nblk);
    double v[KP + 2];
#pragma unroll
    for (int k = 0; k <= KP + 1; ++k) v[k] = 0.0;
    for (int64_t c0 = c_lo + (threadIdx.x >> 5) * (32 / C::GL); c0 < c_hi; c0 += C::CPB) {   // warp-uniform
        const int64_t c = c0 + (threadIdx.x & 31) / C::GL;
        const double wi = (c < c_hi && i < K) ? W[((int64_t)r * N + c) * KP + i] : 0.0;
        double w[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) w[k] = __shfl_sync(0xffffffffu, wi, k, C::GL);
        if (c < c_hi && i == 0) {
#pragma unroll
            for (int k = 0; k < KP; ++k) v[k] += (w[k] == w[k]) ? w[k] : 0.0;
            const uint32_t q = sparsify_row<KP>(w, K);
            hq[(int64_t)r * h_stride + c] = q;
            v[KP + 1] += q == SPARSE_DENSE ? 1.0 : 0.0;
        }
    }
    // [KP] (LL) is left untouched by the caller's reduce (with_ll = 0)
    block_reduce_store<KP + 2, C::GL>(v, s_red, partial + ((int64_t)r * nblk + blockIdx.x) * (KP + 2));
}

// partial[r][blk][NV] -> tail[r][0..NV): column mass [0,KP), LL [KP] (if with_ll), dense-path cell count [KP+1] (if NV = KP+2);
// one block of 256 threads per restart, fixed order.
static __global__ void __launch_bounds__(256) k_reduce_stats(int KP, int NV, const double* __restrict__ partial, int nblk,
                                                      double* __restrict__ stats, int64_t stats_stride, int64_t tail_off, int with_ll,
                                                      const int32_t* __restrict__ done) {
    const int r = blockIdx.x;
    if (done && done[r]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < NV; k += 8) {
        double s = 0.0;
        for (int b = lane; b < nblk; b += 32) s += partial[((int64_t)r * nblk + b) * NV + k];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && (k != KP || with_ll)) stats[(int64_t)r * stats_stride + tail_off + k] = s;
    }
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

// P = (altW + k_alt) / ((altW + refW) + 1);  T[v] = log2 P, T[V+v] = log2(1 - P)   (scSplit:196, 176-177)
__device__ __forceinline__ void finalize_rows(int64_t V, int K, int KP, const double* __restrict__ stats, int64_t stats_stride,
                                              const double* __restrict__ kalt, double* __restrict__ P, double* __restrict__ T, int r) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V * KP) return;   // (the caller's block still takes part in the block count)
    const int64_t v = i / KP;
    const int k = (int)(i - v * KP);
    const double* S = stats + (int64_t)r * stats_stride;
    double p = 0.0, lp = 0.0, lq = 0.0;
    if (k < K) {
        const double a = S[v * KP + k], rf = S[(V + v) * KP + k];
        p = (a + kalt[v]) / ((a + rf) + 1.0);
        lp = log2(p);
        lq = log2(1.0 - p);
    }
    P[((int64_t)r * V + v) * KP + k] = p;
    if (T) {
        T[((int64_t)r * 2 * V + v) * KP + k] = lp;
        T[((int64_t)r * 2 * V + V + v) * KP + k] = lq;
    }
}

// With `prior` (the EM loop) the LAST block of a restart to finish also runs update_prior (block counter in prior->count,
// left at zero again): by then every block has read the stop flag, so a restart that stops NOW still gets this
// iteration's P (scSplit:155-162: E, M, then the test); `done` skips the restarts that had stopped before this iteration.
static __global__ void __launch_bounds__(256) k_finalize(int64_t V, int K, int KP, const double* __restrict__ stats, int64_t stats_stride,
                                                  const double* __restrict__ kalt, double* __restrict__ P, double* __restrict__ T,
                                                  const int32_t* __restrict__ done, const PriorArgs prior) {
    const int r = blockIdx.y;
    if (done && done[r]) return;
    finalize_rows(V, K, KP, stats, stats_stride, kalt, P, T, r);
    if (prior.lps) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(prior.count + r, 1) == (int)gridDim.x - 1) {
            prior.count[r] = 0;
            update_prior(r, K, KP, stats, stats_stride, prior);
        }
    }
}


// log tables from a given P (EM start, scs_em_set(SCS_P), post-EM scores)
static __global__ void __launch_bounds__(256) k_tables_from_P(int64_t V, int K, int KP, const double* __restrict__ P, double* __restrict__ T) {
    const int r = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V * KP) return;
    const int64_t v = i / KP;
    const int k = (int)(i - v * KP);
    double lp = 0.0, lq = 0.0;
    if (k < K) {
        const double p = P[((int64_t)r * V + v) * KP + k];
        lp = log2(p);
        lq = log2(1.0 - p);
    }
    T[((int64_t)r * 2 * V + v) * KP + k] = lp;
    T[((int64_t)r * 2 * V + V + v) * KP + k] = lq;
}

}  // namespace scs
