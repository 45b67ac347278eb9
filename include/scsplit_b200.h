/* scsplit_b200 -- C-ABI of the B200-native `scSplit run` EM hot path.
 *
 * The reference (jon-xu/scSplit, one Python script) has no FFI: its hot path is the
 * `models` class (scSplit:64-336) driven by `run.core` (scSplit:470-490).  This header
 * declares what a maintainer's binding for that path would call; every entry point cites
 * the reference lines it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; scs_last_error() returns the
 *     message of the last failure on the calling thread.
 *   - all pointers are HOST pointers to caller-allocated, C-contiguous buffers unless a
 *     parameter is named *_dev.  The context owns all device memory.
 *   - a context is bound to one CUDA device and is not thread-safe.
 *   - there is no CPU fallback: without a CUDA device scs_create fails.
 *   - matrices are row-major: P[V][K], W[N][K], L[N][K]; restart-batched arrays put the
 *     restart index first: P[R][V][K].
 *   - K counts model states including the doublet state (scSplit:535-538), 1 <= K <= SCS_MAX_K.
 */
#ifndef SCSPLIT_B200_H
#define SCSPLIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define SCS_ABI_VERSION 1
#define SCS_MAX_K 32
#define SCS_MAX_ITER 300 /* scSplit:156 */

typedef struct scs_ctx scs_ctx;

/* which array scs_em_get / scs_em_set moves */
enum scs_array {
    SCS_P = 0,      /* model_af            double[V][K]   scSplit:97,196  */
    SCS_W = 1,      /* P_s_c               double[N][K]   scSplit:89,185  */
    SCS_L = 2,      /* lP_c_s              double[N][K]   scSplit:90,178  */
    SCS_LPS = 3,    /* lP_s                double[K]      scSplit:91,198  */
    SCS_LLHIST = 4, /* sum_log_likelihood[2:]  double[max_iter]  scSplit:161 */
    SCS_STATS = 5   /* packed M-step sufficient statistics double[2V*KP+KP+2] */
};

int scs_abi_version(void);
const char* scs_last_error(void);
int scs_device_count(int* count);

/* Upload the count matrices once and build both device orientations.
 * Replaces scSplit:546-547 (csr_matrix(ref.values), csr_matrix(alt.values)) and the per-iteration
 * transposes/sparse adds of scSplit:176-178,196.  Inputs are the reference's own base_calls_mtx:
 * two SNV-major CSR matrices V x N with sorted, duplicate-free column indices and strictly positive
 * counts.  `data_bytes` is the element size of the count arrays (2, 4 or 8: int16/int32/int64).
 * Counts above INT32_MAX are an error, never wrapped. */
int scs_create(scs_ctx** out, int device, int64_t V, int64_t N,
               const int64_t* ref_indptr, const int32_t* ref_indices, const void* ref_data,
               const int64_t* alt_indptr, const int32_t* alt_indices, const void* alt_data,
               int data_bytes);
int scs_destroy(scs_ctx* ctx);

/* sizes of the device layout: out[0]=nnz(ref) out[1]=nnz(alt) out[2]=nnz(union pattern)
 * out[3]=count-1 ("light") entries out[4]=count>1 ("heavy") entries out[5]=device bytes held */
int scs_layout_info(scs_ctx* ctx, int64_t out[6]);

/* k_alt[v] = (sum_c A + 1) / ((sum_c R + 1) + (sum_c A + 1)); scSplit:100-103. out: double[V] */
int scs_kalt(scs_ctx* ctx, double* out);

/* Hard one-hot M-step that seeds P for R restarts; scSplit:137-145.
 * labels: int32[R][N], -1 = cell not in the initial subset, else cluster in [0,K). out: double[R][V][K] */
int scs_seed_af(scs_ctx* ctx, int R, int K, const int32_t* labels, double* P0_out);

/* Non-zero pattern counts of (A+R) restricted to kept rows/columns, for the initialiser's
 * coverage-trimming loop; scSplit:117-118,124-125.  keep_*: uint8 masks (NULL = keep all).
 * row_cnt: int64[V] (valid for kept rows), col_cnt: int64[N] (valid for kept columns). */
int scs_pattern_counts(scs_ctx* ctx, const uint8_t* keep_rows, const uint8_t* keep_cols,
                       int64_t* row_cnt, int64_t* col_cnt);

/* ---- EM over R batched restarts (run_EM, scSplit:148-162; restart driver scSplit:470-484) ---- */

/* Device memory one batched restart of K states needs (out[0], bytes: tables, posteriors, statistics, per-unit partial sums
 * of both SpMM orientations, seeding temporaries) and what the device can still give (out[1], bytes: free memory plus the
 * unused part of the context's pool).  The restart driver sizes its batches with it (scSplit:470-490 runs restarts serially). */
int scs_em_footprint(scs_ctx* ctx, int K, int64_t out[2]);
/* Allocate EM state for R restarts of K states, upload P0 (double[R][V][K]); lP_s = log2(1/K), scSplit:91. */
int scs_em_begin(scs_ctx* ctx, int R, int K, const double* P0, int max_iter);
/* Run up to n_iter more iterations on every unconverged restart. check_conv=1 applies the reference's
 * stop rule (bitwise-equal consecutive LL, scSplit:156) on the device; 0 runs exactly n_iter (benchmarks). */
int scs_em_iterate(scs_ctx* ctx, int n_iter, int check_conv);
/* Capture the launches of one iteration as CUDA graphs now (both M-step launch sets of the given check_conv), so that the next
 * scs_em_iterate replays them from its first iteration; executes nothing.  Needs at least one iteration to have run since
 * scs_em_begin (the first pass does the lazy builds a capture cannot hold); otherwise, and on a multi-rank context, a no-op.
 * scs_em_iterate captures by itself after 8 plain iterations; this is for callers that time a short loop (bench.py). */
int scs_em_capture(scs_ctx* ctx, int check_conv);
/* Device time in milliseconds (CUDA events on the context's stream) of the launches of the last scs_em_iterate. */
int scs_em_last_ms(scs_ctx* ctx, double* ms);
/* Run to the reference's stopping rule (cap max_iter). */
int scs_em_run(scs_ctx* ctx);
/* One E-step (calculate_cell_likelihood, scSplit:165-188) / one M-step (calculate_model_af,
 * scSplit:191-198) on every restart, regardless of convergence state.  For step-level parity tests. */
int scs_estep(scs_ctx* ctx);
int scs_mstep(scs_ctx* ctx);
/* iterations done, converged flag (iterations < max_iter, scSplit:162) and last LL (lP_c_m) per restart */
int scs_em_status(scs_ctx* ctx, int32_t* iters, int32_t* converged, double* ll);
/* how the E-step sums of the current EM state run: out[0] = lanes per fixed-point table row (4 or 8; 0 = the fp64 kernels do
 * the E-step), out[1] = fraction bits F of the fixed-point tables, out[2] = table tiles, out[3] = restarts whose table
 * currently does not fit the fixed-point format (their E-step runs in fp64), out[4] = form of the fixed-point E-step (0 = none,
 * 1 = a group of lanes per entry + combine kernel, 2 = lock-step: a lane per entry, integer totals by RED, 3 = lock-step in two passes over two tables of 64-byte rows), out[5] = warp
 * tasks of the lock-step stream.  scSplit:175-178; DESIGN.md section 4. */
int scs_em_info(scs_ctx* ctx, int64_t out[6]);
int scs_em_get(scs_ctx* ctx, int restart, int what, double* out);
int scs_em_set(scs_ctx* ctx, int restart, int what, const double* in);
int scs_em_end(scs_ctx* ctx);

/* ---- post-EM reductions ---- */

/* define_doublet, scSplit:210-225: out[i] = sum over cells with mask[c] != 0 of weight[c] * L_i(c), where L uses
 * the given (post-M) P.  mask: uint8[N] = number of clusters holding the cell (0/1 in practice). out: double[K] */
int scs_doublet_scores(scs_ctx* ctx, int K, const double* P, const uint8_t* mask, double* out);
/* refine_doublets, scSplit:243-244: out[c] = sum_v A[v,c] * (1 - P[v, labels[c]]), 0 where labels[c] < 0. */
int scs_refine_scores(scs_ctx* ctx, int K, const double* P, const int32_t* labels, double* out);
/* distinguishing_alleles reductions, scSplit:269-282: hard per-cluster sums and the ternary code
 * (+1 ALT present, -1 REF only, 0 unknown).  labels: int32[N] (-1 = unassigned).
 * n_ref, n_alt: int64[V][K]; pa: int8[V][K] (any of the three outputs may be NULL). */
int scs_cluster_counts(scs_ctx* ctx, int K, const int32_t* labels, int64_t* n_ref, int64_t* n_alt, int8_t* pa);

/* ---- multi-GPU (cells sharded across ranks) ---- */

/* 128-byte NCCL unique id, created on rank 0 and broadcast by the host (torch.distributed). */
int scs_comm_unique_id(uint8_t id[128]);
/* Join the communicator; afterwards k_alt, seed/EM statistics, doublet scores and cluster counts are
 * all-reduced (sum) over ranks inside the calls above.  Every rank must make the same calls. */
int scs_comm_init(scs_ctx* ctx, const uint8_t id[128], int rank, int world);

/* ---- input (host only): the reference's dense count CSV straight to CSR ----
 * Replaces pd.read_csv(header=0, index_col=0) + csr_matrix(frame.values), scSplit:544-546, without the dense V x N
 * int64 intermediate.  Plain non-negative integer fields only; returns 2 for anything else (quotes, floats, blanks) so
 * that the caller can fall back to a general CSV reader.  ids / barcodes are '\n'-joined byte strings. */
typedef struct scs_csv scs_csv;
int scs_csv_parse(const char* path, int threads, scs_csv** out);
/* out[0]=V out[1]=N out[2]=nnz out[3]=bytes of ids out[4]=bytes of barcodes */
int scs_csv_shape(scs_csv* csv, int64_t out[5]);
int scs_csv_fetch(scs_csv* csv, int64_t* indptr, int32_t* indices, int32_t* data, char* ids, char* barcodes);
int scs_csv_free(scs_csv* csv);

/* ---- output (host only): a float64 matrix in the text DataFrame.to_csv gives it ----
 * Replaces model.P_s_c.to_csv(path) (scSplit:596; format 0: every value as the shortest string that reads back as the
 * same double, laid out like Python's repr / numpy's str: "0.5", "1.0", "1e-05", "3.2e+17") and
 * model.dist_matrix / model.pa_matrix .to_csv(path, float_format='%.0f') (scSplit:599-600; format 1).  NaN is an empty
 * field (pandas' na_rep), infinities are "inf" / "-inf".  header: the first line without its newline; labels: the row
 * labels, each followed by '\n'; values: row-major [rows][cols].  The caller passes only labels that need no quoting. */
int scs_csv_write_f64(const char* path, const char* header, const char* labels, int64_t rows, int64_t cols, const double* values,
                      int format);

/* ---- instrumentation ---- */

/* number of kernels this context has launched (for bench.py's gpu_launches) */
int scs_launch_count(scs_ctx* ctx, int64_t* out);
/* average device time in microseconds of the E-step SpMM and M-step SpMM launches recorded with CUDA events
 * on the context's stream since the last reset; pass reset=1 to clear. out[0]=E us, out[1]=M us,
 * out[2]=#E-steps timed, out[3]=#M-steps timed, out[4]=us of statistics all-reduce per timed M-step (cells sharded over
 * ranks; the pieces overlap the M-step kernels), out[5]=#all-reduce pieces timed, out[6]=us of the Bayes kernel (posteriors, LL, statistics),
 * out[7]=#launches of it timed.  scs_set_timing(ctx, n): n = 0 off, n >= 1 times every n-th E-step and
 * every n-th M-step of the EM loop (the event records cost ~11 us per timed iteration at config 3, so a loop that is
 * itself being timed samples, e.g. n = 8); step-level calls outside the loop are always timed while n >= 1. */
int scs_set_timing(scs_ctx* ctx, int enable);
int scs_kernel_times(scs_ctx* ctx, double out[8], int reset);
/* select kernel variant (before scs_em_begin): 0 = auto (fixed-point E-step where the pool fits the format, sparse M-step
 * once posteriors saturate), 1 = v1 (warp per row, gathers from L1/L2; for A/B tests), 2 = tiled fp64 SpMM for both
 * half-steps (sparse M-step off), 3 = tiled fp64 E-step + sparse M-step (the fixed-point E-step off) */
int scs_set_variant(scs_ctx* ctx, int variant);
/* self-test of the device 2^x used by the Bayes step (scSplit:184, 188 evaluate `2 ** x` in numpy): y[i] = 2^x[i] for n host
 * doubles, computed on the context's device with the kernel's own routine */
int scs_selftest_exp2(scs_ctx* ctx, const double* x, double* y, int64_t n);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SCSPLIT_B200_H */
