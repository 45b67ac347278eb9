"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's `scSplit run` EM path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm may import this module; the product (``scsplit_b200``) never does, and has no CPU path.

Every function cites the lines of ``/root/reference/scSplit`` it restates.  The reference
computes in numpy float64 through scipy.sparse and pandas; this oracle uses the same numpy
primitives (``np.log2``, ``np.power(2.0, x)``, scipy CSR products) so that the only freedom
left is floating-point summation order.  Two variants exist where that order matters:
``faithful=True`` reproduces the reference's operation sequence (per-cluster sparse scale,
sparse add, column sum -- slow), the default uses one SpMM per operand (fast).  Both are
pinned against goldens produced by the reference itself (``oracle/make_golden.py``,
``tests/golden/*.npz``, ``tests/test_oracle_vs_golden.py``).

pandas ``skipna`` semantics matter when a cluster dies (hard part 4):
``Series.sum``/``DataFrame.sum`` skip NaN (all-NaN -> 0.0), ``DataFrame.max`` skips NaN
(all-NaN -> NaN).  They are restated with ``np.nansum`` / ``np.nanmax``.
"""
import warnings

import numpy as np
from scipy.sparse import csr_matrix, diags

MAX_ITER = 300          # scSplit:156
ASSIGN_THRESHOLD = 0.99  # scSplit:207
PSEUDO = 1              # scSplit:98


def _f64(m):
    return m.astype(np.float64) if m.dtype != np.float64 else m


def kalt(ref, alt):
    """Background alt fraction per SNV, scSplit:100-103. Returns float64[V]."""
    n_alt = np.asarray(alt.sum(axis=1)).ravel() + PSEUDO
    n_ref = np.asarray(ref.sum(axis=1)).ravel() + PSEUDO
    return n_alt / (n_ref + n_alt)


def seed_af(ref, alt, k_alt, labels, K):
    """Hard (one-hot) M-step that seeds P, scSplit:137-145.

    labels: int[N], -1 = cell not in the initial subset, else its KMeans cluster.
    """
    V = ref.shape[0]
    P = np.zeros((V, K))
    for n in range(K):
        cols = np.flatnonzero(labels == n)
        b_alt = np.asarray(alt[:, cols].sum(axis=1)).ravel()
        b_ref = np.asarray(ref[:, cols].sum(axis=1)).ravel()
        P[:, n] = (b_alt + k_alt) / (b_alt + b_ref + PSEUDO)
    return P


def loglik_matrix(ref, alt, P, faithful=False):
    """L[c,k] = sum_v A[v,c] log2 P[v,k] + R[v,c] log2(1-P[v,k]), scSplit:175-178."""
    with np.errstate(divide="ignore", invalid="ignore"):
        lp = np.log2(P)
        lq = np.log2(1 - P)
    if not faithful:
        return np.asarray(_f64(alt).T @ lp + _f64(ref).T @ lq)
    N, K = ref.shape[1], P.shape[1]
    L = np.zeros((N, K))
    for n in range(K):
        # same shape of computation as scSplit:176-178 (scale stored entries, sparse add, column sum)
        matcalc = diags(lp[:, n]) @ _f64(alt) + diags(lq[:, n]) @ _f64(ref)
        L[:, n] = np.asarray(matcalc.sum(axis=0)).ravel()
    return L


def posterior(L, lPs):
    """W[c,i] = 1 / sum_j 2^(L_j + s_j - L_i - s_i), j ascending, scSplit:181-185."""
    N, K = L.shape
    W = np.zeros((N, K))
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for i in range(K):
            denom = 0
            for j in range(K):
                denom = denom + np.power(2.0, L[:, j] + lPs[j] - L[:, i] - lPs[i])
            W[:, i] = 1 / denom
    return W


def total_loglik(L, lPs):
    """Model log-likelihood with pandas skipna semantics, scSplit:188."""
    with warnings.catch_warnings(), np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        warnings.simplefilter("ignore", RuntimeWarning)
        m = np.nanmax(L, axis=1)                       # DataFrame.max(axis=1): skipna, all-NaN -> NaN
        t = np.power(2.0, (L - m[:, None]) + np.asarray(lPs)[None, :])
        s = np.nansum(np.ascontiguousarray(t.T), axis=0)  # DataFrame.sum(axis=1): column-by-column adds
        ll_c = np.log2(s) + m
        return float(np.nansum(ll_c)), ll_c           # Series.sum(): skipna, all-NaN -> 0.0


def estep(ref, alt, P, lPs, faithful=False):
    """calculate_cell_likelihood, scSplit:165-188. Returns (L, W, LL)."""
    L = loglik_matrix(ref, alt, P, faithful)
    W = posterior(L, lPs)
    LL, _ = total_loglik(L, lPs)
    return L, W, LL


def mstep(ref, alt, W, k_alt):
    """calculate_model_af, scSplit:191-198. Returns (P, lPs)."""
    a = _f64(alt)
    with np.errstate(invalid="ignore", divide="ignore"):
        P = (a.dot(W) + k_alt[:, None]) / ((a + _f64(ref)).dot(W) + PSEUDO)
        col = np.nansum(np.ascontiguousarray(W.T), axis=1)   # P_s_c.sum(axis=0), skipna
        lPs = np.log2(col) - np.log2(np.nansum(col))
    return np.asarray(P), lPs


def run_em(ref, alt, P0, k_alt, max_iter=MAX_ITER, faithful=False, trace=False, fixed_iters=None):
    """run_EM, scSplit:148-162.  ``fixed_iters`` disables the bitwise stop rule (benchmarks)."""
    K = P0.shape[1]
    P = P0.copy()
    lPs = np.array([np.log2(1 / K)] * K)            # scSplit:91
    hist = [1, 2]                                   # scSplit:155
    it = 0
    steps = []
    L = W = None
    while (it < max_iter) and ((hist[-2] != hist[-1]) if fixed_iters is None else it < fixed_iters):
        it += 1
        L, W, LL = estep(ref, alt, P, lPs, faithful)
        P, lPs = mstep(ref, alt, W, k_alt)
        hist.append(LL)
        if trace:
            steps.append(dict(L=L, W=W, LL=LL, P=P, lPs=lPs))
    out = dict(P=P, lPs=lPs, L=L, W=W, LL=hist[-1], hist=hist[2:], iters=it,
               convergence=int(it < max_iter))
    if trace:
        out["steps"] = steps
    return out


def assign_mask(W):
    """assign_cells, scSplit:201-207: boolean [N,K], W >= 0.99 (NaN compares False)."""
    with np.errstate(invalid="ignore"):
        return W >= ASSIGN_THRESHOLD


def define_doublet(ref, alt, P, W, faithful=False):
    """define_doublet, scSplit:210-225: first argmax_i of sum_j sum_{c in assigned[j]} L_i(c), post-M P."""
    mask = assign_mask(W)
    K = P.shape[1]
    if not faithful:
        L = loglik_matrix(ref, alt, P)
        per_cluster = mask.sum(axis=1)               # a cell counted once per cluster that holds it
        result = (L * per_cluster[:, None]).sum(axis=0)
        result = result.tolist()
        return result.index(max(result)), np.array(result)
    with np.errstate(divide="ignore", invalid="ignore"):
        lp, lq = np.log2(P), np.log2(1 - P)
    cross = np.zeros((K, K))
    for i in range(K):
        matcalc = (diags(lp[:, i]) @ _f64(alt) + diags(lq[:, i]) @ _f64(ref)).tocsr()
        for j in range(K):
            idx = np.flatnonzero(mask[:, j])
            cross[j, i] = matcalc[:, idx].sum()
    result = cross.sum(axis=0).tolist()
    return result.index(max(result)), np.array(result)


def refine_scores(alt, P, W):
    """scSplit:243: rps = A^T (1-P) * (W >= 0.99)  -> float64[N,K]."""
    return np.asarray(_f64(alt).T.dot(1 - P)) * assign_mask(W)


def refine_doublets(alt, P, W, doublet, doublets):
    """refine_doublets, scSplit:228-252, on cell indices.

    Returns (assigned, reassigned, doublet): lists of index lists per cluster.  ``assigned`` here is
    the list *after* the reference's aliasing quirk (shallow copy at scSplit:234, ``+=`` at 252).
    Barcode sorting (scSplit:207) is the caller's job; indices are returned in ascending order.
    """
    N, K = W.shape
    mask = assign_mask(W)
    assigned = [np.flatnonzero(mask[:, n]).tolist() for n in range(K)]
    reassigned = list(assigned)
    if doublets == 0:
        return assigned, reassigned, -1
    if doublets > 0:
        rps = refine_scores(alt, P, W)
        rpc = np.delete(rps, doublet, axis=1)
        score = rpc.sum(axis=1)
        lack = N * doublets - len(assigned[doublet])
        if lack > 0:
            order = score.argsort()                 # Series.argsort -> ndarray.argsort (quicksort)
            found = order[int(N - lack):N]
            fset = set(found.tolist())
            for n in range(K):
                if n != doublet:
                    reassigned[n] = [x for x in assigned[n] if x not in fset]
                else:
                    reassigned[n] += found.tolist()  # mutates assigned[doublet] too (aliasing)
    return assigned, reassigned, doublet


def cluster_counts(ref, alt, labels_or_lists, K):
    """Hard per-cluster sums, scSplit:269-275. Accepts per-cluster index lists. int64 [V,K] x2."""
    V = ref.shape[0]
    n_ref = np.zeros((V, K), dtype=np.int64)
    n_alt = np.zeros((V, K), dtype=np.int64)
    for n in range(K):
        idx = sorted(set(labels_or_lists[n]))       # bc_idx is built by membership, scSplit:270
        n_ref[:, n] = np.asarray(ref[:, idx].sum(axis=1)).ravel()
        n_alt[:, n] = np.asarray(alt[:, idx].sum(axis=1)).ravel()
    return n_ref, n_alt


def pa_codes(n_ref, n_alt):
    """scSplit:279-282 before column/row drops: +1 -> ALT present, -1 -> absent (REF only), 0 -> unknown."""
    return ((n_alt >= 10) * 1 - ((n_ref >= 10) & (n_alt == 0)) * 1).astype(np.int8)


def trim_subset(ref, alt, K, rng=np.random):
    """Dense restatement of the coverage-trimming loop, scSplit:106-125. Returns (irows, icols).

    Consumes ``rng.beta`` exactly as the reference consumes the global ``np.random``.
    """
    base_mtx = (alt + ref).toarray()
    rows, cols = [*range(base_mtx.shape[0])], [*range(base_mtx.shape[1])]
    irows, icols = np.array(rows), np.array(cols)
    nrows, ncols = len(rows), len(cols)
    mrows = mcols = 0
    minimum = int(K / 0.9 + 0.5) + 1
    while (mrows < 0.9 * nrows or mcols < 0.9 * ncols) and ncols >= minimum:
        rbrows = np.sort(np.unique(list(map(int, rng.beta(1, 10, int(0.1 * nrows + 0.5)) * nrows))))
        rbcols = np.sort(np.unique(list(map(int, rng.beta(1, 10, int(0.1 * ncols + 0.5)) * ncols))))
        if len(rbrows) == 0 and len(rbcols) == 0:
            break
        rows = np.count_nonzero(base_mtx, axis=1).argsort().tolist()
        cols = np.count_nonzero(base_mtx, axis=0).argsort().tolist()
        srb, scb = set(rbrows.tolist()), set(rbcols.tolist())
        rows = [item for index, item in enumerate(rows) if index not in srb]
        cols = [item for index, item in enumerate(cols) if index not in scb]
        irows, icols = irows[rows], icols[cols]
        nrows, ncols = len(rows), len(cols)
        base_mtx = base_mtx[rows][:, cols]
        mrows = min(np.count_nonzero(base_mtx, axis=0))
        mcols = min(np.count_nonzero(base_mtx, axis=1))
    return irows, icols, nrows, ncols


def initial_labels(ref, alt, irows, icols, nrows, ncols, K):
    """PCA + KMeans hand-off, scSplit:126-135.  Returns labels int[N] (-1 unused) or None (init failed)."""
    from sklearn.cluster import KMeans
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import StandardScaler
    alt_subset = alt[irows][:, icols].todense()
    ref_subset = ref[irows][:, icols].todense()
    alt_prop = (alt_subset + 0.01) / (alt_subset + ref_subset + 0.02)
    alt_pca = StandardScaler().fit_transform(np.array(alt_prop.T))
    pca = PCA(n_components=min(nrows, ncols, 20))
    pca_alt = pca.fit_transform(alt_pca)
    if pca_alt.shape[0] < K:
        return None
    kmeans = KMeans(n_clusters=K, random_state=0).fit(pca_alt)
    labels = np.full(ref.shape[1], -1, dtype=np.int32)
    labels[icols] = kmeans.labels_
    return labels


def initialise(ref, alt, K, rng=np.random):
    """models.__init__ end to end, scSplit:100-145. Returns (P0 or None, labels, k_alt)."""
    k = kalt(ref, alt)
    irows, icols, nrows, ncols = trim_subset(ref, alt, K, rng)
    labels = initial_labels(ref, alt, irows, icols, nrows, ncols, K)
    if labels is None:
        return None, None, k
    return seed_af(ref, alt, k, labels, K), labels, k
