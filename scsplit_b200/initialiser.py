"""Sparse restatement of the reference initialiser (models.__init__, scSplit:105-135).

The reference densifies A+R (V x N int64: 4.8 GB at 30k x 20k, 96 GB at 60k x 200k -- scSplit:106) and then
repeatedly drops low-coverage rows and columns at random sorted positions until the surviving block is 90 % dense
(scSplit:112-125).  Only three things of the dense matrix are ever read: the number of non-zeros per surviving row,
per surviving column, and the final tiny block.  This module keeps the row/column ORDER and the `np.random` /
`np.argsort` call sequence of the reference, so it selects exactly the same cells, while the counts come from a
`count_fn(keep_rows, keep_cols) -> (row_counts, col_counts)` -- in the product the CUDA kernels behind
`scs_pattern_counts` (two launches over the resident entry streams per pass, nothing is densified).

`np.argsort` (introsort / x86-simd-sort) is not stable: tie-breaking depends on the numpy build and the host ISA,
for the reference exactly as for this restatement.  Feeding it the same int64 count vector in the same order is what
makes the two agree on a given machine.
"""
import numpy as np


def _positions_to_drop(rng, n):
    """scSplit:113-114: sorted unique int(beta(1,10) * n) positions, int(0.1*n + 0.5) draws."""
    draws = rng.beta(1, 10, int(0.1 * n + 0.5)) * n
    pos = np.sort(draws.astype(np.int64))         # int() truncation of non-negative floats; sorted unique values (= np.unique)
    return pos[np.concatenate(([True], pos[1:] != pos[:-1]))] if pos.size else pos


def trim_subset(count_fn, V, N, K, rng=np.random):
    """Coverage-trimming loop, scSplit:107-125.  Returns (irows, icols, nrows, ncols) like the reference's locals."""
    irows, icols = np.arange(V), np.arange(N)
    nrows, ncols = V, N
    mrows = mcols = 0
    minimum = int(K / 0.9 + 0.5) + 1
    keep_r = np.ones(V, dtype=np.uint8)
    keep_c = np.ones(N, dtype=np.uint8)
    row_cnt, col_cnt = count_fn(None, None)
    while (mrows < 0.9 * nrows or mcols < 0.9 * ncols) and ncols >= minimum:
        rbrows = _positions_to_drop(rng, nrows)
        rbcols = _positions_to_drop(rng, ncols)
        if len(rbrows) == 0 and len(rbcols) == 0:
            break
        # np.count_nonzero(base_mtx, axis=1).argsort(): an int64 vector in the CURRENT row order (scSplit:117-118)
        rows = np.ascontiguousarray(row_cnt[irows], dtype=np.int64).argsort()
        cols = np.ascontiguousarray(col_cnt[icols], dtype=np.int64).argsort()
        rows = np.delete(rows, rbrows)                  # scSplit:119-120: drop the sorted POSITIONS, keep sorted order
        cols = np.delete(cols, rbcols)
        irows, icols = irows[rows], icols[cols]
        nrows, ncols = len(rows), len(cols)
        keep_r[:] = 0
        keep_r[irows] = 1
        keep_c[:] = 0
        keep_c[icols] = 1
        row_cnt, col_cnt = count_fn(keep_r, keep_c)
        mrows = col_cnt[icols].min()                    # scSplit:124 (named mrows, is the min COLUMN fill)
        mcols = row_cnt[irows].min()                    # scSplit:125
    return irows, icols, nrows, ncols


def initial_labels(ref, alt, irows, icols, nrows, ncols, K):
    """StandardScaler -> PCA(min(nrows, ncols, 20)) -> KMeans(K, random_state=0) on the surviving block, scSplit:126-135.

    Returns int32 labels[N] (-1 = cell not in the block) or None when there are fewer cells than states
    (scSplit:132-133: the reference logs a message and leaves model_af all-zero).
    """
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
