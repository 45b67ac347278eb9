"""Synthetic Poisson-sparse ref/alt allele-count matrices.

The reference ships no data; BASELINE.json quotes its metric on "synthetic Poisson-sparse
count matrices of the named shapes".  The recipe:

  genotype   G[v, s] in {0.01, 0.5, 0.99} with probability {0.5, 0.3, 0.2}
  cell c     drawn from a uniformly chosen sample s(c); a fraction ``doublet_frac`` of
             cells are doublets whose alt probability is the mean of two distinct samples
  depth      D[v, c] ~ Poisson(lam)            (lam = 0.05  ->  4.88 % non-zero)
  alt        A[v, c] ~ Binomial(D, p[v, c]),   R = D - A

Only the non-zero depths are ever materialised: positions are generated per block of
cells as an iid Bernoulli(rho) process through geometric gaps, depths from the
zero-truncated Poisson, so the 60k x 200k configuration never needs a dense matrix.
Every block of ``BLOCK`` cells has its own counter-based seed, so a rank that holds
cells [lo, hi) generates exactly the columns a single process would.
"""
from collections import namedtuple

import numpy as np
from scipy.sparse import csr_matrix, csc_matrix

BLOCK = 1024

# name -> (V, N, n_samples, K_states, doublet_frac, seed)   K_states includes the doublet state
CONFIGS = {
    "c1": dict(V=2000, N=500, samples=2, K=2, doublet_frac=0.0, seed=1001),
    "c2": dict(V=10000, N=5000, samples=4, K=5, doublet_frac=0.05, seed=1002),
    "c3": dict(V=30000, N=20000, samples=8, K=9, doublet_frac=0.05, seed=1003),
    "c4": dict(V=60000, N=200000, samples=16, K=17, doublet_frac=0.05, seed=1004),
    "c5": dict(V=20000, N=10000, samples=8, K=8, doublet_frac=0.0, seed=1005),
}

Pool = namedtuple("Pool", "ref alt snv_ids barcodes truth is_doublet genotypes")


def snv_ids(V):
    return ["%d:%d" % (1 + v % 22, 1000 + v) for v in range(V)]


def barcodes(lo, hi):
    return ["BC%07d-1" % c for c in range(lo, hi)]


def _truncated_poisson(rng, lam, size):
    """Zero-truncated Poisson(lam) by inverse CDF (values >= 1)."""
    u = rng.random(size)
    pmf = np.exp(-lam) * lam / (1.0 - np.exp(-lam))  # P(D=1 | D>=1)
    out = np.ones(size, dtype=np.int64)
    cdf = pmf
    k = 1
    todo = u > cdf
    while todo.any() and k < 64:
        k += 1
        pmf = pmf * lam / k
        cdf = cdf + pmf
        out[todo] = k
        todo = u > cdf
    return out


def make_genotypes(V, samples, seed):
    rng = np.random.default_rng([seed, 2**31 - 1])
    return rng.choice(np.array([0.01, 0.5, 0.99]), size=(V, samples), p=[0.5, 0.3, 0.2])


def _gen_block(args):
    """Non-zero depths of one BLOCK of cells: (SNV rows, local cell columns, alt, ref, truth, truth2, doublet flags, index range)."""
    b, V, N, samples, doublet_frac, lam, seed, cell_lo, cell_hi = args
    G = make_genotypes(V, samples, seed)
    rho = 1.0 - np.exp(-lam)
    lo, hi = b * BLOCK, min((b + 1) * BLOCK, N)
    nb = hi - lo
    rng = np.random.default_rng([seed, b])
    s1 = rng.integers(0, samples, nb)
    s2 = (s1 + 1 + rng.integers(0, max(samples - 1, 1), nb)) % samples
    dbl = rng.random(nb) < doublet_frac
    # iid Bernoulli(rho) over the nb*V linearised (cell-major) positions via geometric gaps
    total = nb * V
    m = int(total * rho + 8 * np.sqrt(total * rho) + 64)
    pos = np.cumsum(rng.geometric(rho, m)) - 1
    while pos[-1] < total:
        extra = np.cumsum(rng.geometric(rho, m)) + pos[-1]
        pos = np.concatenate([pos, extra])
    pos = pos[pos < total]
    c_loc = pos // V
    v = pos - c_loc * V
    depth = _truncated_poisson(rng, lam, pos.size)
    p = G[v, s1[c_loc]]
    pd2 = 0.5 * (G[v, s1[c_loc]] + G[v, s2[c_loc]])
    p = np.where(dbl[c_loc], pd2, p)
    a = rng.binomial(depth, p)
    r = depth - a
    keep = (lo + c_loc >= cell_lo) & (lo + c_loc < cell_hi)
    sel = (np.arange(lo, hi) >= cell_lo) & (np.arange(lo, hi) < cell_hi)
    idx = np.arange(lo, hi)[sel] - cell_lo
    return (v[keep].astype(np.int32), (lo + c_loc[keep] - cell_lo).astype(np.int32), a[keep].astype(np.int16), r[keep].astype(np.int16),
            s1[sel], s2[sel], dbl[sel], idx)


def make_pool(V, N, samples, doublet_frac=0.05, lam=0.05, seed=0, cell_lo=0, cell_hi=None,
              dtype=np.int16, workers=1):
    """Return a Pool for cells [cell_lo, cell_hi) of a V x N experiment (SNV-major CSR, V x n).

    workers > 1 samples the blocks of cells in that many threads (every block has its own seed, so the result is the same)."""
    cell_hi = N if cell_hi is None else cell_hi
    assert 0 <= cell_lo <= cell_hi <= N
    G = make_genotypes(V, samples, seed)
    rows_l, cols_l, alt_l, ref_l = [], [], [], []
    truth = np.zeros(cell_hi - cell_lo, dtype=np.int32)
    truth2 = np.zeros(cell_hi - cell_lo, dtype=np.int32)
    isdbl = np.zeros(cell_hi - cell_lo, dtype=bool)
    b0, b1 = cell_lo // BLOCK, (max(cell_hi, 1) - 1) // BLOCK
    jobs = [(b, V, N, samples, doublet_frac, lam, seed, cell_lo, cell_hi) for b in range(b0, b1 + 1)]
    if workers > 1 and len(jobs) > 1:
        # threads, not processes: the numpy samplers release the GIL on these array sizes, and the caller may hold a CUDA context
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(min(workers, len(jobs))) as pool:
            blocks = list(pool.map(_gen_block, jobs))
    else:
        blocks = map(_gen_block, jobs)
    for v, c, a, r, s1, s2, dbl, idx in blocks:
        rows_l.append(v)
        cols_l.append(c)
        alt_l.append(a)
        ref_l.append(r)
        truth[idx], truth2[idx], isdbl[idx] = s1, s2, dbl
    rows = np.concatenate(rows_l)
    cols = np.concatenate(cols_l)
    a = np.concatenate(alt_l)
    r = np.concatenate(ref_l)
    n = cell_hi - cell_lo
    # cell-major generation order == CSC order of the V x n matrix; convert once to CSR
    def to_csr(vals):
        nz = vals > 0
        indptr = np.zeros(n + 1, dtype=np.int64)
        counts = np.bincount(cols[nz], minlength=n)
        indptr[1:] = np.cumsum(counts)
        m_csc = csc_matrix((vals[nz].astype(dtype), rows[nz].astype(np.int32), indptr), shape=(V, n))
        m_csr = m_csc.tocsr()
        m_csr.sort_indices()
        return m_csr
    return Pool(to_csr(r), to_csr(a), snv_ids(V), barcodes(cell_lo, cell_hi), truth, isdbl, G)


def make_config(name, lam=0.05, cell_lo=0, cell_hi=None, V=None, N=None, workers=1):
    cfg = dict(CONFIGS[name])
    if V is not None:
        cfg["V"] = V
    if N is not None:
        cfg["N"] = N
    return make_pool(cfg["V"], cfg["N"], cfg["samples"], cfg["doublet_frac"], lam, cfg["seed"],
                     cell_lo, cell_hi, workers=workers)


def union_nnz(ref, alt):
    """nnz of the union pattern of A and R (the unit SURVEY.md section 8d counts bytes in)."""
    return int(((ref != 0).astype(np.int8) + (alt != 0).astype(np.int8)).nnz)


def write_csv(pool, ref_path, alt_path):
    """Dense CSVs in the reference's input format (scSplit:59, 462-463, 544-545): header `SNV,<barcodes>`, rows `<id>,<ints>`.

    Rows are rendered from a `0,0,...,0` byte template with the single-digit counts patched in (rows holding a count
    >= 10 take the generic path), so the 2 x 1.2 GB files of config 3 are written in seconds.
    """
    header = ("SNV," + ",".join(pool.barcodes) + "\n").encode()
    N = len(pool.barcodes)
    template = np.frombuffer(b"0," * N, dtype=np.uint8).copy()
    template[-1] = ord("\n")
    for mat, path in ((pool.ref, ref_path), (pool.alt, alt_path)):
        mat = mat.tocsr()
        with open(path, "wb") as f:
            f.write(header)
            for i, sid in enumerate(pool.snv_ids):
                lo, hi = mat.indptr[i], mat.indptr[i + 1]
                cols, vals = mat.indices[lo:hi], mat.data[lo:hi]
                f.write(sid.encode() + b",")
                if vals.size and vals.max() >= 10:
                    dense = np.zeros(N, dtype=np.int64)
                    dense[cols] = vals
                    f.write((",".join(map(str, dense.tolist())) + "\n").encode())
                    continue
                line = template.copy()
                line[2 * cols.astype(np.int64)] = ord("0") + vals.astype(np.uint8)
                f.write(line.tobytes())
