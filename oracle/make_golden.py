"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` by RUNNING THE REFERENCE.

Run in the build container (``/root/reference`` must exist):

    python oracle/make_golden.py

The reference ships no tests or fixtures ("parity unpinned"), so the
goldens are outputs of the reference's own ``models`` class (loaded through
``oracle/ref_loader.py``: in-memory dtype patch, no arithmetic changes) on small seeded
synthetic pools.  ``np.random.seed`` is set by this harness because the reference never
seeds.  Each fixture stores its inputs (CSR arrays), the reference's initial P (so EM parity
does not depend on sklearn / argsort tie-breaking of the machine that replays it), a
per-iteration trace, the state at the reference's own stopping rule and every post-EM product.
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from scsplit_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> pool parameters, model parameters
CASES = {
    # BASELINE.json configs[0]: 2,000 x 500, K=2, -d 0
    "c1_k2_d0": dict(V=2000, N=500, samples=2, dfrac=0.0, lam=0.15, pseed=1001, K=2, dbl=0.0, seed=2001, trace=6),
    # doublet state + refinement (-d 0.25)
    "k5_d025": dict(V=1500, N=400, samples=4, dfrac=0.08, lam=0.2, pseed=11, K=5, dbl=0.25, seed=7, trace=6),
    # doublet state, flag absent (doublets = -1): copy-only refine
    "k4_dnone": dict(V=1200, N=300, samples=3, dfrac=0.05, lam=0.15, pseed=12, K=4, dbl=-1.0, seed=3, trace=4),
    # more states than samples: a cluster dies (lP_s -> -inf, possibly NaN) -- skipna semantics
    "k6_dead": dict(V=900, N=300, samples=2, dfrac=0.0, lam=0.25, pseed=13, K=6, dbl=-1.0, seed=5, trace=8),
    # very sparse, ragged: empty cells and empty SNVs exist
    "k3_ragged": dict(V=700, N=257, samples=3, dfrac=0.0, lam=0.02, pseed=14, K=3, dbl=0.0, seed=9, trace=4),
    # a cluster's mass underflows to 0: lP_s = -inf, then NaN everywhere; the loop exits on bitwise-equal LL
    "k5_nan_a": dict(V=3000, N=250, samples=3, dfrac=0.0, lam=0.4, pseed=21, K=5, dbl=-1.0, seed=3, trace=7, post=False),
    # same, but one more iteration runs: all-NaN -> skipna sums give LL = 0.0 and the run "converges" there
    "k5_nan_b": dict(V=3000, N=250, samples=3, dfrac=0.0, lam=0.4, pseed=21, K=5, dbl=-1.0, seed=4, trace=9, post=False),
}


def _lists_to_index(lists, barcodes):
    pos = {b: i for i, b in enumerate(barcodes)}
    flat = np.array([pos[b] for lst in lists for b in lst], dtype=np.int32)
    ptr = np.cumsum([0] + [len(lst) for lst in lists]).astype(np.int32)
    return flat, ptr


def run_case(name, c, g):
    models = g["models"]
    pool = synth.make_pool(c["V"], c["N"], c["samples"], c["dfrac"], c["lam"], c["pseed"])
    if name == "k3_ragged":
        # a few heavy SNVs so that counts > 1 and both-allele entries are exercised
        rng = np.random.default_rng(99)
        extra = synth.make_pool(c["V"], c["N"], c["samples"], 0.0, 3.0, c["pseed"] + 1000)
        rows = rng.choice(c["V"], 12, replace=False)
        sel = np.zeros(c["V"], dtype=bool)
        sel[rows] = True
        from scipy.sparse import diags
        D = diags(sel.astype(np.int16))
        ref_r, alt_r = (pool.ref + D @ extra.ref).tolil(), (pool.alt + D @ extra.alt).tolil()
        for cc in (5, 100, 256):          # cells without any read
            ref_r[:, cc] = 0; alt_r[:, cc] = 0
        for vv in (0, 17, 350, 699):      # SNVs without any read
            ref_r[vv, :] = 0; alt_r[vv, :] = 0
        ref_r, alt_r = ref_r.tocsr(), alt_r.tocsr()
        ref_r.eliminate_zeros(); alt_r.eliminate_zeros()
        pool = pool._replace(ref=ref_r, alt=alt_r)
    ref = pool.ref.astype(np.int64)
    alt = pool.alt.astype(np.int64)
    ref.sort_indices(); alt.sort_indices()
    mtx = [ref, alt, pd.Index(pool.snv_ids, name="SNV"), pd.Index(pool.barcodes)]
    K = c["K"]
    out = dict(K=K, dbl=c["dbl"], seed=c["seed"], V=c["V"], N=c["N"],
               ref_indptr=ref.indptr.astype(np.int64), ref_indices=ref.indices.astype(np.int32), ref_data=ref.data.astype(np.int16),
               alt_indptr=alt.indptr.astype(np.int64), alt_indices=alt.indices.astype(np.int32), alt_data=alt.data.astype(np.int16),
               snv_ids=np.array(pool.snv_ids), barcodes=np.array(pool.barcodes))
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        # --- step-level trace (manual replay of run_EM's body, scSplit:156-161)
        np.random.seed(c["seed"])
        m = models(mtx, K, tmp)
        assert m.model_af.sum().sum() > 0, "init failed for %s" % name
        out["k_alt"] = np.asarray(m.k_alt).ravel()
        out["P0"] = m.model_af.values.astype(np.float64)
        init_flat, init_ptr = _lists_to_index(m.initial, pool.barcodes)
        out["initial_flat"], out["initial_ptr"] = init_flat, init_ptr
        for t in range(c["trace"]):
            m.calculate_cell_likelihood()
            out["L_%d" % t] = m.lP_c_s.values.astype(np.float64)
            out["W_%d" % t] = m.P_s_c.values.astype(np.float64)
            out["LL_%d" % t] = np.float64(m.lP_c_m)
            m.calculate_model_af()
            out["P_%d" % t] = m.model_af.values.astype(np.float64)
            out["lPs_%d" % t] = np.asarray(m.lP_s, dtype=np.float64)
        out["trace"] = c["trace"]
        # --- the reference's own loop, from the same init
        np.random.seed(c["seed"])
        m = models(mtx, K, tmp)
        assert np.array_equal(m.model_af.values, out["P0"])
        m.run_EM(tmp)
        m.assign_cells()
        out["iters"] = len(m.sum_log_likelihood) - 2
        out["hist"] = np.array(m.sum_log_likelihood[2:], dtype=np.float64)
        out["convergence"] = m.convergence
        out["W_final"] = m.P_s_c.values.astype(np.float64)
        out["L_final"] = m.lP_c_s.values.astype(np.float64)
        out["P_final"] = m.model_af.values.astype(np.float64)
        out["lPs_final"] = np.asarray(m.lP_s, dtype=np.float64)
        out["LL_final"] = np.float64(m.lP_c_m)
        out["assigned_flat"], out["assigned_ptr"] = _lists_to_index(m.assigned, pool.barcodes)
        out["post"] = int(c.get("post", True))
        if not c.get("post", True):
            return out
        # --- post-EM
        m.define_doublet()
        out["doublet_defined"] = m.doublet
        m.refine_doublets(c["dbl"])
        out["doublet"] = m.doublet
        out["reassigned_flat"], out["reassigned_ptr"] = _lists_to_index(m.reassigned, pool.barcodes)
        m.distinguishing_alleles([])
        out["pa_matrix"] = m.pa_matrix.values.astype(np.float64)
        out["pa_index"] = np.array(m.pa_matrix.index.tolist())
        out["pa_columns"] = np.array(m.pa_matrix.columns.tolist(), dtype=np.int32)
        out["dist_variants"] = np.array(sorted(m.dist_variants))
        out["dist_matrix"] = m.dist_matrix.reindex(sorted(m.dist_variants)).values.astype(np.float64)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if os.path.exists("scSplit_dist_variants.txt"):   # CWD side effect of scSplit:323
            os.remove("scSplit_dist_variants.txt")
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    g = ref_loader.load_reference()
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = run_case(name, c, g)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-12s iters=%3d conv=%d doublet=%2d LL=%r nanW=%d  -> %s (%.1f KB)" % (
            name, out["iters"], out["convergence"], out.get("doublet", -9), float(out["LL_final"]),
            int(np.isnan(out["W_final"]).sum()), os.path.relpath(path, ROOT), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
