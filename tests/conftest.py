import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["c1_k2_d0", "k5_d025", "k4_dnone", "k6_dead", "k3_ragged", "k5_nan_a", "k5_nan_b"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    from scipy.sparse import csr_matrix
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    V, N = int(d["V"]), int(d["N"])
    d["ref"] = csr_matrix((d["ref_data"], d["ref_indices"], d["ref_indptr"]), shape=(V, N))
    d["alt"] = csr_matrix((d["alt_data"], d["alt_indices"], d["alt_indptr"]), shape=(V, N))
    d["K"] = int(d["K"])
    return d


def split_lists(flat, ptr):
    return [flat[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]


def assert_close(got, want, rtol=1e-9, what=""):
    """Relative comparison that treats NaN==NaN and inf==inf as equal (the reference produces both)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), "%s: NaN pattern differs (%d vs %d)" % (what, nan_g.sum(), nan_w.sum())
    inf = np.isinf(want)
    assert np.array_equal(got[inf], want[inf]), "%s: inf pattern differs" % what
    fin = ~(nan_w | inf)
    err = np.abs(got[fin] - want[fin])
    tol = rtol * np.maximum(np.abs(want[fin]), 1e-300)
    bad = err > tol
    assert not bad.any(), "%s: %d values off, max rel err %.3e" % (
        what, bad.sum(), (err / np.maximum(np.abs(want[fin]), 1e-300)).max())


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    d = load_golden(request.param)
    d["name"] = request.param
    return d


def scipy_count_fn(ref, alt):
    """Host stand-in for Engine.pattern_counts (scs_pattern_counts) so the initialiser's loop logic can be tested on CPU."""
    U = ((ref != 0).astype(np.int8) + (alt != 0).astype(np.int8)).tocsr()
    U.data[:] = 1
    U = U.astype(np.int64)
    Ut = U.T.tocsr()

    def fn(keep_r, keep_c):
        kr = np.ones(U.shape[0], dtype=np.int64) if keep_r is None else keep_r.astype(np.int64)
        kc = np.ones(U.shape[1], dtype=np.int64) if keep_c is None else keep_c.astype(np.int64)
        return U @ kc, Ut @ kr
    return fn
