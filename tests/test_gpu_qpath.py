"""GPU tests of the fixed-point E-step (csrc/estep_lk.cuh, csrc/estep_q.cuh) through the C-ABI.

The E-step sums L[c,k] = sum_v A[v,c] log2 P[v,k] + R[v,c] log2(1-P[v,k]) (scSplit:175-178) run, by default, on a
55-bit fixed-point copy of the log table with exact integer accumulation.  These tests pin that path against
  * the fp64 kernels of the same library (variant 3) and the numpy oracle, at the north-star tolerance (1e-9 relative),
  * sums of the same table values in extended precision (np.longdouble), to show that the fixed-point path is at least as
    close to the exact sums as fp64 accumulation is,
and exercise its edges: counts above 1 (repeated codes up to 8, fp64 leftovers beyond), every row width (K <= 9: 64-byte rows, K <= 17: 128-byte rows,
K > 17: fp64), tables that do not fit the format (NaN / -inf: per-restart fp64 stand-in), pools the format cannot hold.
"""
import numpy as np
import pytest

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def E():
    from scsplit_b200 import _lib
    if _lib.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    from scsplit_b200.engine import Engine
    return Engine


def _random_P(V, K, seed):
    rng = np.random.default_rng(seed)
    P = rng.choice([0.01, 0.5, 0.99, 1e-5, 0.3], size=(V, K), p=[0.3, 0.3, 0.2, 0.05, 0.15])
    return P * rng.uniform(0.9, 1.0, size=(V, K))


def _estep(E, pool, P, variant):
    from scsplit_b200._lib import SCS_L, SCS_W
    with E(pool.ref, pool.alt) as eng:
        eng.set_variant(variant)
        eng.em_begin(P)
        info = eng.em_info()
        eng.estep()
        return eng.get(SCS_L), eng.get(SCS_W), eng.stats()["ll"], info


@pytest.mark.parametrize("K", [1, 2, 5, 8, 9, 10, 16, 17, 18])
def test_fixed_point_estep_matches_fp64_and_oracle(K, E):
    from oracle import em_oracle as O
    from scsplit_b200 import synth
    pool = synth.make_pool(3000, 700, max(K - 1, 1), 0.05 if K > 2 else 0.0, 0.08, seed=300 + K)
    P = _random_P(3000, K, K)
    Lq, Wq, llq, info = _estep(E, pool, P, 0)
    Lf, Wf, llf, info_f = _estep(E, pool, P, 3)
    assert info_f["q_lanes"] == 0
    assert info["q_lanes"] == (4 if K <= 9 else 8 if K <= 17 else 0), info
    if info["q_lanes"]:
        assert 40 <= info["q_frac_bits"] <= 49, info
    L_o = O.loglik_matrix(pool.ref, pool.alt, P)
    assert_close(Lq, L_o, RTOL, "L fixed point vs oracle K=%d" % K)
    assert_close(Lq, Lf, 1e-12, "L fixed point vs fp64 kernels K=%d" % K)
    assert_close(llq, llf, 1e-11, "LL K=%d" % K)
    assert np.array_equal(Wq >= 0.99, Wf >= 0.99)


def test_fixed_point_sums_are_closer_to_exact_than_fp64(E):
    """Extended-precision sums of the table values are the yardstick: the fixed-point path may be off by at most
    depth * 2^-(F+1) per cell (its only error is the table quantisation), and must not be worse than fp64 accumulation."""
    from scsplit_b200 import synth
    V, N, K = 6000, 400, 9
    pool = synth.make_pool(V, N, 8, 0.05, 0.25, seed=77)           # ~1500 reads per cell
    P = _random_P(V, K, 5)
    Lq, _, _, info = _estep(E, pool, P, 0)
    Lf, _, _, _ = _estep(E, pool, P, 3)
    F = info["q_frac_bits"]
    lp, lq = np.log2(P).astype(np.longdouble), np.log2(1 - P).astype(np.longdouble)
    A, R = pool.alt.tocsc(), pool.ref.tocsc()
    err_q = err_f = 0.0
    for c in range(0, N, 7):
        a, r = A.getcol(c).tocoo(), R.getcol(c).tocoo()
        exact = (a.data.astype(np.longdouble)[:, None] * lp[a.row]).sum(axis=0) + (r.data.astype(np.longdouble)[:, None] * lq[r.row]).sum(axis=0)
        depth = float(a.data.sum() + r.data.sum())
        eq = np.abs(Lq[c].astype(np.longdouble) - exact).max()
        ef = np.abs(Lf[c].astype(np.longdouble) - exact).max()
        bound = depth * 2.0 ** -(F + 1) + 2.0 ** -52 * float(np.abs(exact).max())    # quantisation + the final int -> double rounding
        assert eq <= bound, (c, float(eq), bound)
        err_q, err_f = max(err_q, float(eq)), max(err_f, float(ef))
    print("max |L - exact|: fixed point %.3e, fp64 kernels %.3e (F = %d)" % (err_q, err_f, F))
    assert err_q <= 4 * max(err_f, 1e-13)


@pytest.mark.parametrize("K", [4, 9, 17])
def test_fixed_point_heavy_counts(K, E):
    """Counts up to 8 are repeated codes, larger ones are added by the combine kernel in fp64."""
    from scipy.sparse import csr_matrix
    rng = np.random.default_rng(K)
    V, N = 900, 130
    dense_a = (rng.random((V, N)) < 0.06) * rng.choice([1, 2, 3, 15, 16, 17, 31, 45, 200, 3000], size=(V, N))
    dense_r = (rng.random((V, N)) < 0.09) * rng.choice([1, 1, 1, 2, 14, 15, 16, 30, 46, 999], size=(V, N))
    dense_r[:, 3] = 0
    dense_a[:, 3] = 0                                     # an empty cell
    dense_a[5, :] = 0
    dense_r[5, :] = 0                                     # an empty SNV
    ref, alt = csr_matrix(dense_r.astype(np.int32)), csr_matrix(dense_a.astype(np.int32))

    class Pool:
        pass
    pool = Pool()
    pool.ref, pool.alt = ref, alt
    P = _random_P(V, K, 11)
    Lq, Wq, llq, info = _estep(E, pool, P, 0)
    assert info["q_lanes"] in (4, 8)
    Lf, Wf, llf, _ = _estep(E, pool, P, 3)
    L_o = dense_a.T.astype(np.float64) @ np.log2(P) + dense_r.T.astype(np.float64) @ np.log2(1 - P)
    assert_close(Lq, L_o, RTOL, "L heavy counts")
    assert_close(Lq, Lf, 1e-12, "L heavy counts vs fp64")
    assert np.all(Lq[3] == 0.0) and not np.signbit(Lq[3]).any()


@pytest.mark.parametrize("K,V,N,dens", [(9, 9000, 700, 0.05), (4, 300, 37, 0.3), (17, 5000, 300, 0.04), (12, 2500, 2100, 0.02), (9, 12000, 90, 0.4)])
def test_lockstep_form_equals_row_group_form_bitwise(K, V, N, dens, E, monkeypatch):
    """The two forms of the fixed-point E-step (estep_lk.cuh: a lane per entry, totals by integer RED; estep_q.cuh: a group of
    lanes per entry, per-unit sums + combine) add the same integers, so L must agree bit for bit -- over several table tiles
    (V = 9000 / 12000: 2V rows against 3626 per tile), with counts above 8, empty cells, two restarts, and again when the step
    is repeated (the totals are left at zero by k_bayes)."""
    from scipy.sparse import csr_matrix
    from scsplit_b200._lib import SCS_L, SCS_W
    rng = np.random.default_rng(K * 1000 + N)
    depth = (rng.random((V, N)) < dens) * rng.choice([1, 1, 1, 1, 2, 3, 8, 9, 40], size=(V, N))
    alt = rng.binomial(depth, 0.4)
    ref = depth - alt
    alt[:, N // 2] = 0
    ref[:, N // 2] = 0
    ref_s, alt_s = csr_matrix(ref.astype(np.int32)), csr_matrix(alt.astype(np.int32))
    P = np.stack([_random_P(V, K, 1), _random_P(V, K, 2)])
    out = {}
    forms = ("lock-step", "row groups") + (("lock-step wide",) if K > 9 else ())
    for form in forms:
        monkeypatch.delenv("SCS_Q_GROUPS", raising=False)
        monkeypatch.delenv("SCS_Q_WIDE", raising=False)
        if form == "row groups":
            monkeypatch.setenv("SCS_Q_GROUPS", "1")
        if form == "lock-step wide":                      # one pass over 128-byte rows instead of two over 64-byte rows
            monkeypatch.setenv("SCS_Q_WIDE", "1")
        with E(ref_s, alt_s) as eng:
            eng.em_begin(P)
            info = eng.em_info()
            assert info["q_form"] == {"lock-step": "lock-step" if K <= 9 else "lock-step, two 64-byte tables", "row groups": "row groups",
                                      "lock-step wide": "lock-step"}[form], info
            eng.estep()
            first = [eng.get(SCS_L, r) for r in range(2)]
            eng.estep()
            again = [eng.get(SCS_L, r) for r in range(2)]
            eng.em_iterate(3, check_conv=False)
            out[form] = (first, again, [eng.get(SCS_L, r) for r in range(2)], [eng.get(SCS_W, r) for r in range(2)])
    L_o = alt.T.astype(np.float64) @ np.log2(P[0]) + ref.T.astype(np.float64) @ np.log2(1 - P[0])
    assert_close(out["lock-step"][0][0], L_o, RTOL, "L lock-step vs dense product")
    for r in range(2):
        assert np.array_equal(out["lock-step"][0][r], out["lock-step"][1][r]), "a repeated E-step must give the same sums"
        for form in forms[1:]:
            for part in range(4):
                assert np.array_equal(out["lock-step"][part][r], out[form][part][r], equal_nan=True), (form, part, r)


def test_unfit_table_takes_fp64_per_restart(E):
    """A restart whose table holds NaN / -inf (dying cluster, scSplit:198 -> 184) runs its E-step in fp64; the restart
    batched next to it keeps the fixed-point path; both give what they give alone."""
    from scsplit_b200 import synth
    from scsplit_b200._lib import SCS_L, SCS_W
    V, N, K = 2500, 300, 5
    pool = synth.make_pool(V, N, 4, 0.05, 0.1, seed=9)
    P_ok = _random_P(V, K, 1)
    P_bad = P_ok.copy()
    P_bad[:, 2] = np.nan
    P_bad[7, 0] = 1.0                                      # log2(1 - P) = -inf
    P_far = P_ok.copy()
    P_far[11, 1] = 1e-25                                   # log2 = -83: outside the format, still finite
    single = {}
    for name, P in (("ok", P_ok), ("bad", P_bad), ("far", P_far)):
        with E(pool.ref, pool.alt) as eng:
            eng.set_variant(3)
            eng.em_begin(P)
            eng.em_iterate(3, check_conv=False)
            single[name] = (eng.get(SCS_L), eng.get(SCS_W), eng.em_status()[2][0])
    with E(pool.ref, pool.alt) as eng:
        eng.em_begin(np.stack([P_ok, P_bad, P_far]))
        info = eng.em_info()
        assert info["q_lanes"] == 4 and info["q_unfit_restarts"] == 2, info
        eng.em_iterate(3, check_conv=False)
        for r, name in enumerate(("ok", "bad", "far")):
            L, W, ll = eng.get(SCS_L, r), eng.get(SCS_W, r), eng.em_status()[2][r]
            assert_close(L, single[name][0], 1e-10, "L restart %s" % name)
            assert np.array_equal(np.isnan(W), np.isnan(single[name][1]))
            assert_close(ll, single[name][2], 1e-10, "LL restart %s" % name)
        # the far-away value is gone after one M-step, so that restart is back on the fixed-point path
        assert eng.em_info()["q_unfit_restarts"] == 1


def test_pool_too_deep_for_the_format_keeps_fp64(E):
    """A cell so deep (2^21 reads) that the fixed-point fraction would drop below 40 bits: the context keeps the fp64
    E-step as a whole and still matches the oracle."""
    from scipy.sparse import csr_matrix
    rng = np.random.default_rng(3)
    V, N, K = 400, 60, 3
    a = (rng.random((V, N)) < 0.1) * rng.integers(1, 4, size=(V, N))
    r = (rng.random((V, N)) < 0.1) * rng.integers(1, 4, size=(V, N))
    a[10, 0] = 1 << 21
    ref, alt = csr_matrix(r.astype(np.int32)), csr_matrix(a.astype(np.int32))
    P = _random_P(V, K, 2)
    from scsplit_b200._lib import SCS_L
    with E(ref, alt) as eng:
        eng.em_begin(P)
        assert eng.em_info()["q_lanes"] == 0
        eng.estep()
        L = eng.get(SCS_L)
    L_o = a.T.astype(np.float64) @ np.log2(P) + r.T.astype(np.float64) @ np.log2(1 - P)
    assert_close(L, L_o, RTOL, "L deep pool")


@pytest.mark.parametrize("name", ["k5_nan_a", "k5_nan_b", "k6_dead"])
def test_dead_cluster_runs_match_fp64_variant(name, E):
    """Whole runs through a dying cluster (lP_s -> -inf -> NaN, pandas skipna LL = 0.0): the default path (fixed point with
    the per-restart fp64 stand-in) against the all-fp64 variant: same NaN pattern, same assignments, same stop."""
    from scsplit_b200._lib import SCS_P, SCS_W
    g = load_golden(name)
    out = {}
    for variant in (0, 3):
        with E(g["ref"], g["alt"]) as eng:
            eng.set_variant(variant)
            eng.em_begin(g["P0"])
            eng.em_iterate(int(g["iters"]) + 2, check_conv=False)
            out[variant] = (eng.get(SCS_W), eng.get(SCS_P), eng.em_status()[2][0])
    W0, P0, ll0 = out[0]
    W3, P3, ll3 = out[3]
    assert np.array_equal(np.isnan(W0), np.isnan(W3))
    with np.errstate(invalid="ignore"):
        assert np.array_equal(W0 >= 0.99, W3 >= 0.99)
    assert_close(P0, P3, 1e-9, "P %s" % name)
    assert_close(ll0, ll3, 1e-9, "LL %s" % name)


@pytest.mark.parametrize("seed", range(12))
def test_random_pools_run_against_oracle(seed, E):
    """Randomised differential test of whole runs: shapes, state counts, depths and count distributions drawn at random;
    five iterations from a random seed assignment (stop rule off) against the numpy oracle: L, W, P, lP_s, LL at 1e-9,
    identical >= 0.99 assignments; and the same bits from a second run (graph replay, batched next to a copy)."""
    from scipy.sparse import csr_matrix
    from oracle import em_oracle as O
    from scsplit_b200._lib import SCS_L, SCS_LPS, SCS_P, SCS_W
    rng = np.random.default_rng(1000 + seed)
    V = int(rng.integers(40, 5000))
    N = int(rng.integers(8, 1500))
    K = int(rng.choice([1, 2, 3, 5, 8, 9, 10, 13, 16, 17, 18, 24, 32]))
    dens = float(rng.choice([0.01, 0.05, 0.2, 0.6]))
    top = int(rng.choice([1, 2, 9, 40, 500]))
    n_s = max(1, min(K, 6))
    G = rng.choice([0.01, 0.5, 0.99], size=(V, n_s), p=[0.5, 0.3, 0.2])
    truth = rng.integers(0, n_s, N)
    depth = (rng.random((V, N)) < dens) * rng.integers(1, top + 1, size=(V, N))
    alt = rng.binomial(depth, G[:, truth])
    ref = depth - alt
    if N > 3:
        alt[:, 1] = 0
        ref[:, 1] = 0                                      # an empty cell
    ref_s, alt_s = csr_matrix(ref.astype(np.int32)), csr_matrix(alt.astype(np.int32))
    lab = np.full(N, -1, dtype=np.int32)
    pick = rng.choice(N, min(N, 2 * K), replace=False)
    lab[pick] = np.arange(pick.size) % K
    k_alt = O.kalt(ref_s, alt_s)
    P0 = O.seed_af(ref_s, alt_s, k_alt, lab, K)
    want = O.run_em(ref_s, alt_s, P0, k_alt, fixed_iters=5)
    with E(ref_s, alt_s) as eng:
        assert np.array_equal(eng.seed_af(lab, K), P0)
        eng.em_begin(np.stack([P0, P0]))
        eng.em_iterate(5, check_conv=False)
        it, cv, ll = eng.em_status()
        got = [(eng.get(SCS_L, r), eng.get(SCS_W, r), eng.get(SCS_P, r), eng.get(SCS_LPS, r)) for r in range(2)]
    what = "seed %d: V=%d N=%d K=%d density %.2f counts<=%d" % (seed, V, N, K, dens, top)
    for r in range(2):
        assert_close(got[r][0], want["L"], RTOL, "L " + what)
        assert_close(got[r][2], want["P"], RTOL, "P " + what)
        assert_close(got[r][3], want["lPs"], RTOL, "lP_s " + what)
        assert_close(ll[r], want["LL"], RTOL, "LL " + what)
        with np.errstate(invalid="ignore"):
            assert np.array_equal(got[r][1] >= 0.99, want["W"] >= 0.99), what
    for a, b in zip(got[0], got[1]):
        assert np.array_equal(a, b, equal_nan=True), "batched copies differ: " + what
