"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes -> libscsplit_b200.so), against
(a) goldens produced by the reference itself and (b) the CPU oracle on the same inputs.

Tolerances: bit-exact for integer work (k_alt operands, seed P, cluster sums, P/A codes, assignments);
1e-9 relative for L, LL, P, lP_s (BASELINE.json north_star).
"""
import numpy as np
import pytest

from conftest import assert_close, load_golden, split_lists

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def E():
    from scsplit_b200 import _lib
    if _lib.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    from scsplit_b200.engine import Engine
    return Engine


def _labels_from_lists(lists, N):
    lab = np.full(N, -1, dtype=np.int32)
    for n, lst in enumerate(lists):
        lab[lst] = n
    return lab


def test_layout_kalt_seed(golden, E):
    g = golden
    with E(g["ref"], g["alt"]) as eng:
        info = eng.layout_info()
        assert info["nnz_ref"] == g["ref"].nnz and info["nnz_alt"] == g["alt"].nnz
        union = ((g["ref"] != 0).astype(np.int8) + (g["alt"] != 0).astype(np.int8)).nnz
        assert info["nnz_union"] == union
        assert info["light"] + info["heavy"] == g["ref"].nnz + g["alt"].nnz
        assert info["light"] == int((g["ref"].data == 1).sum() + (g["alt"].data == 1).sum())
        assert np.array_equal(eng.kalt(), g["k_alt"])                       # scSplit:100-103, bit-exact
        lab = _labels_from_lists(split_lists(g["initial_flat"], g["initial_ptr"]), int(g["N"]))
        assert np.array_equal(eng.seed_af(lab, g["K"]), g["P0"])            # scSplit:137-145, bit-exact


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_step_trace(golden, E, variant):
    from oracle import em_oracle as O
    from scsplit_b200._lib import SCS_L, SCS_LPS, SCS_P, SCS_W
    g = golden
    K = g["K"]
    with E(g["ref"], g["alt"]) as eng:
        eng.set_variant(variant)
        eng.em_begin(g["P0"])
        P, lPs = g["P0"], np.array([np.log2(1 / K)] * K)
        for t in range(int(g["trace"])):
            eng.set(SCS_P, P)
            eng.set(SCS_LPS, lPs)
            eng.estep()
            L, W = eng.get(SCS_L), eng.get(SCS_W)
            assert_close(L, g["L_%d" % t], RTOL, "L step %d" % t)
            # Bayes kernel, exactly: against the oracle's posterior evaluated on the GPU's own L
            assert_close(W, O.posterior(L, lPs), RTOL, "W(L_gpu) step %d" % t)
            Wg = g["W_%d" % t]
            assert np.array_equal(np.isnan(W), np.isnan(Wg))
            ok = ~np.isnan(Wg)
            assert np.allclose(W[ok], Wg[ok], rtol=1e-9, atol=1e-300)
            assert_close(eng.stats()["ll"], g["LL_%d" % t], RTOL, "LL step %d" % t)   # device-side skipna LL
            eng.set(SCS_W, Wg)
            eng.mstep()
            assert_close(eng.get(SCS_P), g["P_%d" % t], RTOL, "P step %d" % t)
            assert_close(eng.get(SCS_LPS), g["lPs_%d" % t], RTOL, "lPs step %d" % t)
            P, lPs = g["P_%d" % t], g["lPs_%d" % t]


def test_iterate_matches_reference_trace(golden, E):
    """em_iterate (fused loop, device-side LL) reproduces the reference's LL history over the traced iterations."""
    from scsplit_b200._lib import SCS_LLHIST, SCS_P, SCS_W
    g = golden
    T = int(g["trace"])
    with E(g["ref"], g["alt"]) as eng:
        eng.em_begin(g["P0"])
        eng.em_iterate(T, check_conv=False)
        hist = eng.get(SCS_LLHIST)[:T]
        want = np.array([g["LL_%d" % t] for t in range(T)])
        finite = np.isfinite(want) & (want != 0.0)
        # errors compound over iterations but stay far below the tolerance until the NaN regime starts
        assert_close(hist[finite], want[finite], 1e-9, "LL history")
        it, cv, ll = eng.em_status()
        assert it[0] == T
        if not np.isnan(g["P_%d" % (T - 1)]).any():
            assert_close(eng.get(SCS_P), g["P_%d" % (T - 1)], RTOL, "P after %d iterations" % T)
            W, Wg = eng.get(SCS_W), g["W_%d" % (T - 1)]
            assert np.array_equal(W >= 0.99, Wg >= 0.99)


def test_run_to_convergence(golden, E):
    from scsplit_b200._lib import SCS_LLHIST, SCS_P, SCS_W
    g = golden
    with E(g["ref"], g["alt"]) as eng:
        eng.em_begin(g["P0"])
        eng.em_run()
        it, cv, ll = eng.em_status()
        W = eng.get(SCS_W)
        hist = eng.get(SCS_LLHIST)
        assert cv[0] == int(g["convergence"])
        assert 1 <= it[0] <= 300
        assert ll[0] == hist[it[0] - 1]
        if it[0] >= 2 and it[0] < 300:
            assert hist[it[0] - 1] == hist[it[0] - 2]            # stopped by the bitwise rule, scSplit:156
        if np.isnan(g["W_final"]).any():
            # dead-cluster runs: the end state is one of the two the reference can reach (see oracle test)
            assert ll[0] == 0.0 or abs(ll[0] - g["hist"][int(g["trace"]) - 3]) < 1e-6
            return
        with np.errstate(invalid="ignore"):
            assert np.array_equal(W >= 0.99, g["W_final"] >= 0.99)   # identical assignments
        assert_close(ll[0], g["LL_final"], RTOL, "LL final")
        # (this run stops by ITS bitwise LL equality, possibly an iteration before or after the reference's: P is compared
        # at the reference's own iteration count, at 1e-9, in test_end_state_at_the_reference_iteration_count)
        assert_close(eng.get(SCS_P), g["P_final"], 1e-6, "P final")


def test_end_state_at_the_reference_iteration_count(golden, E):
    """Exactly as many iterations as the reference ran (stop rule off): P, lP_s, L, W and LL against the reference's end
    state at the north-star tolerance.  The bitwise stop rule (scSplit:156) makes iteration COUNTS summation-order
    dependent; the states at equal counts are not."""
    from scsplit_b200._lib import SCS_LLHIST, SCS_P, SCS_W
    g = golden
    if np.isnan(g["W_final"]).any():
        pytest.skip("dead-cluster golden: its NaN end state is covered by test_dead_cluster_runs_match_fp64_variant")
    n = int(g["iters"])
    with E(g["ref"], g["alt"]) as eng:
        eng.em_begin(g["P0"])
        eng.em_iterate(n, check_conv=False)
        it, cv, ll = eng.em_status()
        assert it[0] == n
        assert_close(ll[0], g["LL_final"], RTOL, "LL after %d iterations" % n)
        assert_close(eng.get(SCS_P), g["P_final"], RTOL, "P after %d iterations" % n)
        W, Wg = eng.get(SCS_W), g["W_final"]
        assert np.array_equal(W >= 0.99, Wg >= 0.99)
        assert np.allclose(W, Wg, rtol=1e-9, atol=1e-300)
        assert_close(eng.get(SCS_LLHIST)[:n], g["hist"][:n], RTOL, "LL history")


def test_restart_batch_is_bitwise_independent(E):
    """R batched restarts give bit-identical results to running each alone (restart axis = extra grid dimension)."""
    from scsplit_b200._lib import SCS_LLHIST, SCS_P, SCS_W
    g = load_golden("k5_d025")
    rng = np.random.default_rng(5)
    P0s = [g["P0"]]
    for _ in range(3):
        P0s.append(np.clip(g["P0"] * rng.uniform(0.7, 1.3, g["P0"].shape), 1e-3, 1 - 1e-3))
    P0s = np.stack(P0s)
    with E(g["ref"], g["alt"]) as eng:
        eng.em_begin(P0s)
        eng.em_run()
        it_b, cv_b, ll_b = eng.em_status()
        Wb = [eng.get(SCS_W, r) for r in range(4)]
        Pb = [eng.get(SCS_P, r) for r in range(4)]
        for r in range(4):
            eng.em_begin(P0s[r])
            eng.em_run()
            it, cv, ll = eng.em_status()
            assert it[0] == it_b[r] and ll[0] == ll_b[r]
            assert np.array_equal(eng.get(SCS_W), Wb[r], equal_nan=True)
            assert np.array_equal(eng.get(SCS_P), Pb[r], equal_nan=True)


def test_run_is_deterministic(E):
    from scsplit_b200._lib import SCS_W
    g = load_golden("k6_dead")
    res = []
    for _ in range(2):
        with E(g["ref"], g["alt"]) as eng:
            eng.em_begin(g["P0"])
            eng.em_run()
            res.append((eng.em_status(), eng.get(SCS_W)))
    assert res[0][0][0][0] == res[1][0][0][0]
    assert res[0][0][2][0] == res[1][0][2][0]
    assert np.array_equal(res[0][1], res[1][1], equal_nan=True)


def test_post_em(golden, E):
    from oracle import em_oracle as O
    g = golden
    K, N = g["K"], int(g["N"])
    P, W = g["P_final"], g["W_final"]
    if not int(g["post"]):
        # NaN end state (dead cluster): the reference's own post-EM was not recorded (no cell reaches 0.99, the doublet call
        # is an argmax over NaN); the kernels still have to carry the NaN pattern exactly like the numpy restatement does
        mask = O.assign_mask(W)
        with E(g["ref"], g["alt"]) as eng:
            scores = eng.doublet_scores(P, mask.sum(axis=1).astype(np.uint8))
            d_o, s_o = O.define_doublet(g["ref"], g["alt"], P, W, faithful=True)     # (empty clusters sum to 0, not NaN)
            assert_close(scores, s_o, RTOL, "doublet scores (NaN end state)")
            lab = np.where(mask.any(axis=1), mask.argmax(axis=1), -1).astype(np.int32)
            rs = eng.refine_scores(P, lab)
            want = np.nansum(np.where(mask, O.refine_scores(g["alt"], P, W), 0.0), axis=1) if mask.any() else np.zeros(N)
            assert_close(rs, want, RTOL, "refine scores (NaN end state)")
            lists = [np.flatnonzero(lab == n).tolist() for n in range(K)]
            n_ref, n_alt, pa = eng.cluster_counts(lab, K)
            o_ref, o_alt = O.cluster_counts(g["ref"], g["alt"], lists, K)
            assert np.array_equal(n_ref, o_ref) and np.array_equal(n_alt, o_alt)
            assert np.array_equal(pa, O.pa_codes(o_ref, o_alt))
        return
    mask = O.assign_mask(W)
    with E(g["ref"], g["alt"]) as eng:
        scores = eng.doublet_scores(P, mask.sum(axis=1).astype(np.uint8))
        d_o, s_o = O.define_doublet(g["ref"], g["alt"], P, W)
        assert_close(scores, s_o, RTOL, "doublet scores")
        assert int(np.argmax(scores)) == int(g["doublet_defined"])             # scSplit:225 (first max)
        lab = np.where(mask.any(axis=1), mask.argmax(axis=1), -1).astype(np.int32)
        rs = eng.refine_scores(P, lab)
        want = O.refine_scores(g["alt"], P, W).sum(axis=1)
        assert_close(rs, want, RTOL, "refine scores")
        re_lists = split_lists(g["reassigned_flat"], g["reassigned_ptr"])
        lab2 = _labels_from_lists(re_lists, N)
        n_ref, n_alt, pa = eng.cluster_counts(lab2, K)
        o_ref, o_alt = O.cluster_counts(g["ref"], g["alt"], re_lists, K)
        assert np.array_equal(n_ref, o_ref) and np.array_equal(n_alt, o_alt)    # integer-exact, scSplit:273
        assert np.array_equal(pa, O.pa_codes(o_ref, o_alt))
        cols = [n for n in range(K) if n != int(g["doublet"])]
        paf = pa[:, cols].astype(np.float64)
        paf[paf == 0] = np.nan
        paf[paf == -1] = 0
        keep = [i for i, s in enumerate(g["snv_ids"].tolist()) if s[0] not in ("X", "Y", "M")]
        assert_close(paf[keep], g["pa_matrix"], 0, "pa_matrix")


def test_sparse_mstep_matches_dense(E):
    """The saturated-posterior M-step kernel (variant 0) against the tiled SpMM (variant 2) and the oracle on posteriors that
    are (a) fully saturated, (b) saturated with a few soft cells (gathered as dense rows), (c) mostly soft (gate -> dense)."""
    from oracle import em_oracle as O
    from scsplit_b200._lib import SCS_LPS, SCS_P, SCS_W
    g = load_golden("k5_d025")
    K, N = g["K"], int(g["N"])
    rng = np.random.default_rng(3)
    lab = rng.integers(0, K, N)
    hard = np.full((N, K), 1e-120)
    hard[np.arange(N), lab] = 1.0 - rng.random(N) * 1e-13
    few = hard.copy()
    soft_rows = rng.choice(N, 7, replace=False)
    few[soft_rows] = rng.dirichlet(np.ones(K), 7)
    mostly = rng.dirichlet(np.ones(K), N)
    for name, W in (("saturated", hard), ("few soft", few), ("mostly soft", mostly)):
        out = {}
        for variant in (0, 2):
            with E(g["ref"], g["alt"]) as eng:
                eng.set_variant(variant)
                eng.em_begin(g["P0"])
                eng.set(SCS_W, W)
                eng.mstep()
                out[variant] = (eng.get(SCS_P), eng.get(SCS_LPS), eng.stats()["S"].copy())
        P_o, lPs_o = O.mstep(g["ref"], g["alt"], W, g["k_alt"])
        for variant in (0, 2):
            assert_close(out[variant][0], P_o, RTOL, "%s P variant %d" % (name, variant))
            assert_close(out[variant][1], lPs_o, RTOL, "%s lPs variant %d" % (name, variant))
        assert np.allclose(out[0][2], out[2][2], rtol=1e-12, atol=1e-30)


@pytest.mark.parametrize("K", [2, 8, 9, 16, 17, 24, 32])
def test_sparse_mstep_state_counts(K, E):
    """Saturated posteriors for every row-group width of the sparse M-step kernel (incl. K > 16: all column pairs stored)."""
    from oracle import em_oracle as O
    from scsplit_b200 import synth
    from scsplit_b200._lib import SCS_P, SCS_W
    V, N = 1500, 500
    pool = synth.make_pool(V, N, 3, 0.0, 0.2, seed=300 + K)
    rng = np.random.default_rng(K)
    W = np.full((N, K), 1e-200)
    W[np.arange(N), rng.integers(0, K, N)] = 1.0 + (rng.random(N) - 0.5) * 2e-13
    W[rng.choice(N, 3, replace=False)] = rng.dirichlet(np.ones(K), 3)          # a few dense rows
    res = {}
    for variant in (0, 2):
        with E(pool.ref, pool.alt) as eng:
            eng.set_variant(variant)
            eng.em_begin(np.full((V, K), 0.5))
            eng.set(SCS_W, W)
            eng.mstep()
            res[variant] = (eng.get(SCS_P), eng.stats()["S"].copy())
    P_o, _ = O.mstep(pool.ref, pool.alt, W, O.kalt(pool.ref, pool.alt))
    assert_close(res[0][0], P_o, RTOL, "sparse P K=%d" % K)
    assert_close(res[2][0], P_o, RTOL, "dense P K=%d" % K)
    assert np.allclose(res[0][1], res[2][1], rtol=1e-12, atol=1e-30)


def test_pattern_counts(E):
    g = load_golden("k3_ragged")
    U = ((g["ref"] != 0).astype(np.int8) + (g["alt"] != 0).astype(np.int8)).tocsr()
    U.data[:] = 1
    rng = np.random.default_rng(1)
    with E(g["ref"], g["alt"]) as eng:
        rc, cc = eng.pattern_counts()
        assert np.array_equal(rc, np.asarray(U.sum(axis=1)).ravel())
        assert np.array_equal(cc, np.asarray(U.sum(axis=0)).ravel())
        kr = rng.random(U.shape[0]) < 0.6
        kc = rng.random(U.shape[1]) < 0.5
        rc, cc = eng.pattern_counts(kr, kc)
        assert np.array_equal(rc[kr], np.asarray(U[:, kc].sum(axis=1)).ravel()[kr])
        assert np.array_equal(cc[kc], np.asarray(U[kr].sum(axis=0)).ravel()[kc])


def test_errors(E):
    from scipy.sparse import csr_matrix
    from scsplit_b200._lib import ScsError
    g = load_golden("k3_ragged")
    with pytest.raises(ValueError):
        E(g["ref"], g["alt"][:, :10])
    neg = g["ref"].copy().astype(np.int64)
    neg.data[0] = -1
    with pytest.raises(ValueError):
        E(neg, g["alt"])
    with E(g["ref"], g["alt"]) as eng:
        with pytest.raises(ScsError):
            eng.estep()                                   # EM state not initialised
        with pytest.raises(ScsError):
            eng.em_begin(np.full((int(g["V"]), 40), 0.5))  # K > SCS_MAX_K
        with pytest.raises(ValueError):
            eng.em_begin(np.full((5, 3), 0.5))             # wrong V
        # arrays of the wrong length stop at the Python boundary (the C-ABI takes plain pointers of implied length)
        V, N = int(g["V"]), int(g["N"])
        for call in (lambda: eng.pattern_counts(np.ones(V - 1, np.uint8), None), lambda: eng.pattern_counts(None, np.ones(N + 1, np.uint8)),
                     lambda: eng.seed_af(np.zeros(N - 1, np.int32), 3), lambda: eng.doublet_scores(np.full((V, 3), 0.5), np.ones(N - 1, np.uint8)),
                     lambda: eng.doublet_scores(np.full((V + 1, 3), 0.5), np.ones(N, np.uint8)),
                     lambda: eng.refine_scores(np.full((V, 3), 0.5), np.zeros(N + 2, np.int32)),
                     lambda: eng.cluster_counts(np.zeros(N - 1, np.int32), 3)):
            with pytest.raises(ValueError):
                call()
    empty = csr_matrix((int(g["V"]), int(g["N"])), dtype=np.int16)
    with E(empty, empty) as eng:                           # all-zero input is legal: k_alt = 1/2 everywhere
        assert np.array_equal(eng.kalt(), np.full(int(g["V"]), 0.5))


def test_full_size_properties(E):
    """BASELINE config 2 shape (10k x 5k, K=5): size-independent properties instead of a full oracle run.

    (1) posteriors sum to one per cell, (2) P in (0,1), (3) LL is non-decreasing over iterations (EM),
    (4) linearity of the M-step statistics: S(W1) + S(W2) == S(W1 + W2) to rounding,
    (5) hard cluster sums over all clusters add up to the global row sums (integer-exact).
    """
    from scsplit_b200 import synth
    from scsplit_b200._lib import SCS_LLHIST, SCS_P, SCS_STATS, SCS_W
    pool = synth.make_config("c2")
    K = 5
    with E(pool.ref, pool.alt) as eng:
        V, N = eng.V, eng.N
        rng = np.random.default_rng(0)
        lab = np.full(N, -1, dtype=np.int32)
        pick = rng.choice(N, 40, replace=False)
        lab[pick] = np.arange(40) % K
        P0 = eng.seed_af(lab, K)
        assert ((P0 > 0) & (P0 < 1)).all()
        eng.em_begin(P0)
        eng.em_iterate(12, check_conv=False)
        W, P = eng.get(SCS_W), eng.get(SCS_P)
        hist = eng.get(SCS_LLHIST)[:12]
        assert np.allclose(W.sum(axis=1), 1.0, rtol=0, atol=1e-9)
        assert ((P > 0) & (P < 1)).all()
        assert (np.diff(hist) >= -1e-6 * np.abs(hist[:-1])).all()
        # linearity of the M-step SpMM
        W1 = rng.random((N, K)); W2 = rng.random((N, K))
        stats = []
        for Wx in (W1, W2, W1 + W2):
            eng.set(SCS_W, Wx)
            eng.mstep()
            stats.append(eng.stats()["S"].copy())
        assert np.allclose(stats[0] + stats[1], stats[2], rtol=1e-12, atol=1e-9)
        # hard sums
        labels = rng.integers(0, K, N).astype(np.int32)
        n_ref, n_alt, _ = eng.cluster_counts(labels, K)
        assert np.array_equal(n_ref.sum(axis=1), np.asarray(pool.ref.sum(axis=1)).ravel())
        assert np.array_equal(n_alt.sum(axis=1), np.asarray(pool.alt.sum(axis=1)).ravel())


@pytest.mark.parametrize("K", [1, 3, 7, 8, 11, 12, 16, 17, 24, 32])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_state_count_sweep(K, variant, E):
    """Every compiled table width (KP = 2..32: row groups of 1, 2, 4, 8 and 16 lanes) against the oracle, incl. K = 17 of
    BASELINE config 4; the pool is sized so that the E-step operand spans several shared-memory tiles (2V*KP*8 > 208 KB)."""
    from oracle import em_oracle as O
    from scsplit_b200 import synth
    from scsplit_b200._lib import SCS_L, SCS_LPS, SCS_P, SCS_W
    V, N = 4000, 700
    pool = synth.make_pool(V, N, 4, 0.05, 0.12, seed=77 + K)
    rng = np.random.default_rng(K)
    lab = np.full(N, -1, dtype=np.int32)
    pick = rng.choice(N, 2 * K, replace=False)
    lab[pick] = np.arange(2 * K) % K
    with E(pool.ref, pool.alt) as eng:
        eng.set_variant(variant)
        k_alt = eng.kalt()
        P0 = eng.seed_af(lab, K)
        assert np.array_equal(P0, O.seed_af(pool.ref, pool.alt, k_alt, lab, K))
        eng.em_begin(P0)
        P, lPs = P0, np.array([np.log2(1 / K)] * K)
        for t in range(2):
            eng.set(SCS_P, P)
            eng.set(SCS_LPS, lPs)
            eng.estep()
            L_o, W_o, LL_o = O.estep(pool.ref, pool.alt, P, lPs)
            L = eng.get(SCS_L)
            assert_close(L, L_o, RTOL, "L K=%d" % K)
            assert_close(eng.get(SCS_W), O.posterior(L, lPs), RTOL, "W K=%d" % K)
            assert_close(eng.stats()["ll"], LL_o, RTOL, "LL K=%d" % K)
            eng.set(SCS_W, W_o)
            eng.mstep()
            P_o, lPs_o = O.mstep(pool.ref, pool.alt, W_o, k_alt)
            assert_close(eng.get(SCS_P), P_o, RTOL, "P K=%d" % K)
            assert_close(eng.get(SCS_LPS), lPs_o, RTOL, "lPs K=%d" % K)
            P, lPs = P_o, lPs_o


def test_config3_full_size_properties(E):
    """BASELINE config 3 (30k x 20k, K = 9): the metric's own configuration, checked through size-independent properties
    and against the oracle on a column sample (the oracle's full E-step takes ~5 s, so one step is affordable)."""
    from oracle import em_oracle as O
    from scsplit_b200 import synth
    from scsplit_b200._lib import SCS_L, SCS_LLHIST, SCS_P, SCS_W
    pool = synth.make_config("c3")
    K = 9
    with E(pool.ref, pool.alt) as eng:
        V, N = eng.V, eng.N
        info = eng.layout_info()
        assert info["nnz_union"] == 29264030                       # the figure SURVEY.md section 8d quotes
        rng = np.random.default_rng(0)
        lab = np.full(N, -1, dtype=np.int32)
        lab[rng.choice(N, 27, replace=False)] = np.arange(27) % K
        k_alt = eng.kalt()
        assert np.array_equal(k_alt, O.kalt(pool.ref, pool.alt))
        P0 = eng.seed_af(lab, K)
        eng.em_begin(P0)
        eng.estep()
        L = eng.get(SCS_L)
        L_o = O.loglik_matrix(pool.ref, pool.alt, P0)
        assert_close(L, L_o, RTOL, "L at config 3")
        W = eng.get(SCS_W)
        assert np.allclose(np.nansum(W, axis=1), 1.0, rtol=0, atol=1e-9)
        eng.mstep()
        P1_o, lPs_o = O.mstep(pool.ref, pool.alt, W, k_alt)
        assert_close(eng.get(SCS_P), P1_o, RTOL, "P at config 3")
        eng.em_begin(P0)
        eng.em_iterate(10, check_conv=False)
        hist = eng.get(SCS_LLHIST)[:10]
        assert (np.diff(hist) >= -1e-9 * np.abs(hist[:-1])).all()   # EM never decreases the likelihood
        labels = rng.integers(0, K, N).astype(np.int32)
        n_ref, n_alt, _ = eng.cluster_counts(labels, K)
        assert np.array_equal(n_ref.sum(axis=1), np.asarray(pool.ref.sum(axis=1)).ravel())
        assert np.array_equal(n_alt.sum(axis=1), np.asarray(pool.alt.sum(axis=1)).ravel())


def test_skewed_coverage_against_oracle(E):
    """Non-Poisson input: SNV coverage and cell depth spread over two orders of magnitude, counts up to the hundreds, some
    SNVs with int32-sized totals -- the work plan (equal-cost CTA ranges / group cuts) and the count>1 path must hold."""
    from oracle import em_oracle as O
    from scipy.sparse import csr_matrix
    from scsplit_b200._lib import SCS_L, SCS_LPS, SCS_P, SCS_W
    rng = np.random.default_rng(11)
    V, N, K = 6000, 1500, 6
    snv_rate = np.exp(rng.normal(-3.5, 1.4, V))[:, None]            # log-normal coverage per SNV
    cell_depth = np.exp(rng.normal(0.0, 0.9, N))[None, :]            # and per cell
    lam = np.minimum(snv_rate * cell_depth, 40.0)
    depth = rng.poisson(lam)
    depth[rng.choice(V, 5, replace=False)[:, None], rng.choice(N, 40, replace=False)[None, :]] += 300   # a few huge counts
    geno = rng.choice([0.01, 0.5, 0.99], size=(V, 4), p=[0.5, 0.3, 0.2])
    p = geno[:, rng.integers(0, 4, N)]
    a = rng.binomial(depth, p)
    ref, alt = csr_matrix((depth - a).astype(np.int32)), csr_matrix(a.astype(np.int32))
    lab = np.full(N, -1, dtype=np.int32)
    lab[rng.choice(N, 3 * K, replace=False)] = np.arange(3 * K) % K
    for variant in (0, 1):
        with E(ref, alt) as eng:
            eng.set_variant(variant)
            info = eng.layout_info()
            assert info["heavy"] > 0 and info["light"] > 0
            k_alt = eng.kalt()
            assert np.array_equal(k_alt, O.kalt(ref, alt))
            P0 = eng.seed_af(lab, K)
            assert np.array_equal(P0, O.seed_af(ref, alt, k_alt, lab, K))
            eng.em_begin(P0)
            P, lPs = P0, np.array([np.log2(1 / K)] * K)
            for t in range(3):
                eng.set(SCS_P, P)
                eng.set(SCS_LPS, lPs)
                eng.estep()
                L_o, W_o, LL_o = O.estep(ref, alt, P, lPs)
                assert_close(eng.get(SCS_L), L_o, RTOL, "L skewed")
                assert_close(eng.stats()["ll"], LL_o, RTOL, "LL skewed")
                eng.set(SCS_W, W_o)
                eng.mstep()
                P_o, lPs_o = O.mstep(ref, alt, W_o, k_alt)
                assert_close(eng.get(SCS_P), P_o, RTOL, "P skewed")
                assert_close(eng.get(SCS_LPS), lPs_o, RTOL, "lPs skewed")
                P, lPs = P_o, lPs_o
            eng.em_begin(P0)
            eng.em_run()
            want = O.run_em(ref, alt, P0, k_alt)
            with np.errstate(invalid="ignore"):
                assert np.array_equal(eng.get(SCS_W) >= 0.99, want["W"] >= 0.99)


@pytest.mark.parametrize("V,N,K", [(1, 1, 1), (3, 2, 2), (5, 40, 3), (40, 5, 4), (2, 300, 2)])
def test_tiny_shapes(V, N, K, E):
    """Degenerate sizes: single SNV / single cell / fewer cells than a warp, K = 1; everything against the oracle."""
    from oracle import em_oracle as O
    from scipy.sparse import csr_matrix
    from scsplit_b200._lib import SCS_L, SCS_P, SCS_W
    rng = np.random.default_rng(V * 1000 + N)
    depth = rng.poisson(1.2, size=(V, N))
    a = rng.binomial(depth, 0.4)
    ref, alt = csr_matrix((depth - a).astype(np.int16)), csr_matrix(a.astype(np.int16))
    lab = (np.arange(N) % (K + 1) - 1).astype(np.int32)            # includes -1 (unused) cells
    with E(ref, alt) as eng:
        k_alt = eng.kalt()
        assert np.array_equal(k_alt, O.kalt(ref, alt))
        P0 = eng.seed_af(lab, K)
        assert np.array_equal(P0, O.seed_af(ref, alt, k_alt, lab, K))
        eng.em_begin(P0)
        eng.em_iterate(4, check_conv=False)
        want = O.run_em(ref, alt, P0, k_alt, fixed_iters=4)
        assert_close(eng.get(SCS_L), want["L"], RTOL, "L tiny")
        assert_close(eng.get(SCS_W), want["W"], 1e-7, "W tiny")
        assert_close(eng.get(SCS_P), want["P"], RTOL, "P tiny")
        assert_close(eng.em_status()[2][0], want["LL"], RTOL, "LL tiny")
        rc, cc = eng.pattern_counts()
        U = ((ref != 0).astype(np.int8) + (alt != 0).astype(np.int8))
        assert np.array_equal(rc, np.asarray((U != 0).sum(axis=1)).ravel())
        assert np.array_equal(cc, np.asarray((U != 0).sum(axis=0)).ravel())


def test_device_exp2_matches_numpy(E):
    """The Bayes step's own 2**x (scSplit:184, 188 use numpy's) over the whole exponent range, specials included."""
    from scipy.sparse import csr_matrix
    rng = np.random.default_rng(7)
    x = np.concatenate([
        rng.uniform(-1080, 1030, 200000), rng.uniform(-2, 2, 100000), rng.normal(0, 1e-13, 50000),
        rng.uniform(-1075, -1020, 50000),                       # results in the denormal range
        np.arange(-1080, 1030, dtype=np.float64), np.arange(-1080, 1030) + 0.5,
        [0.0, -0.0, 1023.9999999999999, 1024.0, 1e300, -1e300, np.inf, -np.inf, np.nan, -1074.0, -1074.5, -1075.0, -1076.0],
    ])
    m = csr_matrix(np.ones((2, 2), dtype=np.int16))
    with E(m, m) as eng:
        y = eng.selftest_exp2(x)
    with np.errstate(over="ignore", under="ignore"):
        want = np.exp2(x)
    assert np.array_equal(np.isnan(y), np.isnan(want))
    fin = np.isfinite(want) & ~np.isnan(want)
    assert np.array_equal(y[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)])          # +inf where numpy overflows
    normal = fin & (want >= 2.0 ** -1022)
    ulp = np.spacing(want[normal])
    assert np.max(np.abs(y[normal] - want[normal]) / ulp) <= 2.0
    sub = fin & (want < 2.0 ** -1022)                                                       # denormals and exact zeros
    assert np.max(np.abs(y[sub] - want[sub])) <= 2.0 ** -1074


def test_sparse_mstep_long_rows_with_soft_cells(E):
    """Pseudo-rows far longer than the sparse M-step's per-lane replay mask covers (few SNVs seen in most of 6000 cells), with
    2 % unsaturated cells: the flagged-element replay and its whole-row fallback against the tiled SpMM and the oracle."""
    from oracle import em_oracle as O
    from scsplit_b200 import synth
    from scsplit_b200._lib import SCS_P, SCS_W
    V, N, K = 40, 6000, 5
    pool = synth.make_pool(V, N, 3, 0.0, 1.5, seed=77)                 # lam = 1.5: ~78 % of the cells per SNV
    rng = np.random.default_rng(5)
    W = np.full((N, K), 1e-200)
    W[np.arange(N), rng.integers(0, K, N)] = 1.0 + (rng.random(N) - 0.5) * 2e-13
    soft = rng.choice(N, N // 50, replace=False)
    W[soft] = rng.dirichlet(np.ones(K), soft.size)
    res = {}
    for variant in (0, 2):
        with E(pool.ref, pool.alt) as eng:
            eng.set_variant(variant)
            eng.em_begin(np.full((V, K), 0.5))
            eng.set(SCS_W, W)
            eng.mstep()
            res[variant] = (eng.get(SCS_P), eng.stats())
    assert res[0][1]["dense_cells"] == soft.size
    P_o, _ = O.mstep(pool.ref, pool.alt, W, O.kalt(pool.ref, pool.alt))
    assert_close(res[0][0], P_o, RTOL, "sparse P long rows")
    assert_close(res[2][0], P_o, RTOL, "tiled P long rows")
    assert np.allclose(res[0][1]["S"], res[2][1]["S"], rtol=1e-12, atol=1e-30)


@pytest.mark.parametrize("N,K", [(40000, 9), (47000, 9), (60000, 9), (130000, 9), (61000, 17)])
def test_sparse_mstep_thread_budget(N, K, E):
    """Pools whose packed posteriors leave room for fewer than 1024 sparse-M-step threads (40k cells: 512; 47k: 256) or do not
    fit one shared-memory table at all (60k: two blocks of cells, 130k: five, 61k at K = 17: three -- one launch per block,
    each adding its share of every pseudo-row): six EM iterations against the tiled-only variant."""
    from scsplit_b200 import synth
    from scsplit_b200._lib import SCS_L, SCS_P
    V = 1200
    pool = synth.make_pool(V, N, K - 1, 0.05, 0.3, seed=N)
    lab = np.full(N, -1, dtype=np.int32)
    for k in range(K - 1):                                # a few true members per state: saturates within 3 iterations
        lab[np.flatnonzero((pool.truth == k) & ~pool.is_doublet)[:40]] = k
    lab[np.flatnonzero(pool.is_doublet)[:40]] = K - 1
    out = {}
    for variant in (0, 2):
        with E(pool.ref, pool.alt) as eng:
            eng.set_variant(variant)
            eng.em_begin(eng.seed_af(lab, K))
            eng.em_iterate(6, check_conv=False)
            out[variant] = (eng.get(SCS_P), eng.get(SCS_L), eng.em_status()[2][0], eng.stats()["dense_cells"])
    assert out[0][3] < N / 32, "the pool did not saturate: the sparse path was not exercised"
    assert_close(out[0][0], out[2][0], 1e-10, "P sparse vs tiled, N=%d" % N)
    assert_close(out[0][1], out[2][1], 1e-10, "L sparse vs tiled, N=%d" % N)
    assert_close(out[0][2], out[2][2], 1e-10, "LL sparse vs tiled, N=%d" % N)


@pytest.mark.parametrize("V,lam", [(6000, 0.1), (3000, 0.5), (20000, 0.01)])
def test_estep_k9_unit_lengths(V, lam, E):
    """K = 9 E-step across unit lengths: ~120 entries per (cell, tile) unit (packed groups, fewer units per order-pass warp),
    ~470 (too long for the order pass: the plain 8-lane form is chosen) and ~3 (mostly padding); three iterations vs the oracle."""
    from oracle import em_oracle as O
    from scsplit_b200 import synth
    from scsplit_b200._lib import SCS_L, SCS_P, SCS_W
    N, K = 384, 9
    pool = synth.make_pool(V, N, 8, 0.05, lam, seed=V)
    rng = np.random.default_rng(V)
    lab = np.full(N, -1, dtype=np.int32)
    pick = rng.choice(N, 3 * K, replace=False)
    lab[pick] = np.arange(3 * K) % K
    with E(pool.ref, pool.alt) as eng:
        k_alt = eng.kalt()
        P0 = eng.seed_af(lab, K)
        eng.em_begin(P0)
        eng.em_iterate(3, check_conv=False)
        want = O.run_em(pool.ref, pool.alt, P0, k_alt, fixed_iters=3)
        assert_close(eng.get(SCS_L), want["L"], RTOL, "L V=%d" % V)
        assert_close(eng.get(SCS_W), want["W"], 1e-7, "W V=%d" % V)
        assert_close(eng.get(SCS_P), want["P"], RTOL, "P V=%d" % V)


def test_config4_shard_step_against_oracle(E):
    """One E-step and one M-step at the size one of 8 GPUs holds of BASELINE config 4 (60,000 SNVs x 25,000 cells, K = 17:
    128-byte fixed-point rows, 768-thread sparse M-step) against the numpy oracle on the same shard."""
    from oracle import em_oracle as O
    from scsplit_b200 import synth
    from scsplit_b200._lib import SCS_L, SCS_LPS, SCS_P, SCS_W
    pool = synth.make_config("c4", cell_lo=50000, cell_hi=75000, workers=8)
    V, N = pool.ref.shape
    K = 17
    rng = np.random.default_rng(4)
    lab = np.full(N, -1, dtype=np.int32)
    lab[rng.choice(N, 3 * K, replace=False)] = np.arange(3 * K) % K
    k_alt = O.kalt(pool.ref, pool.alt)
    P0 = O.seed_af(pool.ref, pool.alt, k_alt, lab, K)
    lPs = np.array([np.log2(1 / K)] * K)
    with E(pool.ref, pool.alt) as eng:
        assert np.array_equal(eng.seed_af(lab, K), P0)
        eng.em_begin(P0)
        assert eng.em_info()["q_lanes"] == 8
        eng.estep()
        L, W = eng.get(SCS_L), eng.get(SCS_W)
        L_o = O.loglik_matrix(pool.ref, pool.alt, P0)
        assert_close(L, L_o, RTOL, "L, config-4 shard")
        assert_close(W, O.posterior(L, lPs), RTOL, "W(L_gpu), config-4 shard")
        assert_close(eng.stats()["ll"], O.total_loglik(L_o, lPs)[0], RTOL, "LL, config-4 shard")
        eng.mstep()
        P_o, lPs_o = O.mstep(pool.ref, pool.alt, W, k_alt)
        assert_close(eng.get(SCS_P), P_o, RTOL, "P, config-4 shard")
        assert_close(eng.get(SCS_LPS), lPs_o, RTOL, "lP_s, config-4 shard")
        # two more iterations so that the saturated-posterior (sparse) M-step also runs at this shape
        eng.em_iterate(3, check_conv=False)               # (continues from the state after the step-level E/M above)
        it, cv, ll = eng.em_status()
    Lq, Wq, LLq = O.estep(pool.ref, pool.alt, P_o, lPs_o)
    P2, lPs2 = O.mstep(pool.ref, pool.alt, Wq, k_alt)
    for _ in range(2):
        Lq, Wq, LLq = O.estep(pool.ref, pool.alt, P2, lPs2)
        P2, lPs2 = O.mstep(pool.ref, pool.alt, Wq, k_alt)
    assert_close(ll[0], LLq, RTOL, "LL after 3 more iterations, config-4 shard")
