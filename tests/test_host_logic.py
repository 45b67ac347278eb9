"""CPU tests of the host-side logic: initialiser restatement, variant selection, C-ABI surface, CSV formats."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT, load_golden, scipy_count_fn, split_lists
from oracle import em_oracle as O


@pytest.mark.parametrize("name", ["c1_k2_d0", "k5_d025", "k3_ragged", "k4_dnone"])
def test_sparse_trim_matches_dense_oracle(name):
    """initialiser.trim_subset (counts only) selects exactly the rows/cells of the dense loop (scSplit:106-125)."""
    from scsplit_b200 import initialiser
    g = load_golden(name)
    ref, alt, K = g["ref"].astype(np.int64), g["alt"].astype(np.int64), g["K"]
    for seed in (int(g["seed"]), 11, 12):
        np.random.seed(seed)
        want = O.trim_subset(ref, alt, K)
        state_after = np.random.get_state()[1].copy()
        np.random.seed(seed)
        got = initialiser.trim_subset(scipy_count_fn(ref, alt), ref.shape[0], ref.shape[1], K)
        assert np.array_equal(np.random.get_state()[1], state_after)      # same number of beta draws consumed
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        assert got[2:] == want[2:]
        lab_w = O.initial_labels(ref, alt, *want, K)
        lab_g = initialiser.initial_labels(ref, alt, *got, K)
        assert np.array_equal(lab_w, lab_g)


def test_select_distinguishing_matches_reference_golden():
    from scsplit_b200.models import select_distinguishing
    for name in ("c1_k2_d0", "k5_d025", "k4_dnone", "k6_dead", "k3_ragged"):
        g = load_golden(name)
        pa = pd.DataFrame(g["pa_matrix"], index=g["pa_index"].tolist(), columns=g["pa_columns"].tolist())
        K = g["K"]
        doublet = int(g["doublet"])
        ncols = K - 1 + (doublet < 0) * 1
        cwd = os.getcwd()
        got = select_distinguishing(pa, K, ncols)
        assert sorted(set(got)) == g["dist_variants"].tolist()
        if os.path.exists(os.path.join(cwd, "scSplit_dist_variants.txt")):
            os.remove(os.path.join(cwd, "scSplit_dist_variants.txt"))


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly the functions include/scsplit_b200.h declares."""
    from scsplit_b200 import _lib
    with open(os.path.join(ROOT, "include", "scsplit_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(scs_[a-z0-9_]+)\s*\(", header, re.M))
    assert len(declared) >= 25
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.scs_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product refuses to run instead of silently computing on the host."""
    from scsplit_b200 import _lib
    from scsplit_b200.engine import Engine
    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    g = load_golden("k3_ragged")
    with pytest.raises(_lib.ScsError):
        Engine(g["ref"], g["alt"])


def test_product_never_imports_oracle():
    import ast
    pkg = os.path.join(ROOT, "scsplit_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            tree = ast.parse(open(os.path.join(pkg, fn)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                assert not any(n.split(".")[0] == "oracle" for n in names), fn


def test_synth_is_shard_consistent():
    """Cells [lo, hi) generated alone equal the same columns of the full pool (what cell-sharded ranks rely on)."""
    from scsplit_b200 import synth
    full = synth.make_pool(300, 2500, 3, 0.05, 0.05, seed=5)
    part = synth.make_pool(300, 2500, 3, 0.05, 0.05, seed=5, cell_lo=700, cell_hi=1900)
    assert (full.ref[:, 700:1900] != part.ref).nnz == 0
    assert (full.alt[:, 700:1900] != part.alt).nnz == 0
    assert np.array_equal(full.truth[700:1900], part.truth)
    dens = (full.ref + full.alt).nnz / (300 * 2500)
    assert abs(dens - (1 - np.exp(-0.05))) < 0.01


def test_csv_parser_matches_pandas(tmp_path):
    """csrc/csv.cu against pd.read_csv + csr_matrix (scSplit:544-546) on the reference's own file format."""
    from scipy.sparse import csr_matrix
    from scsplit_b200 import cli, synth
    pool = synth.make_pool(700, 257, 3, 0.05, 0.3, seed=3)
    rp, ap = str(tmp_path / "ref.csv"), str(tmp_path / "alt.csv")
    synth.write_csv(pool, rp, ap)
    for path in (rp, ap):
        frame = pd.read_csv(path, header=0, index_col=0)
        want = csr_matrix(frame.values)
        for threads in (1, 3):
            got, idx, cols = cli.read_count_csv(path, threads)
            assert got.shape == want.shape and (got != want).nnz == 0
            assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
            assert idx.tolist() == frame.index.tolist() and cols.tolist() == frame.columns.tolist()
    # a bigger file exercises the multi-threaded row split (> 1 MB), CRLF line ends and blanks around fields
    big = synth.make_pool(3000, 900, 3, 0.0, 0.2, seed=4)
    bp = str(tmp_path / "big.csv")
    synth.write_csv(big, bp, str(tmp_path / "big_alt.csv"))
    txt = open(bp).read().replace("\n", "\r\n").replace(",1,", ", 1 ,", 50)
    open(bp, "w", newline="").write(txt)
    got, idx, cols = cli.read_count_csv(bp, 4)
    assert (got != big.ref).nnz == 0 and idx.tolist() == big.snv_ids and cols.tolist() == big.barcodes
    # declined inputs fall back to pandas (floats), malformed ones raise
    fp = str(tmp_path / "float.csv")
    open(fp, "w").write("SNV,a,b\n1:5,1.0,0\n1:9,0,2\n")
    got, idx, cols = cli.read_count_csv(fp)
    assert got.toarray().tolist() == [[1.0, 0.0], [0.0, 2.0]]
    rg = str(tmp_path / "ragged.csv")
    open(rg, "w").write("SNV,a,b\n1:5,1\n")
    with pytest.raises(Exception):
        cli.read_count_csv(rg)


def test_positions_to_drop_equals_np_unique():
    """initialiser._positions_to_drop (scSplit:113-114) returns what np.unique(int(beta * n)) returns, from the same draws."""
    from scsplit_b200 import initialiser as I
    for n in (0, 4, 50, 3000, 30000):
        r1, r2 = np.random.RandomState(n + 1), np.random.RandomState(n + 1)
        got = I._positions_to_drop(r1, n)
        want = np.unique((r2.beta(1, 10, int(0.1 * n + 0.5)) * n).astype(np.int64))
        assert np.array_equal(got, want)
        assert r1.rand() == r2.rand()                  # same number of variates consumed


def test_synthetic_pool_is_the_same_from_threads():
    from scsplit_b200 import synth
    a = synth.make_pool(700, 2500, 3, 0.05, 0.05, seed=11)
    b = synth.make_pool(700, 2500, 3, 0.05, 0.05, seed=11, workers=3)
    assert (a.ref != b.ref).nnz == 0 and (a.alt != b.alt).nnz == 0 and np.array_equal(a.truth, b.truth)
    c = synth.make_pool(700, 2500, 3, 0.05, 0.05, seed=11, cell_lo=1000, cell_hi=2100, workers=2)
    assert (a.ref[:, 1000:2100] != c.ref).nnz == 0 and (a.alt[:, 1000:2100] != c.alt).nnz == 0


def test_fixed_point_format_helpers():
    """Host mirror of the fixed-point E-step's format arithmetic (csrc/estep_q.cuh): fraction bits from the deepest cell, the
    byte-pair packing of the extra value, and the wrap-around recovery of the sums."""
    def frac_bits(depth):
        b = 0
        while (1 << b) <= depth and b < 62:
            b += 1
        return min(49, 63 - 6 - b)
    assert frac_bits(1600) == 46 and frac_bits(3100) == 45 and frac_bits(1) == 49 and frac_bits((1 << 17)) == 39
    rng = np.random.default_rng(0)
    F, n = 46, 1500
    x = -rng.uniform(0, 40, size=(n, 9))
    q = np.rint(-x * 2.0 ** F).astype(np.uint64)
    assert (q < (1 << 55)).all()
    cnt = rng.integers(1, 9, size=n).astype(np.uint64)
    # lane j: q0 = a | pa << 56, q1 = b | pb << 56 with pa/pb bytes 2j, 2j+1 of the ninth value; all sums mod 2^64
    tot = np.zeros(9, dtype=object)
    for j in range(4):
        a, b, ex = q[:, 2 * j], q[:, 2 * j + 1], q[:, 8]
        pa, pb = (ex >> np.uint64(16 * j)) & np.uint64(255), (ex >> np.uint64(16 * j + 8)) & np.uint64(255)
        s0 = int(sum(int(c) * (int(v) | (int(p) << 56)) for c, v, p in zip(cnt, a, pa))) % (1 << 64)
        s1 = int(sum(int(c) * (int(v) | (int(p) << 56)) for c, v, p in zip(cnt, b, pb))) % (1 << 64)
        spa, spb = int(sum(int(c) * int(p) for c, p in zip(cnt, pa))), int(sum(int(c) * int(p) for c, p in zip(cnt, pb)))
        assert spa < (1 << 24)
        tot[2 * j] = (s0 - (spa << 56)) % (1 << 64)
        tot[2 * j + 1] = (s1 - (spb << 56)) % (1 << 64)
        tot[8] = tot[8] + (spa << (16 * j)) + (spb << (16 * j + 8))
    want = [int(sum(int(c) * int(v) for c, v in zip(cnt, q[:, k]))) for k in range(9)]
    assert [int(t) for t in tot] == want
    L = -np.array(want, dtype=np.float64) * 2.0 ** -F
    exact = (cnt[:, None].astype(np.float64) * x).sum(axis=0)
    assert np.allclose(L, exact, rtol=1e-12, atol=0)


def _same_bytes(frame, tmp_path, float_format=None, native=True):
    from scsplit_b200 import cli
    a, b = str(tmp_path / "fast.csv"), str(tmp_path / "pandas.csv")
    assert cli.frame_to_csv(frame, a, float_format=float_format) == native
    frame.to_csv(b, float_format=float_format)
    got, want = open(a, "rb").read(), open(b, "rb").read()
    if got != want:
        for i, (x, y) in enumerate(zip(got.split(b"\n"), want.split(b"\n"))):
            assert x == y, "line %d: %r vs %r" % (i, x[:200], y[:200])
    assert got == want


def test_float_writer_equals_pandas_to_csv(tmp_path):
    """csrc/csv.cu `scs_csv_write_f64` against DataFrame.to_csv, byte for byte, in the two forms the reference uses:
    P_s_c.to_csv(path) (scSplit:596, shortest round-trip floats) and pa_matrix.to_csv(path, float_format='%.0f')
    (scSplit:599-600)."""
    rng = np.random.default_rng(11)
    # every exponent, random mantissas (denormals included), both signs
    bits = rng.integers(0, 2**63, 60000, dtype=np.uint64) | (rng.integers(0, 2, 60000).astype(np.uint64) << np.uint64(63))
    vals = bits.view(np.float64).copy()
    special = [0.0, -0.0, 1.0, -1.0, 0.5, 0.1, 0.99, 1 - 1e-13, 1e-4, 9.999e-5, 1e-5, 1.5e-5, 1e15, 1e16, 9999999999999998.0,
               1.2345678901234567e16, 123456789012345678.0, 1e22, 1e23, 1e100, 1e-100, 1.5e300, 5e-324, 2.2250738585072014e-308,
               1.7976931348623157e308, np.inf, -np.inf, np.nan, 3.0, 1234.5, 100.0, 1e-300, 0.3333333333333333, 2 / 3, 1e-40]
    vals[:len(special)] = special
    vals[100:200] = rng.integers(-10**9, 10**9, 100).astype(np.float64)              # integral values print as "x.0"
    vals[200:5000] = 2.0 ** rng.uniform(-60, 60, 4800)                               # the everyday range
    vals[5000:9000] = rng.random(4000)
    frame = pd.DataFrame(vals.reshape(-1, 6), index=["BC%06d-1" % i for i in range(10000)], columns=range(6))
    _same_bytes(frame, tmp_path)
    # saturated posteriors as `scSplit run` writes them, index named, string column labels
    W = np.full((3000, 9), 1e-250) * rng.random((3000, 9))
    W[np.arange(3000), rng.integers(0, 9, 3000)] = 1 - rng.integers(0, 50, 3000) * 2.0 ** -53
    W[5] = np.nan
    named = pd.DataFrame(W, index=pd.Index(["c%d" % i for i in range(3000)], name="Barcode"), columns=[str(c) for c in range(9)])
    _same_bytes(named, tmp_path)
    # the ternary P/A matrix: 0, 1, NaN with '%.0f'; then arbitrary values through the same format
    pa = rng.choice([0.0, 1.0, np.nan], size=(4000, 8), p=[0.5, 0.4, 0.1])
    snv = ["%d:%d" % (1 + i % 22, 1000 + i) for i in range(4000)]
    _same_bytes(pd.DataFrame(pa, index=snv, columns=[0, 1, 2, 3, 5, 6, 7, 8]), tmp_path, "%.0f")
    odd = np.array([0.5, 1.5, 2.5, -0.4, -0.5, -0.0, 1e20, 123456.789, np.inf, -np.inf, np.nan, 1e300, 0.49999999999999994, 7.0, -3.5, 1e-9])
    _same_bytes(pd.DataFrame(odd.reshape(4, 4), index=list("abcd"), columns=range(4)), tmp_path, "%.0f")
    # empty frames (no distinguishing variants found) and frames the writer leaves to pandas
    _same_bytes(pd.DataFrame(pa[:0], index=[], columns=range(8)), tmp_path, "%.0f")
    _same_bytes(pd.DataFrame(pa[:0], index=pd.Index([], dtype=object), columns=range(8)), tmp_path)
    _same_bytes(pd.DataFrame(pa[:3], index=['a,b', 'c"d', 'e'], columns=range(8)), tmp_path, "%.0f", native=False)
    _same_bytes(pd.DataFrame(pa[:3], index=[1, 2, 3], columns=range(8)), tmp_path, native=False)
    _same_bytes(pd.DataFrame({"x": [1, 2], "y": [0.5, np.nan]}, index=["a", "b"]), tmp_path, native=False)
    _same_bytes(pd.DataFrame(pa[:2], index=["a", np.nan], columns=range(8)), tmp_path, native=False)
    _same_bytes(pd.DataFrame(pa[:2], index=["a", "b"], columns=range(8)), tmp_path, "%.3f", native=False)


def test_pattern_representatives_equal_the_text_grouping(monkeypatch):
    """models.pattern_representatives (integer patterns, one sort) against the reference's grouping by concatenated value
    text with a label look-up per group (scSplit:291-296, kept as models._representatives_by_text): same representatives on
    random ternary matrices, and the same distinguishing variants from the whole selection with either of them."""
    from scsplit_b200 import models as M
    rng = np.random.default_rng(5)
    cwd_note = os.path.join(os.getcwd(), "scSplit_dist_variants.txt")
    for trial in range(12):
        rows, cols = int(rng.integers(40, 1500)), int(rng.integers(2, 9))
        p_nan = float(rng.choice([0.0, 0.05, 0.3]))
        vals = rng.choice([0.0, 1.0, np.nan], size=(rows, cols), p=[(1 - p_nan) * 0.6, (1 - p_nan) * 0.4, p_nan])
        frame = pd.DataFrame(vals, index=["%d:%d" % (1 + i % 22, 1000 + i) for i in range(rows)], columns=range(cols))
        for width in range(cols, 1, -1):
            sub = frame.iloc[:, 0:width]
            informative = ((sub.var(axis=1) > 0) & (sub.count(axis=1) == width)).values
            if not informative.any():
                continue
            assert sorted(M.pattern_representatives(frame, sub, informative)) == sorted(M._representatives_by_text(frame, sub, informative))
        fast = M.select_distinguishing(frame, cols + 1, cols)
        with monkeypatch.context() as mp:
            mp.setattr(M, "pattern_representatives", M._representatives_by_text)
            slow = M.select_distinguishing(frame, cols + 1, cols)
        assert fast == slow
    # repeated labels and values other than 0 / 1 take the text grouping itself
    dup = pd.DataFrame(rng.choice([0.0, 1.0], size=(30, 3)), index=["1:%d" % (i % 20) for i in range(30)], columns=range(3))
    inf = (dup.var(axis=1) > 0).values
    assert sorted(map(str, M.pattern_representatives(dup, dup, inf))) == sorted(map(str, M._representatives_by_text(dup, dup, inf)))
    if os.path.exists(cwd_note):
        os.remove(cwd_note)


def test_csv_parser_block_path_on_random_files(tmp_path):
    """The 16-byte block path of csrc/csv.cu (eight one-digit fields per step) against pd.read_csv on small random files:
    every row width from 1 to 40 columns (line ends at every offset inside a block), counts of several digits, leading
    zeros, blanks, CRLF, no newline at the end; malformed rows are refused, not misread."""
    from scipy.sparse import csr_matrix
    from scsplit_b200 import cli
    from scsplit_b200._lib import ScsError
    rng = np.random.default_rng(21)
    for trial in range(60):
        ncols, nrows = int(rng.integers(1, 41)), int(rng.integers(1, 30))
        vals = rng.choice([0, 0, 0, 0, 0, 0, 0, 1, 2, 9, 10, 37, 2000, 123456], size=(nrows, ncols))
        if trial % 5 == 0:
            vals[:] = rng.integers(0, 10, size=vals.shape) * (rng.random(vals.shape) < 0.3)      # one-digit fields only
        lines = ["SNV," + ",".join("c%d" % j for j in range(ncols))]
        for i in range(nrows):
            fields = [str(int(v)) for v in vals[i]]
            if trial % 7 == 3:
                fields = [("0" + f if rng.random() < 0.2 else f) for f in fields]                  # "00", "07"
            if trial % 7 == 5:
                fields = [(" " + f + " " if rng.random() < 0.2 else f) for f in fields]
            lines.append("%d:%d," % (1 + i % 22, 1000 + i) + ",".join(fields))
        eol = "\r\n" if trial % 4 == 1 else "\n"
        text = eol.join(lines) + (eol if trial % 3 else "")
        path = str(tmp_path / ("r%d.csv" % trial))
        open(path, "w", newline="").write(text)
        got, idx, cols = cli.read_count_csv(path, 1 + trial % 3)
        assert got.shape == (nrows, ncols) and np.array_equal(got.toarray(), vals), trial
        assert got.has_canonical_format and (got.data != 0).all()
        assert idx.tolist() == ["%d:%d" % (1 + i % 22, 1000 + i) for i in range(nrows)] and cols.tolist() == ["c%d" % j for j in range(ncols)]
    # malformed rows: a field too many / too few (also far inside a run of zeros), an empty field, a trailing comma
    head = "SNV," + ",".join("c%d" % j for j in range(20)) + "\n"
    good = "1:1," + ",".join(["0"] * 20) + "\n"
    for bad in ("1:2," + ",".join(["0"] * 21), "1:2," + ",".join(["0"] * 19), "1:2," + ",".join(["0"] * 10 + [""] + ["0"] * 9),
                "1:2," + ",".join(["0"] * 20) + ",", "1:2," + ",".join(["0"] * 19) + ","):
        path = str(tmp_path / "bad.csv")
        open(path, "w").write(head + good + bad + "\n" + good)
        try:
            got, _, _ = cli.read_count_csv(path)          # declined files go to pandas, which may accept an empty field as NaN
        except (ScsError, Exception):
            continue
        want = pd.read_csv(path, header=0, index_col=0)
        assert got.shape == want.shape and np.array_equal(np.nan_to_num(got.toarray(), nan=-1), np.nan_to_num(want.values.astype(float), nan=-1))
    # counts beyond int32 are refused, never wrapped
    open(str(tmp_path / "big.csv"), "w").write("SNV,a,b\n1:5,4294967297,0\n")
    with pytest.raises(Exception):
        cli.read_count_csv(str(tmp_path / "big.csv"))


def test_write_outputs_equals_the_pandas_writers(tmp_path):
    """cli.write_outputs against the reference's own writer calls (scSplit:581-600, restated with pandas here): the five
    result files byte for byte, with and without a doublet cluster, including an empty dist_matrix."""
    from scsplit_b200 import cli
    rng = np.random.default_rng(8)
    V, N, K = 400, 300, 5
    snv = ["%d:%d" % (1 + v % 22, 1000 + v) for v in range(V)]
    bc = ["BC%05d-1" % c for c in range(N)]

    class Model:
        pass

    for doublets, dbl in ((-1, 2), (0, -1), (0.1, 0)):
        m = Model()
        lab = rng.integers(0, K, N)
        W = np.full((N, K), 1e-200) * rng.random((N, K))
        W[np.arange(N), lab] = 1 - rng.integers(0, 9, N) * 2.0 ** -53
        W[7] = np.nan
        m.doublet = dbl
        m.P_s_c = pd.DataFrame(W, index=bc, columns=range(K))
        m.reassigned = [sorted(np.asarray(bc, dtype=object)[lab == k].tolist()) for k in range(K)]
        cols = [k for k in range(K) if k != dbl]
        pa = pd.DataFrame(rng.choice([0.0, 1.0, np.nan], size=(V, len(cols)), p=[0.5, 0.4, 0.1]), index=snv, columns=cols)
        m.pa_matrix = pa
        m.dist_variants = list(rng.choice(snv, 4, replace=False)) if doublets != 0 else []
        m.dist_matrix = pa.reindex(m.dist_variants)
        got_dir, want_dir = tmp_path / ("got%s" % dbl), tmp_path / ("want%s" % dbl)
        got_dir.mkdir()
        want_dir.mkdir()
        cli.write_outputs(m, K, doublets, str(got_dir))
        frames = [pd.DataFrame({"Barcode": m.reassigned[n], "Cluster": "SNG-" + str(n)}) for n in range(K) if n != dbl]
        if doublets != 0:
            frames.append(pd.DataFrame({"Barcode": m.reassigned[dbl], "Cluster": "DBL-" + str(dbl)}))
        pd.concat(frames).to_csv(want_dir / "scSplit_result.csv", sep="\t", index=False)
        m.P_s_c.to_csv(want_dir / "scSplit_P_s_c.csv")
        with open(want_dir / "scSplit_dist_variants.txt", "w") as f:
            for item in m.dist_variants:
                f.write(str(item) + "\n")
        m.dist_matrix.to_csv(want_dir / "scSplit_dist_matrix.csv", float_format="%.0f")
        m.pa_matrix.to_csv(want_dir / "scSplit_PA_matrix.csv", float_format="%.0f")
        for fn in ("scSplit_result.csv", "scSplit_P_s_c.csv", "scSplit_dist_variants.txt", "scSplit_dist_matrix.csv", "scSplit_PA_matrix.csv"):
            assert open(got_dir / fn, "rb").read() == open(want_dir / fn, "rb").read(), (fn, doublets)
