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
