"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard plans, the packed statistics buffer that is
all-reduced per M-step, the unique-id broadcast, the best-restart selection, and the drop-in's restart driver
(models.core_batched) sharded over two ranks.  The arithmetic on each fake rank is
the CPU oracle's (test infrastructure); the product's ranks run the same plan with the CUDA kernels + NCCL."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import em_oracle as O
    from scsplit_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden(name)
        ref, alt, K = g["ref"].astype(np.float64), g["alt"].astype(np.float64), g["K"]
        V, N = ref.shape
        reads = np.asarray((ref + alt).sum(axis=0)).ravel()
        lo, hi = D.shard_ranges(N, world, weights=reads)[rank]
        # unique id travels from rank 0
        uid = D.broadcast_unique_id(lambda: np.arange(128, dtype=np.uint8)[::-1].copy(), rank)
        assert np.array_equal(uid, np.arange(128, dtype=np.uint8)[::-1])
        # k_alt needs GLOBAL per-SNV sums: all-reduce of integer sums (scs_comm_init does this on the device)
        sums = torch.from_numpy(np.stack([np.asarray(alt[:, lo:hi].sum(axis=1)).ravel(), np.asarray(ref[:, lo:hi].sum(axis=1)).ravel()]))
        dist.all_reduce(sums)
        k_alt = (sums[0].numpy() + 1) / ((sums[1].numpy() + 1) + (sums[0].numpy() + 1))
        assert np.array_equal(k_alt, g["k_alt"])
        P, lPs = g["P0"], np.array([np.log2(1 / K)] * K)
        for t in range(3):
            L, W, ll = O.estep(ref[:, lo:hi], alt[:, lo:hi], P, lPs)          # local cells only
            altW = np.asarray(alt[:, lo:hi].dot(W))
            refW = np.asarray(ref[:, lo:hi].dot(W))
            buf = torch.from_numpy(D.pack_stats(altW, refW, np.nansum(W, axis=0), ll))
            assert buf.numel() == D.stats_layout(V, K)["size"]
            dist.all_reduce(buf)                                              # the one exchange step per iteration
            a, r, colsum, ll_all = D.unpack_stats(buf.numpy(), V, K)
            P, lPs = D.finalize_from_stats(a, r, colsum, k_alt)
            np.save(os.path.join(out_dir, "P_%d_%d.npy" % (rank, t)), P)
            np.save(os.path.join(out_dir, "ll_%d_%d.npy" % (rank, t)), np.array([ll_all]))
            np.save(os.path.join(out_dir, "lps_%d_%d.npy" % (rank, t)), lPs)
        # restart sharding: LL of every restart gathered, reference selection rule applied identically everywhere
        mine = D.shard_restarts(7, world)[rank]
        lls = torch.full((7,), float("nan"), dtype=torch.float64)
        for i in mine:
            lls[i] = -100.0 - (i % 3)
        gathered = [torch.zeros(7, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, lls)
        merged = np.nanmax(np.stack([x.numpy() for x in gathered]), axis=0)
        assert D.select_best(merged.tolist()) == 0        # restarts 0, 3, 6 tie at -100: the FIRST wins
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["k5_d025", "k3_ragged"])
def test_cell_sharded_iteration_matches_unsharded(name, tmp_path):
    import torch.multiprocessing as mp
    from oracle import em_oracle as O
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    g = load_golden(name)
    K = g["K"]
    for t in range(3):
        P0 = np.load(tmp_path / ("P_0_%d.npy" % t))
        P1 = np.load(tmp_path / ("P_1_%d.npy" % t))
        assert np.array_equal(P0, P1)                     # every rank ends the iteration with identical bits
        want_P, want_lps, want_ll = g["P_%d" % t], g["lPs_%d" % t], float(g["LL_%d" % t])
        assert np.allclose(P0, want_P, rtol=1e-9, atol=0)
        assert np.allclose(np.load(tmp_path / ("lps_0_%d.npy" % t)), want_lps, rtol=1e-9, atol=0)
        assert abs(np.load(tmp_path / ("ll_0_%d.npy" % t))[0] - want_ll) <= 1e-9 * abs(want_ll)


def test_shard_plans():
    from scsplit_b200 import dist as D
    for n, w in ((10, 3), (7, 8), (200000, 8), (1, 1)):
        r = D.shard_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
    wts = np.r_[np.ones(50), 9 * np.ones(50)]
    r = D.shard_ranges(100, 2, wts)
    assert r[0][1] > 50 and r == [(0, r[0][1]), (r[0][1], 100)]
    assert sorted(sum(D.shard_restarts(30, 8), [])) == list(range(30))
    assert D.select_best([None, -5.0, -3.0, -3.0, float("nan")]) == 2
    assert D.select_best([None, None]) is None
    lay = D.stats_layout(30000, 9)
    assert lay["KP"] == 10 and lay["size"] == 2 * 30000 * 10 + 10 + 2        # 4.8 MB all-reduce payload at config 3


# ---- the drop-in's restart driver under a process group (scSplit:470-490 sharded over ranks) --------------------------------

class _OracleEngine:
    """Host stand-in for scsplit_b200.engine.Engine, for driving models.core_batched on CPU ranks: the same methods, the
    arithmetic is the CPU oracle's (test infrastructure; the product's ranks hold a CUDA engine each)."""

    def __init__(self, ref, alt):
        from conftest import scipy_count_fn
        from oracle import em_oracle as O
        self.O, self.ref, self.alt = O, ref, alt
        self.V, self.N = ref.shape
        self.h = 1
        self._counts = scipy_count_fn(ref, alt)
        self._kalt = O.kalt(ref, alt)
        self.begun = []                                   # restarts per em_begin call, for the test to look at

    def close(self):
        self.h = None

    def kalt(self):
        return self._kalt

    def pattern_counts(self, keep_rows=None, keep_cols=None):
        return self._counts(keep_rows, keep_cols)

    def seed_af(self, labels, K):
        labels = np.asarray(labels)
        if labels.ndim == 1:
            return self.O.seed_af(self.ref, self.alt, self._kalt, labels, K)
        return np.stack([self.O.seed_af(self.ref, self.alt, self._kalt, lab, K) for lab in labels])

    def em_footprint(self, K):
        return 1, 10 ** 12

    def em_begin(self, P0):
        self.P0 = P0 if P0.ndim == 3 else P0[None]
        self.begun.append(len(self.P0))

    def em_run(self):
        self.res = [self.O.run_em(self.ref, self.alt, p, self._kalt) for p in self.P0]

    def em_status(self):
        return (np.array([r["iters"] for r in self.res], dtype=np.int32), np.array([r["convergence"] for r in self.res], dtype=np.int32),
                np.array([r["LL"] for r in self.res], dtype=np.float64))

    def get(self, what, restart=0):
        r = self.res[restart]
        if what == 4:
            return np.asarray(r["hist"] + [0.0] * (300 - len(r["hist"])), dtype=np.float64)
        return np.asarray({0: r["P"], 1: r["W"], 2: r["L"], 3: r["lPs"]}[what], dtype=np.float64)


def _run_core(name, ems, out_dir):
    """models.core_batched on the golden pool `name` with the oracle engine; returns what rank 0 (or the only process) got."""
    import pandas as pd
    from scsplit_b200 import models as M
    g = load_golden(name)
    mtx = [g["ref"], g["alt"], pd.Index(["1:%d" % (1000 + v) for v in range(g["ref"].shape[0])]),
           pd.Index(["BC%05d-1" % c for c in range(g["ref"].shape[1])])]
    eng = _OracleEngine(g["ref"], g["alt"])
    real = M.get_engine
    M.get_engine = lambda base_calls_mtx, device=0: eng
    try:
        np.random.seed(int(g["seed"]))
        model, best = M.core_batched(mtx, g["K"], out_dir, ems)
    finally:
        M.get_engine = real
    return model, best, eng


def _core_worker(rank, world, port, name, ems, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rank_dir = os.path.join(out_dir, "rank%d" % rank)
        os.makedirs(rank_dir)
        model, best, eng = _run_core(name, ems, rank_dir)
        assert sum(eng.begun) == len(range(rank, ems, world))          # this rank ran its share of the restarts, nothing else
        np.save(os.path.join(out_dir, "best_%d.npy" % rank), np.array([best]))
        if rank == 0:
            np.savez(os.path.join(out_dir, "rank0.npz"), W=model.P_s_c.values, P=model.model_af.values, L=model.lP_c_s.values,
                     lP_c_m=model.lP_c_m, assigned=np.array([len(a) for a in model.assigned]),
                     first=np.array([a[0] if a else "" for a in model.assigned], dtype=str),
                     initial=np.array([",".join(i) for i in model.initial], dtype=str))
            assert os.path.exists(os.path.join(rank_dir, "scSplit.log"))
        else:
            assert model is None                                         # only rank 0 returns the populated model
            assert not os.path.exists(os.path.join(rank_dir, "scSplit.log"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,ems", [("k4_dnone", 5), ("k3_ragged", 4)])
def test_restart_driver_sharded_over_two_ranks_equals_one_process(name, ems, tmp_path):
    """models.core_batched under torchrun semantics (world 2, gloo): rank 0 draws every initialisation and broadcasts the
    labels, the ranks run restarts r, r + 2, ..., the reference's selection rule (scSplit:482-484) runs on the gathered LLs,
    the winner and the last restart's arrays travel to rank 0 -- whose model must equal, bit for bit, the one a single
    process builds from the same seed."""
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_core_worker, args=(2, port, name, ems, str(tmp_path)), nprocs=2, join=True)
    solo_dir = tmp_path / "solo"
    solo_dir.mkdir()
    model, best, eng = _run_core(name, ems, str(solo_dir))
    assert sum(eng.begun) == ems
    got = np.load(tmp_path / "rank0.npz")
    assert int(np.load(tmp_path / "best_0.npy")[0]) == best == int(np.load(tmp_path / "best_1.npy")[0])
    assert np.array_equal(got["W"], model.P_s_c.values, equal_nan=True)
    assert np.array_equal(got["P"], model.model_af.values, equal_nan=True)
    assert np.array_equal(got["L"], model.lP_c_s.values, equal_nan=True)
    assert float(got["lP_c_m"]) == model.lP_c_m
    assert got["assigned"].tolist() == [len(a) for a in model.assigned]
    assert got["first"].tolist() == [a[0] if a else "" for a in model.assigned]
    assert got["initial"].tolist() == [",".join(i) for i in model.initial]
    # the log of rank 0 lists every restart and every log-likelihood, like the single process's
    log2, log1 = open(tmp_path / "rank0" / "scSplit.log").read(), open(solo_dir / "scSplit.log").read()
    assert log2.count("Model Log-Likelihood = ") == log1.count("Model Log-Likelihood = ") == ems
    assert [l for l in log2.splitlines() if l.startswith("Model Log")] == [l for l in log1.splitlines() if l.startswith("Model Log")]


def test_trim_loops_ahead_of_the_clustering_keep_the_reference_stream(tmp_path, monkeypatch):
    """models._initialise_all runs the trimming loops of all restarts before any block is clustered.  Same labels and the
    same final `np.random` state as the reference's alternating order (trim, cluster, trim, cluster, ...: scSplit:473-477
    with 106-135) -- also when a clustering step itself draws from `np.random` (PCA's randomised solver would), which must
    send the later restarts back to the alternating order."""
    import pandas as pd
    from scsplit_b200 import initialiser, models as M
    g = load_golden("k4_dnone")
    ref, alt, K = g["ref"], g["alt"], g["K"]
    barcodes = ["BC%05d-1" % c for c in range(ref.shape[1])]
    mtx = [ref, alt, pd.Index(["1:%d" % (1000 + v) for v in range(ref.shape[0])]), pd.Index(barcodes)]
    eng = _OracleEngine(ref, alt)
    monkeypatch.setattr(M, "get_engine", lambda base_calls_mtx, device=0: eng)
    real_labels = initialiser.initial_labels
    for draws_in in ((), (1,), (0, 3)):                       # restarts whose clustering consumes random numbers
        calls = []

        def noisy(*a, **k):
            if len(calls) in draws_in:
                np.random.random(3)
            calls.append(1)
            return real_labels(*a, **k)

        monkeypatch.setattr(initialiser, "initial_labels", noisy)
        np.random.seed(99)
        want = [M._initialise(eng, mtx, barcodes, K, str(tmp_path)) for _ in range(5)]      # the alternating order
        want_state = np.random.get_state()
        del calls[:]
        np.random.seed(99)
        model, labels, initial = M._initialise_all(eng, mtx, barcodes, K, str(tmp_path), 5, 0)
        assert M._same_rng_state(np.random.get_state(), want_state), draws_in
        for i in range(5):
            assert np.array_equal(labels[i], want[i][0]) and initial[i] == want[i][1], (draws_in, i)
        assert np.array_equal(model._labels, want[4][0]) and model.initial == want[4][1]
    log = open(tmp_path / "scSplit.log").read()
    assert log.count("Repeat 5 of 5 in demultiplexing") == 3 and log.count("Building model... ") == 15
