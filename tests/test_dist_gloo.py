"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard plans, the packed statistics buffer that is
all-reduced per M-step, the unique-id broadcast and the best-restart selection.  The arithmetic on each fake rank is
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
