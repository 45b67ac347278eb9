"""2-GPU test of the cell-sharded path: two processes, one GPU each, NCCL all-reduce of the M-step statistics inside
libscsplit_b200.so.  Skipped on a box with fewer than 2 GPUs (the single-GPU tests cover the kernels; the gloo tests in
test_dist_gloo.py cover the host logic)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from scsplit_b200 import dist as D
    from scsplit_b200._lib import SCS_LLHIST, SCS_P, SCS_W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g = load_golden(name)
        K, N = g["K"], int(g["N"])
        eng, (lo, hi) = D.open_cell_shard(g["ref"], g["alt"], rank, world, rank)
        assert np.array_equal(eng.kalt(), g["k_alt"])                      # global sums were all-reduced
        lab = np.full(N, -1, dtype=np.int32)
        from conftest import split_lists
        for n, lst in enumerate(split_lists(g["initial_flat"], g["initial_ptr"])):
            lab[lst] = n
        P0 = eng.seed_af(lab[lo:hi], K)
        assert np.array_equal(P0, g["P0"])                                 # integer-exact across the all-reduce
        eng.em_begin(P0)
        eng.em_run()
        it, cv, ll = eng.em_status()
        np.save(os.path.join(out_dir, "W_%d.npy" % rank), eng.get(SCS_W))
        np.save(os.path.join(out_dir, "P_%d.npy" % rank), eng.get(SCS_P))
        np.save(os.path.join(out_dir, "meta_%d.npy" % rank), np.array([it[0], cv[0], ll[0], lo, hi], dtype=np.float64))
        n_ref, n_alt, pa = eng.cluster_counts(np.arange(lo, hi, dtype=np.int32) % K, K)
        np.save(os.path.join(out_dir, "nref_%d.npy" % rank), n_ref)
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["k5_d025", "c1_k2_d0"])
def test_two_gpu_cell_sharded_em(name, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from oracle import em_oracle as O
    world = 2
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    g = load_golden(name)
    K = g["K"]
    m0, m1 = np.load(tmp_path / "meta_0.npy"), np.load(tmp_path / "meta_1.npy")
    assert m0[0] == m1[0] and m0[2] == m1[2]                               # same iteration count and LL bits on both ranks
    W = np.concatenate([np.load(tmp_path / "W_0.npy"), np.load(tmp_path / "W_1.npy")])
    assert np.array_equal(np.load(tmp_path / "P_0.npy"), np.load(tmp_path / "P_1.npy"))
    with np.errstate(invalid="ignore"):
        assert np.array_equal(W >= 0.99, g["W_final"] >= 0.99)             # identical assignments to the reference
    assert abs(m0[2] - float(g["LL_final"])) <= 1e-9 * abs(float(g["LL_final"]))
    assert np.allclose(np.load(tmp_path / "P_0.npy"), g["P_final"], rtol=1e-6, atol=0)
    lists = [np.flatnonzero(np.arange(int(g["N"])) % K == n).tolist() for n in range(K)]
    o_ref, _ = O.cluster_counts(g["ref"], g["alt"], lists, K)
    assert np.array_equal(np.load(tmp_path / "nref_0.npy"), o_ref)         # int64 all-reduce of the hard sums
