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


def _worker_batch(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from scsplit_b200 import dist as D, synth
    from scsplit_b200._lib import SCS_W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pool, labels, K = _batch_case()
        eng, (lo, hi) = D.open_cell_shard(pool.ref, pool.alt, rank, world, rank)
        P0 = eng.seed_af(labels[:, lo:hi], K)
        eng.em_begin(P0)
        eng.em_run()
        eng.em_iterate(24, check_conv=True)        # every restart has stopped: these rounds must not touch anything any more
        it, cv, ll = eng.em_status()
        np.save(os.path.join(out_dir, "bW_%d.npy" % rank), np.stack([eng.get(SCS_W, r) for r in range(len(labels))]))
        np.save(os.path.join(out_dir, "bS_%d.npy" % rank), np.stack([eng.stats(r)["S"] for r in range(len(labels))]))
        np.save(os.path.join(out_dir, "bmeta_%d.npy" % rank), np.stack([it, cv, ll]).astype(np.float64))
        # each restart alone on the same sharded pool: what the batch must reproduce bit for bit
        alone_S, alone_it = [], []
        for r in range(len(labels)):
            eng.em_begin(P0[r])
            eng.em_run()
            alone_it.append(eng.em_status()[0][0])
            alone_S.append(eng.stats(0)["S"])
        np.save(os.path.join(out_dir, "bS1_%d.npy" % rank), np.stack(alone_S))
        np.save(os.path.join(out_dir, "bit1_%d.npy" % rank), np.array(alone_it, dtype=np.float64))
        eng.close()
    finally:
        dist.destroy_process_group()


def _batch_case():
    from scsplit_b200 import synth
    K, N = 4, 1200
    pool = synth.make_pool(4000, N, 3, 0.05, 0.08, seed=41)
    rng = np.random.default_rng(5)
    labels = np.full((3, N), -1, dtype=np.int32)
    for r in range(3):
        labels[r, rng.choice(N, 3 * K, replace=False)] = np.arange(3 * K) % K
    return pool, labels, K


def test_two_gpu_restart_batch_statistics_survive_convergence(tmp_path):
    """Three restarts batched on a cell-sharded pool: restarts stop at different iterations; the summed statistics of a
    stopped restart must stay what they were (the collectives are out of place and leave stopped restarts out)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from scsplit_b200._lib import SCS_W
    from scsplit_b200.engine import Engine
    world = 2
    port = 31600 + (os.getpid() % 2000)
    mp.spawn(_worker_batch, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    pool, labels, K = _batch_case()
    with Engine(pool.ref, pool.alt) as eng:
        eng.em_begin(eng.seed_af(labels, K))
        eng.em_run()
        it, cv, ll = eng.em_status()
        W1 = np.stack([eng.get(SCS_W, r) for r in range(3)])
        S1 = np.stack([eng.stats(r)["S"] for r in range(3)])
    m0, m1 = np.load(tmp_path / "bmeta_0.npy"), np.load(tmp_path / "bmeta_1.npy")
    assert np.array_equal(m0, m1)
    W2 = np.concatenate([np.load(tmp_path / "bW_0.npy"), np.load(tmp_path / "bW_1.npy")], axis=1)
    S2 = np.load(tmp_path / "bS_0.npy")
    assert np.array_equal(S2, np.load(tmp_path / "bS_1.npy"))
    assert np.isfinite(S2).all()
    with np.errstate(invalid="ignore"):
        assert np.array_equal(W1 >= 0.99, W2 >= 0.99)
    assert np.allclose(m0[2], ll, rtol=1e-9, atol=0)
    # the statistics belong to each restart's own last M-step: a restart run ALONE on the same two shards stops at the same
    # iteration and leaves the same bits (the 24 extra rounds and the restarts batched next to it changed nothing)
    assert np.array_equal(np.load(tmp_path / "bit1_0.npy"), m0[0])
    assert np.array_equal(np.load(tmp_path / "bS1_0.npy"), S2)
    # against the single-GPU run: iteration counts may differ by the stop rule's last ulp (the LL is summed in another order),
    # so the tight comparison is made where both runs stopped at the same iteration, a loose one everywhere (both have converged)
    same = m0[0] == it
    print("restarts stopping at the same iteration on 1 and 2 GPUs: %d of 3 (%s vs %s)" % (int(same.sum()), m0[0], it))
    assert np.allclose(S2[same], S1[same], rtol=1e-9, atol=1e-12)
    assert np.allclose(S2, S1, rtol=1e-4, atol=1e-7)


def _worker_cli(rank, world, port, argv, seed):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from scsplit_b200 import cli
    np.random.seed(seed)               # only rank 0 draws initialisations; it broadcasts the labels
    cli.run(list(argv))


def test_two_gpu_cli_run_matches_reference_partition(tmp_path):
    """`scSplit run` launched one process per GPU (torchrun environment): restarts sharded over the ranks, the reference's
    selection rule applied to the gathered log-likelihoods, rank 0 writes the outputs.  Same partition as the reference's
    own result file for the same CSVs (golden cli_k4, scSplit:470-490 + 582-594)."""
    import ast
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import pandas as pd
    import torch.multiprocessing as mp
    from conftest import GOLDEN_DIR
    from scsplit_b200 import synth
    d = dict(np.load(os.path.join(GOLDEN_DIR, "cli_k4.npz")))
    c = ast.literal_eval(str(d["params"][0]))
    pool = synth.make_pool(c["V"], c["N"], c["samples"], c["dfrac"], c["lam"], c["pseed"])
    rp, ap = str(tmp_path / "ref_filtered.csv"), str(tmp_path / "alt_filtered.csv")
    synth.write_csv(pool, rp, ap)
    argv = ["-r", rp, "-a", ap, "-n", str(c["n"]), "-o", str(tmp_path), "-e", str(c["e"])]
    if c["d"] is not None:
        argv += ["-d", c["d"]]
    mp.spawn(_worker_cli, args=(2, 33600 + (os.getpid() % 2000), tuple(argv), int(c["seed"])), nprocs=2, join=True)
    res = pd.read_csv(tmp_path / "scSplit_result.csv", sep="\t")

    def partition(barcodes, clusters):
        out = {}
        for b, cl in zip(barcodes, clusters):
            out.setdefault(cl, []).append(b)
        return out
    got = partition(res["Barcode"].astype(str), res["Cluster"].astype(str))
    want = partition(d["result_barcode"].tolist(), d["result_cluster"].tolist())
    assert sorted(map(sorted, got.values())) == sorted(map(sorted, want.values()))
    log = open(tmp_path / "scSplit.log").read()
    assert log.count("Building model...") == c["e"] and log.count("Model Log-Likelihood = ") >= 1
