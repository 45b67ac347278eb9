#!/usr/bin/env python3
"""BASELINE config 5: restart sweep -- R random-initialisation EM runs at 20k x 10k, K = 8, batched as the grid's y
dimension and (under torchrun) sharded over the GPUs with no collective on the data path.

  python tools/restart_sweep.py [--restarts 256] [--chunk 64] [--init reference|random] [--check-winner]
  python -m torch.distributed.run --nproc-per-node 8 ... tools/restart_sweep.py --restarts 256

--init reference (default): the 256 initialisations come from ONE np.random.seed(2005) followed by 256 serial calls of the
reference's initialiser (coverage trimming -> PCA -> KMeans, scSplit:106-142), drawn on rank 0.
--check-winner re-runs the chosen restart on the numpy oracle from the same seed P and compares LL and assignments.

Every restart runs to the reference's own stop rule (bitwise-equal LL, cap 300; scSplit:156) on the device; the host
only polls the per-restart flags.  Prints one JSON line: wall time, total restart-iterations, restart-iterations/s and
the winner by the reference's "first strictly greater LL" rule (scSplit:482).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--restarts", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--init", default="reference", choices=["reference", "random"])
    ap.add_argument("--check-winner", action="store_true")
    args = ap.parse_args()
    import torch
    from scsplit_b200 import dist as D, synth
    from scsplit_b200.engine import Engine
    import bench
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg, pool = bench.workload(args.workload)
    K, (V, N) = cfg["K"], pool.ref.shape
    mine = D.shard_restarts(args.restarts, world)[rank]
    eng = Engine(pool.ref, pool.alt, device=local)
    t_init = time.perf_counter()
    if args.init == "random":
        all_labels = np.stack([bench.random_seed_labels(N, K, 5000 + i) for i in range(args.restarts)])
    else:
        import pandas as pd
        import tempfile
        from scsplit_b200 import models as M
        all_labels = np.full((args.restarts, N), -1, dtype=np.int32)
        if rank == 0:
            mtx = [pool.ref, pool.alt, pd.Index(pool.snv_ids), pd.Index(pool.barcodes)]
            out = tempfile.mkdtemp(prefix="sweep_")
            np.random.seed(2005)
            for i in range(args.restarts):
                lab, _ = M._initialise(eng, mtx, list(pool.barcodes), K, out)
                if lab is not None:
                    all_labels[i] = lab
        if world > 1:
            t = torch.from_numpy(all_labels).cuda()
            dist.broadcast(t, 0)
            all_labels = t.cpu().numpy()
    init_s = time.perf_counter() - t_init
    labels = all_labels[mine]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lls = np.full(args.restarts, np.nan)
    iters_total = 0
    dev_ms = 0.0
    for c0 in range(0, len(mine), args.chunk):
        part = mine[c0:c0 + args.chunk]
        P0 = eng.seed_af(labels[c0:c0 + len(part)], K)
        eng.em_begin(P0)
        eng.em_run()
        dev_ms += eng.last_ms()
        it, cv, ll = eng.em_status()
        iters_total += int(it.sum())
        lls[part] = ll
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    stats = torch.tensor([wall, dev_ms / 1e3, float(iters_total)], dtype=torch.float64, device="cuda")
    ll_t = torch.from_numpy(np.nan_to_num(lls, nan=-np.inf)).cuda()
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dist.all_reduce(ll_t, op=dist.ReduceOp.MAX)
        wall, dev_s, iters_total = float(mx[0]), float(mx[1]), int(sm[2])
    else:
        dev_s = dev_ms / 1e3
    if rank == 0:
        ll_all = [None if not np.isfinite(x) else float(x) for x in ll_t.cpu().numpy()]
        best = D.select_best(ll_all)
        check = None
        if args.check_winner and best is not None:
            from oracle import em_oracle as O
            from scsplit_b200._lib import SCS_W
            k_alt = O.kalt(pool.ref, pool.alt)
            P0 = O.seed_af(pool.ref, pool.alt, k_alt, all_labels[best], K)
            want = O.run_em(pool.ref, pool.alt, P0, k_alt)
            eng.em_begin(eng.seed_af(all_labels[best], K))
            eng.em_run()
            W = eng.get(SCS_W)
            with np.errstate(invalid="ignore"):
                check = {"oracle_ll": float(want["LL"]), "ll_rel_err": abs(ll_all[best] - want["LL"]) / max(abs(want["LL"]), 1e-300),
                         "same_assignments": bool(np.array_equal(W >= 0.99, want["W"] >= 0.99)),
                         "assigned_cells": int((W >= 0.99).any(axis=1).sum())}
        print(json.dumps({"init": args.init, "init_wall_s": init_s, "degenerate_ll_zero": int(sum(1 for x in ll_all if x == 0.0)),
                          "winner_vs_numpy_oracle": check, "workload": "%s: %d SNVs x %d cells, K=%d, %d restarts to convergence" % (args.workload, V, N, K, args.restarts),
                          "n_gpus": world, "wall_s": wall, "device_s": dev_s, "restart_iterations": iters_total,
                          "restart_iterations_per_s": iters_total / wall, "mean_iterations_per_restart": iters_total / args.restarts,
                          "chosen_restart": best, "best_ll": ll_all[best] if best is not None else None,
                          "distinct_final_ll": len(set(x for x in ll_all if x is not None))}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
