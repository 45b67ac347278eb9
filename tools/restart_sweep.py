#!/usr/bin/env python3
"""BASELINE config 5: restart sweep -- R random-initialisation EM runs at 20k x 10k, K = 8, batched as the grid's y
dimension and (under torchrun) sharded over the GPUs with no collective on the data path.

  python tools/restart_sweep.py [--restarts 256] [--chunk 64]
  python -m torch.distributed.run --nproc-per-node 8 ... tools/restart_sweep.py --restarts 256

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
    labels = np.stack([bench.random_seed_labels(N, K, 5000 + i) for i in mine])
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
        print(json.dumps({"workload": "%s: %d SNVs x %d cells, K=%d, %d restarts to convergence" % (args.workload, V, N, K, args.restarts),
                          "n_gpus": world, "wall_s": wall, "device_s": dev_s, "restart_iterations": iters_total,
                          "restart_iterations_per_s": iters_total / wall, "mean_iterations_per_restart": iters_total / args.restarts,
                          "chosen_restart": best, "best_ll": ll_all[best] if best is not None else None,
                          "distinct_final_ll": len(set(x for x in ll_all if x is not None))}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
