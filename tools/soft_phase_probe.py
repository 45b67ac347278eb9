"""Device time of the first iterations of an EM run (posteriors still soft) and the per-iteration count of unsaturated
cells (the sparse M-step takes over below N/32 of them; the gate was chosen with a build that made the divisor a knob)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from scsplit_b200.engine import Engine
cfg, pool = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
K = cfg["K"]; N = pool.ref.shape[1]
with Engine(pool.ref, pool.alt) as eng:
    for seed in (0, 2):
        lab = bench.random_seed_labels(N, K, seed)
        P0 = eng.seed_af(lab, K)
        for rep in range(2):
            eng.em_begin(P0, max_iter=1000)
            ts, ds = [], []
            for it in range(10):
                eng.em_iterate(1, check_conv=False)
                ts.append(eng.last_ms() * 1000); ds.append(int(eng.stats()["dense_cells"]))
            eng.em_end()
        print("seed", seed, "first 10 iterations: %.0f us" % sum(ts),
              "per-iter", [int(t) for t in ts], "dense", ds, flush=True)
