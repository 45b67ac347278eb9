"""Iteration time and saturation state at config 3 for the eight random initialisations `bench.py --gpus 8` gives its ranks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from scsplit_b200.engine import Engine
cfg, pool = bench.workload("c3")
K = cfg["K"]
N = pool.ref.shape[1]
with Engine(pool.ref, pool.alt) as eng:
    for seed in range(8):
        lab = bench.random_seed_labels(N, K, seed)
        P0 = eng.seed_af(lab, K)
        eng.em_begin(P0, max_iter=1000)
        eng.em_iterate(20, check_conv=False)
        eng.em_iterate(200, check_conv=False)
        st = eng.stats()
        print("seed", seed, "us/iter %.1f" % (eng.last_ms() * 5), "dense cells", st["dense_cells"], "colsum", np.round(st["colsum"]).astype(int).tolist(), flush=True)
        eng.em_end()
