#!/usr/bin/env python3
"""How many cells have an EXACTLY one-hot posterior row per EM iteration (decides whether a hard-count M-step pays)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scsplit_b200.engine import Engine
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg, pool = bench.workload(name)
K = cfg["K"]
with Engine(pool.ref, pool.alt) as eng:
    P0 = eng.seed_af(bench.random_seed_labels(pool.ref.shape[1], K, 0), K)
    eng.em_begin(P0)
    for it in range(14):
        eng.em_iterate(1, check_conv=False)
        W = eng.get(1)
        srt = np.sort(W, axis=1)
        onehot = (srt[:, -1] == 1.0) & (srt[:, -2] < 1e-40)
        one = (srt[:, -1] == 1.0)
        print("        max==1.0 exactly: %.4f ; second max < 1e-40 | < 1e-20 | < 1e-10: %.4f %.4f %.4f" % (one.mean(), (srt[:, -2] < 1e-40).mean(), (srt[:, -2] < 1e-20).mean(), (srt[:, -2] < 1e-10).mean()))
        near = (W.max(axis=1) >= 0.99)
        print("iter %2d  exactly one-hot %.4f  >=0.99 %.4f  LL %r" % (it + 1, onehot.mean(), near.mean(), float(eng.em_status()[2][0])))
