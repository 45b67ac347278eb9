#!/usr/bin/env python3
"""Small end-to-end pass over every kernel for `compute-sanitizer --tool memcheck python tools/sanitize.py`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scsplit_b200 import synth
from scsplit_b200.engine import Engine

for K, variant in ((2, 2), (5, 2), (9, 2), (9, 1), (17, 2)):
    pool = synth.make_pool(3000, 600, 3, 0.05, 0.1, seed=K)
    rng = np.random.default_rng(K)
    lab = np.full(600, -1, dtype=np.int32)
    lab[rng.choice(600, 2 * K, replace=False)] = np.arange(2 * K) % K
    with Engine(pool.ref, pool.alt) as eng:
        eng.set_variant(variant)
        eng.pattern_counts(rng.random(3000) < 0.5, rng.random(600) < 0.5)
        P0 = eng.seed_af(np.stack([lab, lab]), K)
        eng.em_begin(P0)
        eng.em_run()
        it, cv, ll = eng.em_status()
        W, P = eng.get(1, 1), eng.get(0, 1)
        with np.errstate(invalid="ignore"):
            hit = W >= 0.99
        eng.doublet_scores(P, hit.any(axis=1).astype(np.uint8))
        eng.refine_scores(P, np.where(hit.any(axis=1), hit.argmax(axis=1), -1).astype(np.int32))
        eng.cluster_counts(rng.integers(-1, K, 600).astype(np.int32), K)
        print("K=%d variant=%d iters=%s ll=%r" % (K, variant, it.tolist(), ll.tolist()))
print("sanitize pass complete")
