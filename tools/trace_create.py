#!/usr/bin/env python3
"""SCS_TRACE=1 python tools/trace_create.py -- host-side phase times of Engine creation (layout build) at config 3."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scipy.sparse import csr_matrix
from scsplit_b200.engine import Engine
cfg, pool = bench.workload("c3")
pin = bench.pin
mk = lambda m: csr_matrix((pin(m.data.astype(np.int16)), pin(m.indices.astype(np.int32)), pin(m.indptr.astype(np.int64))), shape=m.shape)
pr, pa = mk(pool.ref), mk(pool.alt)
for rep in range(3):
    t = time.perf_counter()
    eng = Engine(pr, pa, canonicalize=False)
    print("create %.2f ms" % (1e3 * (time.perf_counter() - t)), file=sys.stderr)
    eng.close()
