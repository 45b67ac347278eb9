import sys, time, cProfile, pstats; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
import bench
from scsplit_b200 import models as M
cfg, pool = bench.workload('c3')
mtx=[pool.ref, pool.alt, pd.Index(pool.snv_ids), pd.Index(pool.barcodes)]
np.random.seed(1)
import tempfile; out=tempfile.mkdtemp()
m = M.models(mtx, 9, out, _defer_seed=True)   # warm
pr=cProfile.Profile(); pr.enable()
t=time.perf_counter()
for i in range(3): m = M.models(mtx, 9, out, _defer_seed=True)
print('per init %.1f ms' % ((time.perf_counter()-t)/3*1e3))
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
