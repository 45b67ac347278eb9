import sys, time, numpy as np, torch, ctypes
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scsplit_b200 import synth
from scsplit_b200.engine import Engine
from scsplit_b200._lib import SCS_P, SCS_W, check
from scipy.sparse import csr_matrix
import bench
cfg, pool = bench.workload('c3'); K=cfg['K']
pin=bench.pin
ref, alt = pool.ref, pool.alt
pr = csr_matrix((pin(ref.data.astype(np.int16)), pin(ref.indices.astype(np.int32)), pin(ref.indptr.astype(np.int64))), shape=ref.shape)
pa = csr_matrix((pin(alt.data.astype(np.int16)), pin(alt.indices.astype(np.int32)), pin(alt.indptr.astype(np.int64))), shape=alt.shape)
eng = Engine(pr, pa, canonicalize=False)
P0 = eng.seed_af(bench.random_seed_labels(pool.ref.shape[1], K, 0), K); eng.close()
P0p, outW, outP = pin(P0), pin(np.empty((ref.shape[1], K))), pin(np.empty((ref.shape[0], K)))
for rep in range(10):
    t=[time.perf_counter()]
    eng = Engine(pr, pa, canonicalize=False); t.append(time.perf_counter())
    eng.em_begin(P0p); t.append(time.perf_counter())
    eng.em_iterate(1, check_conv=False); t.append(time.perf_counter())
    eng.em_iterate(7, check_conv=False); t.append(time.perf_counter())
    check(eng.lib.scs_em_get(eng.h, 0, SCS_W, outW.ctypes.data_as(ctypes.c_void_p)))
    check(eng.lib.scs_em_get(eng.h, 0, SCS_P, outP.ctypes.data_as(ctypes.c_void_p))); t.append(time.perf_counter())
    eng.em_status(); t.append(time.perf_counter())
    eng.close(); t.append(time.perf_counter())
    print('create %.2f  begin %.2f  iter1 %.2f  iter7 %.2f  get %.2f  status %.2f  close %.2f ms' % tuple(1e3*(b-a) for a,b in zip(t[:-1],t[1:])))
# raw H2D bandwidth
x = torch.empty(178_000_000, dtype=torch.uint8).pin_memory(); y = torch.empty_like(x, device='cuda')
torch.cuda.synchronize(); t0=time.perf_counter(); y.copy_(x, non_blocking=True); torch.cuda.synchronize(); print('H2D 178MB pinned: %.2f ms' % (1e3*(time.perf_counter()-t0)))
