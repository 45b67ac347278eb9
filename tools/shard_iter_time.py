"""EM iteration time of one cell shard of a config on one GPU, auto variant (sparse M-step when it fits) vs tiled-only.

  python tools/shard_iter_time.py c4 0 25000        # the shard one of 8 GPUs holds at config 4
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scsplit_b200 import synth
from scsplit_b200.engine import Engine

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synth.CONFIGS[name]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else cfg["N"]
t0 = time.time()
pool = synth.make_config(name, cell_lo=lo, cell_hi=hi)
print("generated %s cells [%d,%d): %d + %d stored entries in %.1f s" % (name, lo, hi, pool.ref.nnz, pool.alt.nnz, time.time() - t0), flush=True)
K = cfg["K"]
rng = np.random.default_rng(0)
lab = np.full(hi - lo, -1, dtype=np.int32)
pick = rng.choice(hi - lo, 3 * K, replace=False)
lab[pick] = np.arange(3 * K) % K
with Engine(pool.ref, pool.alt) as eng:
    P0 = eng.seed_af(lab, K)
    for variant in (0, 2, 0, 2):
        eng.set_variant(variant)
        eng.em_begin(P0, max_iter=2000)
        eng.em_iterate(30, check_conv=False)
        eng.set_timing(True, every=4)
        eng.kernel_times(reset=True)
        eng.em_iterate(100, check_conv=False)
        kt = eng.kernel_times(reset=True)
        eng.set_timing(False)
        print("variant %d: %.1f us/iter  (E %.1f us, M %.1f us)" % (variant, eng.last_ms() * 10, kt["e_us"], kt["m_us"]), flush=True)
        eng.em_end()
