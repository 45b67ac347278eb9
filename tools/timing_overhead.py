"""Iteration time at a config with and without the per-kernel CUDA-event timing (what `bench.py` pays for its roofline numbers)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scsplit_b200 import synth
from scsplit_b200.engine import Engine

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synth.CONFIGS[name]
pool = synth.make_config(name)
rng = np.random.default_rng(0)
lab = rng.integers(0, cfg["K"], pool.ref.shape[1]).astype(np.int32)
with Engine(pool.ref, pool.alt) as eng:
    P0 = eng.seed_af(lab, cfg["K"])
    eng.em_begin(P0, max_iter=2000)
    eng.em_iterate(30, check_conv=False)
    for on in (False, True, False, True):
        eng.set_timing(on)
        eng.em_iterate(300, check_conv=False)
        print("timing", on, "us/iter %.2f" % (eng.last_ms() * 1000 / 300), flush=True)
