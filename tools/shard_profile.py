"""A few EM iterations on one cell shard of a config, once per form of the fixed-point E-step (for A/B timing and ncu):
python tools/shard_profile.py c4 0 25000 [iterations] [forms: lk,groups]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scsplit_b200 import synth
from scsplit_b200.engine import Engine
name, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 12
forms = sys.argv[5].split(",") if len(sys.argv) > 5 else ["lk"]
cfg = synth.CONFIGS[name]
pool = synth.make_config(name, cell_lo=lo, cell_hi=hi, workers=8)
K = cfg["K"]
rng = np.random.default_rng(0)
lab = np.full(hi - lo, -1, dtype=np.int32)
lab[rng.choice(hi - lo, 3 * K, replace=False)] = np.arange(3 * K) % K
with Engine(pool.ref, pool.alt) as eng:
    P0 = eng.seed_af(lab, K)
    for form in forms:
        if form == "groups":
            os.environ["SCS_Q_GROUPS"] = "1"
        else:
            os.environ.pop("SCS_Q_GROUPS", None)
        eng.em_begin(P0, max_iter=2000)
        eng.em_iterate(8, check_conv=False)
        eng.set_timing(True, every=1)
        eng.kernel_times(reset=True)
        eng.em_iterate(n, check_conv=False)
        kt = eng.kernel_times(reset=True)
        eng.set_timing(False)
        print("%s %s: %.1f us/iter (E %.1f, bayes %.1f, M %.1f) LL %.6f" % (
            form, eng.em_info(), eng.last_ms() * 1000 / n, kt["e_us"], kt["bayes_us"], kt["m_us"], eng.em_status()[2][0]), flush=True)
        eng.em_end()
