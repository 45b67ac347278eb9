"""cProfile of `scSplit run` (this repo's CLI) on a BASELINE config: python tools/profile_cli.py [c3]"""
import cProfile, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scsplit_b200 import cli, synth, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synth.CONFIGS[name]
pool = synth.make_config(name)
tmp = tempfile.mkdtemp(prefix="cli_prof_")
rp, ap = os.path.join(tmp, "ref_filtered.csv"), os.path.join(tmp, "alt_filtered.csv")
synth.write_csv(pool, rp, ap)
_lib.load(); import torch; torch.zeros(1, device="cuda")
np.random.seed(2003)
pr = cProfile.Profile(); pr.enable(); t = time.perf_counter()
cli.run(["-r", rp, "-a", ap, "-n", str(cfg["samples"]), "-o", tmp, "-e", "30"])
print("wall %.3f s" % (time.perf_counter() - t)); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
