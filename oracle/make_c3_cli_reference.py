"""TEST INFRASTRUCTURE ONLY -- runs the REFERENCE `scSplit run` CLI end to end at BASELINE config 3 (30k SNVs x 20k cells,
-n 8 + doublet state) with a REDUCED number of restarts (-e 2 by default; the default -e 30 would take over an hour on one
core), on the synthetic CSVs of scsplit_b200.synth, and records the wall time, the log's event times and the result file:

    python oracle/make_c3_cli_reference.py [-e 2]   ->   profiles/r02_ref_cli_c3.json, tests/golden/cli_c3_e<E>.npz
    python oracle/make_c3_cli_reference.py -e 8     ->   profiles/r02_ref_cli_c3_e8.json (timing record only, no golden)
    python oracle/make_c3_cli_reference.py -e 30    ->   profiles/r02_ref_cli_c3_e30.json, tests/golden/cli_c3_e30.npz (about two hours)

The same CSVs, seed and -e go through this repo's CLI in tools/cli_e2e.py --workload c3 -e <E> --seed 2003."""
import json, os, re, shutil, sys, tempfile, time, warnings
import numpy as np, pandas as pd
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE); sys.path.insert(0, ROOT)
from oracle import ref_loader
from scsplit_b200 import synth


def main():
    ems = int(sys.argv[sys.argv.index("-e") + 1]) if "-e" in sys.argv else 2
    pool = synth.make_config("c3")
    tmp = tempfile.mkdtemp(prefix="c3cli_")
    try:
        rp, ap = os.path.join(tmp, "ref_filtered.csv"), os.path.join(tmp, "alt_filtered.csv")
        synth.write_csv(pool, rp, ap)
        np.random.seed(2003)
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_loader.run_reference_cli(["run", "-r", rp, "-a", ap, "-n", "8", "-o", tmp, "-e", str(ems)])
        wall = time.perf_counter() - t0
        res = pd.read_csv(os.path.join(tmp, "scSplit_result.csv"), sep="\t")
        log = open(os.path.join(tmp, "scSplit.log")).read()
        out = {"what": "REFERENCE scSplit run, config 3 (30000 x 20000, -n 8, -e %d), one host core of the build container" % ems,
               "wall_s": wall, "ems": ems, "em_iterations": log.count("E-M Iteration "), "seed": 2003,
               "clusters": res["Cluster"].value_counts().to_dict(), "assigned_cells": int(len(res)),
               "log_tail": log.strip().split("\n")[-3:], "host_cores_used": 1, "csv_mb": os.path.getsize(rp) / 1e6}
        with open(os.path.join(ROOT, "profiles", "r02_ref_cli_c3.json" if ems == 2 else "r02_ref_cli_c3_e%d.json" % ems), "w") as f:
            json.dump(out, f)
        if ems in (2, 30):
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cli_c3_e%d.npz" % ems),
                                result_barcode=np.array(res["Barcode"].tolist(), dtype=str), result_cluster=np.array(res["Cluster"].tolist(), dtype=str),
                                wall_s=np.float64(wall), em_iterations=np.int64(log.count("E-M Iteration ")), seed=np.int64(2003), ems=np.int64(ems))
        print(json.dumps(out))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if os.path.exists("scSplit_dist_variants.txt"):
            os.remove("scSplit_dist_variants.txt")


if __name__ == "__main__":
    main()
