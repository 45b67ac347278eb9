"""TEST INFRASTRUCTURE ONLY -- runs the REFERENCE `scSplit run` CLI end to end at BASELINE config 2 (10k x 5k, -n 4,
doublet state, -e 30) on the synthetic CSVs of scsplit_b200.synth, records wall time and stores the reference's
scSplit_result.csv as tests/golden/cli_c2.npz (barcode -> cluster).  Takes ~10-20 minutes on one core."""
import os, shutil, sys, tempfile, time, warnings
import numpy as np, pandas as pd
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE); sys.path.insert(0, ROOT)
from oracle import ref_loader
from scsplit_b200 import synth

def main():
    pool = synth.make_config("c2")
    tmp = tempfile.mkdtemp(prefix="c2cli_")
    try:
        rp, ap = os.path.join(tmp, "ref_filtered.csv"), os.path.join(tmp, "alt_filtered.csv")
        synth.write_csv(pool, rp, ap)
        np.random.seed(2002)
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_loader.run_reference_cli(["run", "-r", rp, "-a", ap, "-n", "4", "-o", tmp, "-e", "30"])
        wall = time.perf_counter() - t0
        res = pd.read_csv(os.path.join(tmp, "scSplit_result.csv"), sep="\t")
        log = open(os.path.join(tmp, "scSplit.log")).read()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cli_c2.npz"),
                            result_barcode=np.array(res["Barcode"].tolist(), dtype=str), result_cluster=np.array(res["Cluster"].tolist(), dtype=str),
                            wall_s=np.float64(wall), em_iterations=np.int64(log.count("E-M Iteration ")), seed=np.int64(2002),
                            truth=pool.truth, is_doublet=pool.is_doublet)
        print("reference CLI at config 2: %.1f s wall, %d EM iterations, clusters %s" % (wall, log.count("E-M Iteration "), res["Cluster"].value_counts().to_dict()))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if os.path.exists("scSplit_dist_variants.txt"): os.remove("scSplit_dist_variants.txt")

if __name__ == "__main__":
    main()
