"""TEST INFRASTRUCTURE ONLY -- goldens of the reference's `scSplit run` CLI outputs (tests/golden/cli_*.npz).

Runs the reference CLI (oracle/ref_loader.py) on small synthetic CSVs with np.random seeded by this harness and stores
the CSV inputs' generator parameters plus the parsed output files.
"""
import os
import shutil
import sys
import tempfile
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from scsplit_b200 import synth  # noqa: E402

CASES = {
    "cli_c1": dict(V=2000, N=500, samples=2, dfrac=0.0, lam=0.15, pseed=1001, n=2, d="0", e=3, seed=2001),
    "cli_k4": dict(V=1200, N=300, samples=3, dfrac=0.06, lam=0.2, pseed=31, n=3, d="0.1", e=3, seed=2002, genotype=True),
    "cli_k4_nod": dict(V=1200, N=300, samples=3, dfrac=0.06, lam=0.2, pseed=31, n=3, d=None, e=2, seed=2003),
    # autodetect (-n 0): sweep K = 2..s, elbow on the LL curve (scSplit:550-558, incl. the elbow quirk of 493-501)
    "cli_auto": dict(V=1200, N=300, samples=3, dfrac=0.0, lam=0.2, pseed=32, n=0, d="0", e=2, seed=2004, s=4),
}


def main():
    for name, c in CASES.items():
        pool = synth.make_pool(c["V"], c["N"], c["samples"], c["dfrac"], c["lam"], c["pseed"])
        tmp = tempfile.mkdtemp(prefix="cli_")
        try:
            rp, ap = os.path.join(tmp, "ref_filtered.csv"), os.path.join(tmp, "alt_filtered.csv")
            synth.write_csv(pool, rp, ap)
            argv = ["run", "-r", rp, "-a", ap, "-n", str(c["n"]), "-o", tmp, "-e", str(c["e"])]
            if c["d"] is not None:
                argv += ["-d", c["d"]]
            if "s" in c:
                argv += ["-s", str(c["s"])]
            np.random.seed(c["seed"])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref_loader.run_reference_cli(argv)
            res = pd.read_csv(os.path.join(tmp, "scSplit_result.csv"), sep="\t")
            psc = pd.read_csv(os.path.join(tmp, "scSplit_P_s_c.csv"), index_col=0)
            pa = pd.read_csv(os.path.join(tmp, "scSplit_PA_matrix.csv"), index_col=0)
            with open(os.path.join(tmp, "scSplit_dist_variants.txt")) as f:
                dv = [l.strip() for l in f if l.strip()]
            with open(os.path.join(tmp, "scSplit.log")) as f:
                log = f.read()
            with open(os.path.join(tmp, "scSplit_result.csv")) as f:
                head = f.readline()
            out = dict(params=np.array([str(c)]), result_barcode=np.array(res["Barcode"].tolist(), dtype=str), result_cluster=np.array(res["Cluster"].tolist(), dtype=str),
                       result_header=np.array([head]), psc=psc.values, psc_index=np.array(list(map(str, psc.index.tolist())), dtype=str), psc_columns=np.array(list(map(str, psc.columns.tolist())), dtype=str),
                       pa=pa.values.astype(np.float64), pa_index=np.array(list(map(str, pa.index.tolist())), dtype=str), pa_columns=np.array(list(map(str, pa.columns.tolist())), dtype=str),
                       dist_variants=np.array(sorted(dv)), log=np.array([log]))
            if c.get("genotype"):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ref_loader.run_reference_cli(["genotype", "-r", rp, "-a", ap, "-p", os.path.join(tmp, "scSplit_P_s_c.csv"), "-o", tmp])
                with open(os.path.join(tmp, "scSplit.vcf")) as f:
                    out["vcf"] = np.array([f.read()])
                with open(os.path.join(tmp, "scSplit_P_s_c.csv")) as f:
                    out["psc_csv"] = np.array([f.read()])
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
            print(name, res["Cluster"].value_counts().to_dict(), len(dv), "variants")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
            if os.path.exists("scSplit_dist_variants.txt"):
                os.remove("scSplit_dist_variants.txt")


if __name__ == "__main__":
    main()
