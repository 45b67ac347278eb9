#!/usr/bin/env python3
"""End-to-end `scSplit run` through this repo's CLI on synthetic CSVs of a BASELINE config, with a phase breakdown.

  python tools/cli_e2e.py --workload c2 [-e 30] [--keep DIR]

Prints one JSON line (wall seconds per phase, EM iterations, cluster sizes, agreement with the simulated truth and --
when tests/golden/cli_<workload>.npz exists -- agreement with the REFERENCE's own result file for the same CSVs and its
recorded wall time)."""
import argparse, json, os, shutil, sys, tempfile, time
import numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)


def partition_agreement(a_labels, b_labels):
    """Fraction of items whose cluster in `a` maps to the majority cluster in `b` (permutation-free agreement)."""
    df = pd.DataFrame({"a": a_labels, "b": b_labels})
    return float(df.groupby("a")["b"].agg(lambda s: s.value_counts().iloc[0]).sum() / len(df))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("-e", "--ems", type=int, default=30)
    ap.add_argument("--seed", type=int, default=2002)
    ap.add_argument("--keep", default=None)
    args = ap.parse_args()
    print(json.dumps(run(args)))
    if os.path.exists("scSplit_dist_variants.txt"):
        os.remove("scSplit_dist_variants.txt")


def run(args):
    """args: workload, ems, seed, keep.  Returns the dict that main() prints."""
    from scsplit_b200 import cli, synth, models as M
    cfg = synth.CONFIGS[args.workload]
    pool = synth.make_config(args.workload)
    tmp = args.keep or tempfile.mkdtemp(prefix="cli_e2e_")
    os.makedirs(tmp, exist_ok=True)
    rp, ap_ = os.path.join(tmp, "ref_filtered.csv"), os.path.join(tmp, "alt_filtered.csv")
    t = time.perf_counter(); synth.write_csv(pool, rp, ap_); t_write = time.perf_counter() - t
    n = cfg["samples"]
    argv = ["-r", rp, "-a", ap_, "-n", str(n), "-o", tmp, "-e", str(args.ems)] + (["-d", "0"] if cfg["doublet_frac"] == 0 else [])
    # warm the CUDA context / library outside the timed region (the reference has no such start-up either)
    from scsplit_b200 import _lib; _lib.load(); import torch; torch.cuda.init(); torch.zeros(1, device="cuda")
    phases = {}
    orig_load, orig_core = cli.load_counts, M.core_batched
    def timed(name, fn):
        def wrap(*a, **k):
            t0 = time.perf_counter(); out = fn(*a, **k); phases[name] = phases.get(name, 0.0) + time.perf_counter() - t0; return out
        return wrap
    cli.load_counts = timed("load_csv", orig_load)
    M.core_batched = timed("init_and_em", orig_core)
    np.random.seed(args.seed)
    t0 = time.perf_counter()
    model = cli.run(argv)
    wall = time.perf_counter() - t0
    cli.load_counts, M.core_batched = orig_load, orig_core
    phases["post_and_write"] = wall - sum(phases.values())
    res = pd.read_csv(os.path.join(tmp, "scSplit_result.csv"), sep="\t")
    log = open(os.path.join(tmp, "scSplit.log")).read()
    pos = {b: i for i, b in enumerate(pool.barcodes)}
    idx = np.array([pos[b] for b in res["Barcode"]])
    truth = np.where(pool.is_doublet[idx], -1, pool.truth[idx])
    out = {"workload": "%s: %d SNVs x %d cells, -n %d, -e %d" % (args.workload, cfg["V"], cfg["N"], n, args.ems),
           "wall_s": wall, "phases_s": phases, "em_iterations": log.count("E-M Iteration "), "csv_mb": os.path.getsize(rp) / 1e6,
           "clusters": res["Cluster"].value_counts().to_dict(), "assigned_cells": int(len(res)),
           "agreement_with_truth": partition_agreement(res["Cluster"].values, truth)}
    gpath = os.path.join(ROOT, "tests", "golden", "cli_%s_e%d.npz" % (args.workload, args.ems))   # reference run with the same -e
    if not os.path.exists(gpath):
        gpath = os.path.join(ROOT, "tests", "golden", "cli_%s.npz" % args.workload)
    if os.path.exists(gpath) and int(np.load(gpath)["ems"] if "ems" in np.load(gpath) else 30) == args.ems:
        g = np.load(gpath)
        ref = pd.Series(g["result_cluster"], index=g["result_barcode"])
        both = res[res["Barcode"].isin(ref.index)]
        out["reference_wall_s"] = float(g["wall_s"])
        out["speedup_vs_reference_cli"] = float(g["wall_s"]) / wall
        out["reference_assigned_cells"] = int(len(ref))
        out["cells_assigned_by_both"] = int(len(both))
        out["agreement_with_reference"] = partition_agreement(both["Cluster"].values, ref[both["Barcode"]].values)
    if not args.keep:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


if __name__ == "__main__":
    main()
