"""SASS opcode histogram of the built library (cuobjdump -sass): python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "scsplit_b200", "libscsplit_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
total, per = collections.Counter(), collections.defaultdict(collections.Counter)
inst = collections.Counter()
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        k = re.search(r"(k_[a-z_0-9]+?)(I[LS]|E[vPi]|$)", name)
        fn = k.group(1) if k else None
        if fn:
            inst[fn] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(\.[A-Z0-9_.]+)?\s", line)
    if m:
        op = m.group(1)
        total[op] += 1
        if m.group(2) and op in ("LDS", "LDG", "STG", "RED", "UBLKCP", "SYNCS", "ATOMG", "REDG"):
            total[op + m.group(2)] += 0
        if fn:
            per[fn][op] += 1
            if op in ("LDS", "RED", "REDG") and m.group(2):
                per[fn][op + m.group(2).split(".E")[0][:8]] += 0
print("# SASS opcode histogram of scsplit_b200/libscsplit_b200.so (cuobjdump -sass; tools/sass_histogram.py), round 2, final library")
print("# architectures in the fat binary: %s" % archs)
print("# evidence to look for: UBLKCP (cp.async.bulk, TMA bulk copies of the table tiles / packed posteriors) with SYNCS (mbarrier), LDS.128 row gathers,")
print("# IADD3 (64-bit integer accumulation of the fixed-point E-step), RED/REDG (integer totals of the lock-step E-step), DADD/DFMA (fp64 M-step, Bayes),")
print("# no UTC*MMA / HMMA: the contraction is K <= 17 wide\n")
print("## whole library")
for op, n in total.most_common(48):
    if n:
        print("%-12s %d" % (op, n))
for want in ("UBLKCP", "SYNCS", "RED", "REDG", "ATOMG", "HMMA", "UTCHMMA", "UTCQMMA"):
    print("%-12s %d   (named evidence)" % (want, total.get(want, 0)))
for fn in ("k_estep_lk", "k_estep_q", "k_mstep_sparse", "k_spmm_tiled", "k_bayes", "k_finalize", "k_combine_q", "k_lk_emit"):
    if fn in per:
        print("\n## %s (%d instantiations)" % (fn, inst[fn]))
        print(", ".join("%s %d" % (op, n) for op, n in per[fn].most_common(18) if n))
