"""`scSplit run` drop-in: same flags, same input CSVs, same five output files and log lines (scSplit:467-600).

`run` is the hot path; `genotype` (section 8f rank 3, a direct consumer of scSplit_P_s_c.csv) reuses the
hard-sum kernel; `count` (BAM pile-up, pysam) is upstream of the path and is reported as out of scope.
"""
import argparse
import datetime
import os
import sys

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

from . import models as M


def _log(out, text):
    with open(os.path.join(out, r'scSplit.log'), 'a') as logfile:
        logfile.write(text)


def read_count_csv(path, threads=0):
    """One dense count CSV -> (csr int32 V x N, SNV index, barcode index) through the library's streaming parser
    (csrc/csv.cu); files it declines (quoted / non-integer fields) go through pandas exactly like scSplit:544-546."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.scs_csv_parse(os.fsencode(path), int(threads), ctypes.byref(h))
    if rc == 2:
        frame = pd.read_csv(path, header=0, index_col=0)
        return csr_matrix(frame.values), frame.index, frame.columns
    _lib.check(rc)
    try:
        shp = (ctypes.c_int64 * 5)()
        _lib.check(lib.scs_csv_shape(h, shp))
        V, N, nnz, nid, nbc = (int(x) for x in shp)
        indptr, indices, data = np.empty(V + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.int32)
        ids, bcs = ctypes.create_string_buffer(max(nid, 1)), ctypes.create_string_buffer(max(nbc, 1))
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(lib.scs_csv_fetch(h, vp(indptr), vp(indices), vp(data), ids, bcs))
    finally:
        lib.scs_csv_free(h)
    snv = ids.raw[:nid].decode().split('\n')[:-1] if nid else []
    barcodes = bcs.raw[:nbc].decode().split('\n')[:-1] if nbc else []
    m = csr_matrix((data, indices, indptr), shape=(V, N))
    m.has_sorted_indices = True
    return m, pd.Index(snv, name='SNV'), pd.Index(barcodes)


def load_counts(ref_path, alt_path):
    """ref_filtered.csv / alt_filtered.csv -> base_calls_mtx = [csr ref, csr alt, SNV index, barcode index] (scSplit:544-547)."""
    ref_s, ref_index, ref_columns = read_count_csv(ref_path)
    alt_s, _, _ = read_count_csv(alt_path)
    return [ref_s, alt_s, ref_index, ref_columns]


def _preload_sklearn():
    try:
        from sklearn.cluster import KMeans            # noqa: F401
        from sklearn.decomposition import PCA         # noqa: F401
        from sklearn.preprocessing import StandardScaler  # noqa: F401
    except Exception:
        pass                                          # (the initialiser imports them itself and reports any real error)


def elbow(points):
    """Elbow of the LL curve, including the reference's quirk that `m` is never updated (scSplit:493-501)."""
    points = [-x for x in points]
    p1, p2, m, e = np.array([0, points[0]]), np.array([len(points) - 1, points[-1]]), 0, 0
    for i in range(1, (len(points) - 1)):
        p3 = np.array([i, points[i]])
        a, b = p2 - p1, p1 - p3
        d = abs(a[0] * b[1] - a[1] * b[0]) / np.linalg.norm(p2 - p1)
        if d > m:
            e = i
    return e


def frame_to_csv(frame, path, float_format=None):
    """`frame.to_csv(path)` / `frame.to_csv(path, float_format='%.0f')` (scSplit:596, 599-600), byte for byte, through the
    library's float writer (`scs_csv_write_f64`) instead of one Python string object per value; frames it does not cover
    (non-float columns, labels that are not plain strings or would need quoting) go through pandas itself.
    Returns True when the library wrote the file."""
    import ctypes
    from . import _lib
    index, columns = frame.index, frame.columns
    plain = (float_format in (None, '%.0f') and index.nlevels == 1 and columns.nlevels == 1 and len(columns) > 0
             and (index.name is None or isinstance(index.name, str))
             and all(dt == np.float64 for dt in frame.dtypes)
             and all(isinstance(c, (int, np.integer, str)) and not isinstance(c, (bool, np.bool_)) for c in columns))
    if plain:
        rows = index.tolist()
        head = ','.join([index.name or ''] + [str(c) for c in columns])
        plain = all(type(x) is str for x in rows)
    if plain:
        labels = '\n'.join(rows) + '\n' if rows else ''
        # a label holding a separator, a quote or a line break would be quoted by pandas: leave those frames to it
        plain = (labels.count('\n') == len(rows) and head.count(',') == len(columns)
                 and not any(ch in text for text in (labels, head) for ch in ('"', '\r')) and ',' not in labels and '\n' not in head)
    if not plain:
        frame.to_csv(path, float_format=float_format)
        return False
    values = np.ascontiguousarray(frame.values, dtype=np.float64)
    _lib.check(_lib.load().scs_csv_write_f64(os.fsencode(path), head.encode(), labels.encode(), len(rows), len(columns),
                                             values.ctypes.data_as(ctypes.c_void_p), 0 if float_format is None else 1))
    return True


def write_outputs(model, num, doublets, out):
    """scSplit_result.csv, scSplit_P_s_c.csv, scSplit_dist_variants.txt, scSplit_dist_matrix.csv, scSplit_PA_matrix.csv
    in the reference's formats (scSplit:582-600)."""
    frames = []
    for n in range(num):
        if n != model.doublet:
            frames.append(pd.DataFrame({'Barcode': model.reassigned[n], 'Cluster': 'SNG-' + str(n)}))
    if doublets != 0:
        frames.append(pd.DataFrame({'Barcode': model.reassigned[model.doublet], 'Cluster': 'DBL-' + str(model.doublet)}))
    frames = [f for f in frames if len(f)] or [pd.DataFrame({'Barcode': [], 'Cluster': []})]
    pd.concat(frames).to_csv(os.path.join(out, r'scSplit_result.csv'), sep='\t', index=False)
    frame_to_csv(model.P_s_c, os.path.join(out, r'scSplit_P_s_c.csv'))
    with open(os.path.join(out, r'scSplit_dist_variants.txt'), 'w') as f:
        for item in model.dist_variants:
            f.write(str(item) + '\n')
    frame_to_csv(model.dist_matrix, os.path.join(out, r'scSplit_dist_matrix.csv'), float_format='%.0f')
    frame_to_csv(model.pa_matrix, os.path.join(out, r'scSplit_PA_matrix.csv'), float_format='%.0f')


def run(argv):
    parser = argparse.ArgumentParser(prog='scSplit run', description='Demultiplex the scRNA-Seq using REF/ALT count matrices')
    parser.add_argument('-r', '--ref', required=True, help='REF count CSV input')
    parser.add_argument('-a', '--alt', required=True, help='ALT count CSV input')
    parser.add_argument('-n', '--num', type=int, required=True, help='Number of known mixed samples, if set to 0 system will find the most likely number')
    parser.add_argument('-o', '--out', required=True, help='Output directory')
    parser.add_argument('-s', '--sub', type=int, required=False, default='10', help='(Optional) Largest possible number of subpopulations, default: 10')
    parser.add_argument('-e', '--ems', type=int, required=False, default='30', help='(Optional) Number of EM repeats to avoid local maximum, default: 30')
    parser.add_argument('-d', '--dbl', type=float, required=False, help='(Optional) Doublet proportion')
    parser.add_argument('-v', '--vcf', required=False, help='(Optional) VCF file for filtering distinguishing variants')
    parser.add_argument('--device', type=int, default=0, help='CUDA device ordinal (scsplit_b200 extension)')
    args = parser.parse_args(argv)

    if not os.path.exists(args.out):
        raise ValueError('Specified output directory does not exist')
    # launched by torchrun (one process per GPU): the restarts are sharded over the ranks (models.core_batched) and rank 0
    # alone post-processes and writes the outputs
    rank, world, own_group = 0, int(os.environ.get('WORLD_SIZE', '1')), False
    if world > 1:
        import torch
        import torch.distributed as dist
        args.device = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(args.device)
        if not dist.is_initialized():
            dist.init_process_group('nccl', device_id=torch.device('cuda', args.device))
            own_group = True
        rank = dist.get_rank()
    doublets = float(args.dbl) if args.dbl is not None else -1
    num = args.num if doublets == 0 else args.num + 1          # additional doublet state, scSplit:535-538
    # The initialiser's PCA / KMeans are scikit-learn's (scSplit:14-16 imports them at start-up; ~1 s of module loading):
    # load them while the CSV parser's threads and the device upload run, instead of in the first restart.
    import threading
    threading.Thread(target=_preload_sklearn, daemon=True).start()

    if rank == 0:
        _log(args.out, '\n' + ' '.join(['scSplit', 'run'] + list(argv)) + '\n')
        _log(args.out, 'Loading allele count matrices: ' + str(datetime.datetime.now()) + '\n')
    base_calls_mtx = load_counts(args.ref, args.alt)
    if rank == 0:
        _log(args.out, 'Allele counts matrices loaded: ' + str(datetime.datetime.now()) + '\n')

    if args.num == 0:                                            # autodetect sweep, scSplit:550-558
        all_models, LL = [], []
        for sub in range(2, args.sub + 1):
            model, _ = M.core_batched(base_calls_mtx, sub, args.out, args.ems, device=args.device)
            all_models.append(model)
            if rank == 0:
                LL.append(model.lP_c_s.max(axis=1).sum())
        if rank == 0:
            elb = elbow(LL)
            model = all_models[elb]
            num = elb + 2
    else:
        model, _ = M.core_batched(base_calls_mtx, num, args.out, args.ems, device=args.device)
    if rank != 0:                                                 # the other ranks have done their share of the restarts
        M.release_engines()
        if own_group:
            dist.destroy_process_group()
        return None

    model.define_doublet()
    model.refine_doublets(doublets)

    pos = []
    if args.vcf:
        try:
            import vcf                                            # PyVCF, optional exactly as in scSplit:566-575
            for record in vcf.Reader(open(args.vcf, 'r')):
                try:
                    if float(record.INFO['R2'][0]) > 0.9:
                        pos.append(model.all_POS.index(str(record.CHROM) + ':' + str(record.POS)))
                except Exception:
                    continue
        except Exception:
            pass
    model.distinguishing_alleles(pos)
    write_outputs(model, num, doublets, args.out)
    M.release_engines()
    if own_group:
        dist.destroy_process_group()
    return model


# GRCh37 contig table of the VCF header the reference writes (scSplit:679-692); kept verbatim for byte-compatible output
_VCF_CONTIGS = [("1", 249250621), ("10", 135534747), ("11", 135006516), ("12", 133851895), ("13", 115169878), ("14", 107349540),
                ("15", 102531392), ("16", 90354753), ("17", 81195210), ("18", 78077248), ("19", 59128983), ("2", 243199373),
                ("20", 63025520), ("21", 48129895), ("22", 51304566), ("3", 198022430), ("4", 191154276), ("5", 180915260),
                ("6", 171115067), ("7", 159138663), ("8", 146364022), ("9", 141213431), ("MT", 16569), ("X", 155270560),
                ("Y", 59373566)]


def _vcf_header():
    lines = ['##fileformat=VCFv4.2', '##fileDate=' + str(datetime.datetime.today()).split(' ')[0], '##source=scSplit']
    lines += ['##contig=<ID=%s,length=%d>' % c for c in _VCF_CONTIGS]
    lines += ['##FILTER=<ID=PASS,Description="All filters passed">',
              '##INFO=<ID=AN,Number=1,Type=Integer,Description="Total Allele Count">',
              '##INFO=<ID=AC,Number=A,Type=Integer,Description="Alternate Allele Count">',
              '##INFO=<ID=AF,Number=A,Type=Float,Description="Estimated Alternate Allele Frequency">',
              '##FORMAT=<ID=GL,Number=3,Type=Float,Description="Genotype Likelihood for RR/RA/AA">']
    return '\n'.join(lines) + '\n'


def genotype(argv):
    """`scSplit genotype` (scSplit:603-696): per-cluster genotype likelihoods from the hard (P(s|c) >= 0.9) allele sums.

    The sums alt.dot(A), (alt+ref).dot(A) (scSplit:637-639) are the integer cluster sums of `scs_cluster_counts`; the
    binomial pmf at p in {err, 0.5, 1-err}, the GP/GL rounding and the VCF text follow the reference on the host.
    """
    import csv
    from scipy.stats import binom
    from .engine import Engine
    parser = argparse.ArgumentParser(prog='scSplit genotype', description='Generate sample genotypes in VCF format')
    parser.add_argument('-r', '--ref', required=True, help='REF count CSV input')
    parser.add_argument('-a', '--alt', required=True, help='ALT count CSV input')
    parser.add_argument('-p', '--psc', required=True, help='generated P(S|C)')
    parser.add_argument('-o', '--out', required=True, help='Output directory')
    parser.add_argument('--device', type=int, default=0, help='CUDA device ordinal (scsplit_b200 extension)')
    args = parser.parse_args(argv)
    if not os.path.exists(args.out):
        raise ValueError('Specified output directory does not exist')
    ref_s, all_POS, _ = read_count_csv(args.ref)
    alt_s, _, _ = read_count_csv(args.alt)
    P_s_c = pd.read_csv(args.psc, header=0, index_col=0)
    num = len(P_s_c.columns)
    hit = (P_s_c.values >= 0.9)
    if (hit.sum(axis=1) > 1).any():
        raise ValueError('a barcode has P(s|c) >= 0.9 for two clusters')
    labels = np.where(hit.any(axis=1), hit.argmax(axis=1), -1).astype(np.int32)
    with Engine(ref_s, alt_s, device=args.device) as eng:
        n_ref, n_alt, _ = eng.cluster_counts(labels, num)
    k = pd.DataFrame(n_alt.astype(np.float64))
    n = pd.DataFrame((n_alt + n_ref).astype(np.float64))
    err = 0.01
    ids = all_POS.tolist()

    def log_pmf(p):
        return pd.DataFrame(binom.pmf(k, n, [p] * num), index=ids, columns=range(num)).apply(np.log10)

    lp_rr, lp_ra, lp_aa = log_pmf(err), log_pmf(0.5), log_pmf(1 - err)
    gl_rr, gl_ra, gl_aa = 10 ** lp_rr.astype(float), 10 ** lp_ra.astype(float), 10 ** lp_aa.astype(float)
    nom = gl_rr + gl_ra + gl_aa
    text = lambda frame: round(frame, 3).astype(str)
    gp = [text(gl_rr / nom), text(gl_ra / nom), text(gl_aa / nom)]
    lg = [text(lp_rr.astype(float)), text(lp_ra.astype(float)), text(lp_aa.astype(float))]
    cols = {'#CHROM': [i.split(':')[0] for i in ids], 'POS': [i.split(':')[1] for i in ids], 'ID': ids}
    for name in ('REF', 'ALT', 'QUAL', 'FILTER', 'INFO'):
        cols[name] = '.'
    cols['FORMAT'] = 'GP:GL'
    vcf_content = pd.DataFrame(cols, index=ids)
    for c in range(num):
        vcf_content[c] = gp[0][c] + ',' + gp[1][c] + ',' + gp[2][c] + ':' + lg[0][c] + ',' + lg[1][c] + ',' + lg[2][c]
    with open(os.path.join(args.out, r'scSplit.vcf'), 'w+') as f:
        f.write(_vcf_header())
        vcf_content.to_csv(f, index=False, sep='\t', quoting=csv.QUOTE_NONE, escapechar='"')
    return vcf_content


def main(argv=None):
    argv = sys.argv[1:] if argv is None else list(argv)
    if not argv or argv[0] in ('-h', '--help'):
        print('usage: scSplit <command> [<args>]\n\nCommands:\n   run       Demultiplex the scRNA-Seq using REF/ALT count matrices (B200 CUDA path)\n   genotype  Generate sample genotypes in VCF format')
        return 0
    if argv[0] == 'run':
        run(argv[1:])
        return 0
    if argv[0] == 'genotype':
        genotype(argv[1:])
        return 0
    if argv[0] == 'count':
        raise SystemExit('scsplit_b200 accelerates `scSplit run` (and `genotype`); `count` (BAM pile-up) is outside the EM hot path')
    raise SystemExit('unknown command %r' % argv[0])


if __name__ == '__main__':
    sys.exit(main())
