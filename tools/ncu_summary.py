"""Summarise an ncu launch list (csv) or a full-set report (.ncu-rep) on stdout: python tools/ncu_summary.py <file> [kernel-substring]"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get('Metric Name') == 'gpu__time_duration.sum':
                v = float(d['Metric Value'].replace(',', ''))
                v = v / 1000 if d['Metric Unit'] == 'ns' else v * 1000 if d['Metric Unit'] == 'ms' else v
                agg[d['Kernel Name'][:70]].append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-72s n=%4d avg=%9.1f us share=%5.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))


WANT = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'inst_executed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active']


def full(path, pat):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r))
        if pat and pat not in d.get('Kernel Name', ''):
            continue
        print(d.get('Kernel Name', '')[:90])
        u = dict(zip(h, units))
        for w in WANT:
            if w in d:
                print('    %-70s %s %s' % (w, d[w], u.get(w, '')))
        st = [(k, v) for k, v in d.items() if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k]
        st.sort(key=lambda kv: -float(kv[1].replace(',', '') or 0))
        print('    stalls:', ', '.join('%s=%s' % (k.replace('smsp__pcsamp_warps_issue_stalled_', ''), v) for k, v in st[:8]))


if __name__ == '__main__':
    p = sys.argv[1]
    if p.endswith('.csv'):
        launches(p)
    else:
        full(p, sys.argv[2] if len(sys.argv) > 2 else '')
