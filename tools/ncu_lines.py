#!/usr/bin/env python3
"""Summarise an ncu report: headline metrics + per-source-line instruction / stall-sample shares.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u, v = rows[0], rows[1], rows[2]
for k in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'lts__t_sector_hit_rate.pct',
          'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'] + \
         [x for x in h if x.startswith('smsp__average_warps_issue_stalled') and x.endswith('per_issue_active.ratio')]:
    if k in h:
        i = h.index(k)
        try:
            if 'stalled' in k and float(v[i]) < 0.3: continue
        except ValueError: pass
        print('%-86s %-10s %s' % (k, u[i], v[i]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, 0, '', collections.Counter()])
cur = hdr = None
for row in csv.reader(io.StringIO(src)):
    if not row: continue
    if row[0] == 'File Path': cur = row[1].split('/')[-1]; continue
    if row[0] == 'Function Name': continue
    if row[0] == 'Line No': hdr = row; continue
    if hdr is None or cur is None: continue
    try: ln = int(row[0])
    except ValueError: continue
    try:
        a = agg[(cur, ln)]
        a[0] += int(row[hdr.index('Instructions Executed')] or 0); a[1] += int(row[hdr.index('# Samples')] or 0)
        a[2] += int(row[hdr.index('Thread Instructions Executed')] or 0); a[3] = row[1][:110]
        for i, hh in enumerate(hdr):
            if hh.startswith('stall_') and '(' not in hh and row[i]: a[4][hh] += int(row[i])
    except (ValueError, IndexError): pass
tot = sum(x[0] for x in agg.values()) or 1; ts = sum(x[1] for x in agg.values()) or 1
print('total warp-instructions %d, stall samples %d' % (tot, ts))
byf = collections.Counter(); byfs = collections.Counter()
for (f, l), x in agg.items(): byf[f] += x[0]; byfs[f] += x[1]
for f, c in byf.most_common(): print('  %-28s inst %5.1f%%  samples %5.1f%%' % (f, 100 * c / tot, 100 * byfs[f] / ts))
print('--- top lines by instructions (cum%)')
cum = 0
for (f, l), x in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    cum += x[0]
    top = ','.join('%s %d' % (k.replace('stall_', ''), c) for k, c in x[4].most_common(2))
    print('%-18s %4d i%5.1f%% c%5.1f%% s%5.1f%% t%4.1f [%s] | %s' % (f, l, 100 * x[0] / tot, 100 * cum / tot, 100 * x[1] / ts, x[2] / max(1, x[0]), top, x[3].strip()[:70]))
# optional: region sums  (usage: ... report top_n file:lo-hi:name,...)
if len(sys.argv) > 3:
    print('--- regions')
    for spec in sys.argv[3].split(','):
        f, rng, name = spec.split(':'); lo, hi = map(int, rng.split('-'))
        i = sum(x[0] for (ff, l), x in agg.items() if ff == f and lo <= l <= hi); s = sum(x[1] for (ff, l), x in agg.items() if ff == f and lo <= l <= hi)
        print('  %-28s inst %5.1f%%  samples %5.1f%%' % (name, 100 * i / tot, 100 * s / ts))
