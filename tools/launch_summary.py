#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total ms, share.
Kernels are grouped into the ENCODE step (what bench.py's `value` / `roofline` time) and the decode block that
bench.py runs afterwards for its `decode_batch` key, so that the dominant kernel's share of the encode step can be
compared with bench.py's live `kernel_share_of_device_time`.
usage: python tools/launch_summary.py launches.csv ['# header line' ...]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
def group(name):
    if any(s in name for s in ('k_dec_', 'k_clean_', 'ctk::ToU64', 'ctk::U8ToU32', 'ScanTileState<unsigned long')): return 'decode block (bench.py decode_batch key)'
    if name.startswith('void at::') or 'at::native' in name: return 'torch (bench harness)'
    return 'encode step'
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault((group(r[4]), r[4]), [0, 0.0]); a[0] += 1; a[1] += float(r[14]) / 1e6
for h in sys.argv[2:]: print(h)
for g in ('encode step', 'decode block (bench.py decode_batch key)', 'torch (bench harness)'):
    items = [(k[1], v) for k, v in agg.items() if k[0] == g]
    if not items: continue
    tot = sum(v[1] for _, v in items) or 1
    print('== %s: %.3f ms in %d launches' % (g, tot, sum(v[0] for _, v in items)))
    for k, v in sorted(items, key=lambda kv: -kv[1][1]):
        print('%-72s launches=%3d total_ms=%10.3f share=%5.1f%%' % (k[:72], v[0], v[1], 100 * v[1] / tot))
