#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total ms, share.
usage: python tools/launch_summary.py launches.csv ['# header line' ...]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4], [0, 0.0]); a[0] += 1; a[1] += float(r[14]) / 1e6
tot = sum(v[1] for v in agg.values()) or 1
for h in sys.argv[2:]: print(h)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-72s launches=%3d total_ms=%10.3f share=%5.1f%%' % (k[:72], v[0], v[1], 100 * v[1] / tot))
