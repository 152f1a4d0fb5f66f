"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv) of tools/diag_train.py."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'^void |ctk::<unnamed>::|cub::CUB_\w+::', '', name)[:56]
        v = float(row['Metric Value'].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 'ps': v / 1e6}.get(row['Metric Unit'], v)
        a = agg.setdefault(name, [0, 0.0, 0.0, 1e18])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v); a[3] = min(a[3], v)
    print('%-56s %6s %12s %10s %10s %10s' % ('kernel', 'n', 'total us', 'avg us', 'min us', 'max us'))
    for k, (n, t, mx, mn) in agg.items():
        print('%-56s %6d %12.1f %10.2f %10.2f %10.2f' % (k, n, t, t / n, mn, mx))


if __name__ == '__main__':
    main(sys.argv[1])
