#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [steps_in_capture]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        t = float(row['Metric Value'].replace(',', ''))
        t *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(row['Metric Unit'], 1.0)
        name = re.sub(r'\(.*', '', row['Kernel Name'])[:100]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print('| share | total us | launches | avg us | kernel |')
    print('|---|---|---|---|---|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.001:
            continue
        print('| %5.1f%% | %9.1f | %4d | %8.1f | `%s` |' % (100 * v[1] / tot, v[1], v[0], v[1] / v[0], k))
    print('\ntotal %.1f us over %d launches' % (tot, sum(v[0] for v in agg.values())))


if __name__ == '__main__':
    main()
