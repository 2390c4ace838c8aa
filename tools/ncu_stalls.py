#!/usr/bin/env python
"""Summarise the per-instruction warp-stall samples of an `ncu --set full --import-source on` capture.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass [--kernel-name regex:..] > sass.csv
       python tools/ncu_stalls.py sass.csv [top_n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[h]
    data = []
    for r in rows[h + 1:]:                     # first kernel section only
        if len(r) != len(hdr) or r[0] == 'Address':
            break
        data.append(r)
    isamp, isrc, iex = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stalls = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
    tot = sum(int(r[isamp]) for r in data)
    print('kernel:', rows[0][1][:100] if rows[0] else '?')
    print('total samples', tot, ' instructions', len(data))
    agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stalls}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print('  %-24s %8d  %5.1f%%' % (k, v, 100.0 * v / max(tot, 1)))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]
    print('top instructions (sass index, samples, executed, instruction, top stalls):')
    for i in sorted(top):
        r = data[i]
        st = sorted(((int(r[j]), hdr[j][6:]) for j in stalls), reverse=True)[:2]
        print('%5d %7s %9s  %-72s %s' % (i, r[isamp], r[iex], r[isrc][:72], st))


if __name__ == '__main__':
    main()
