#!/usr/bin/env python3
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (second half of the list = the last step).
usage: python tools/launch_summary.py <launches.csv> [fraction of the list to keep from the end, default 0.5]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ci = {h: i for i, h in enumerate(hdr)}
    names = [(r[ci['Kernel Name']], float(r[ci['Metric Value']])) for r in data if r[ci['Metric Name']] == 'gpu__time_duration.sum']
    keep = names[int(len(names) * (1 - frac)):]
    tot = sum(v for _, v in keep)
    agg = collections.OrderedDict()
    for k, v in keep:
        k = k.split('(')[0][:64]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    print(f"launches: {len(keep)}  sum of kernel durations: {tot / 1e3:.1f} us (cold-cache, serialised under ncu: compare shares)")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:64s} {c:3d} {v / 1e3:9.1f} us {100 * v / tot:5.1f}%")


if __name__ == "__main__":
    main()
