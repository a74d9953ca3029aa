#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':70s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        name = k.replace("<unnamed>::", "").split("(")[0]
        print(f"{name:70s} {a[0]:8d} {a[1] / 1e6:10.3f} {a[1] / a[0] / 1e3:9.1f} {a[1] / tot * 100:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
