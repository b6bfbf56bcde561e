#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total ms and share."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    print(f"{'kernel':40s} {'launches':>8s} {'total ms':>12s} {'share':>8s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k:40s} {cnt[k]:8d} {v:12.3f} {100 * v / s:7.2f}%")
    print(f"{'all':40s} {sum(cnt.values()):8d} {s:12.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
