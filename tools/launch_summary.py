#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path, only="setk::"):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        if only and only not in k:
            continue
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print(f"{k:80s} n={len(v):4d} mean={sum(v) / len(v):10.1f} us  total={sum(v) / 1e3:9.2f} ms "
              f"{100 * sum(v) / tot:5.1f}%")
    print(f"total {tot / 1e3:.2f} ms")


if __name__ == "__main__":
    main(*sys.argv[1:])
