#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, agg, n = None, collections.defaultdict(lambda: [0, 0.0]), 0
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            name = re.sub(r"\(.*", "", d["Kernel Name"])
            v = float(d["Metric Value"].replace(",", ""))
            v = v / 1e6 if d["Metric Unit"] == "ns" else v / 1e3 if d["Metric Unit"] == "us" else v
            agg[name][0] += 1
            agg[name][1] += v
            n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {n} launches, {tot:.3f} ms total (ncu-serialised, cold cache: compare SHARES)")
    print(f"{'kernel':72s} {'n':>5s} {'ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:72]:72s} {v[0]:5d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
