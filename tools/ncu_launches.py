#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count/avg/min/max (us)."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    d, order = collections.defaultdict(list), []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("d4pg::", "")
        d[name].append(v)
        order.append((name, v, row.get("Grid Size", "")))
    tot = sum(sum(v) for v in d.values())
    print("%-44s %5s %9s %9s %9s %7s" % ("kernel", "n", "avg_us", "min_us", "max_us", "share"))
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print("%-44s %5d %9.2f %9.2f %9.2f %6.1f%%" % (k[:44], len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))
    return order


if __name__ == "__main__":
    order = main(sys.argv[1])
    if len(sys.argv) > 2:
        for o in order[:int(sys.argv[2])]:
            print(o)
