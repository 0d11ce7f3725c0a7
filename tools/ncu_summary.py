#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (for profiles/)."""
import collections
import csv
import sys


def main(path, title):
    print("# " + title)
    print("# per-launch times are cold-cache and serialised under ncu: compare SHARES of the step, not absolutes")
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        ns = v if unit.startswith("n") else v * 1e3 if unit.startswith("u") else v * 1e6 if unit.startswith("m") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print("launches %d   total %.2f ms" % (sum(a[0] for a in agg.values()), tot / 1e6))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s launches %4d  total %9.3f ms  avg %9.1f us  share %5.1f%%" % (k[:44], n, t / 1e6, t / n / 1e3, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
