#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        val = float(r[-1].replace(",", ""))
        unit = r[-2]
        ms = val / 1e6 if unit == "ns" else (val / 1e3 if unit in ("us", "usecond") else val)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
        total += ms
    print(f"| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|")
    for k, (n, ms) in agg.items():
        print(f"| `{k}` | {n} | {ms:.3f} | {ms / n:.3f} | {100 * ms / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
