#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if not row["Metric Name"].startswith("gpu__time_duration"):
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T:.0f} us over {sum(cnt.values())} launches")
    for k, v in tot.most_common(top):
        print(f"{v:10.0f} us {100 * v / T:5.1f}% n={cnt[k]:5d} {k[:120]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
