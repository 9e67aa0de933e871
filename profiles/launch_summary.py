#!/usr/bin/env python
"""Per-kernel shares of an ncu launch list (``ncu --metrics gpu__time_duration.sum --clock-control none --csv``).

    python profiles/launch_summary.py profiles/launches_r1_bench_configC.csv > profiles/launches_r1_summary.txt
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = [ln for ln in open(path, newline="") if ln.startswith('"')]
    rd = csv.DictReader(rows)
    agg = collections.defaultdict(list)
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        us = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v * (1e3 if r["Metric Unit"] in ("ms", "msecond") else 1.0)
        name = re.sub(r"^void ", "", r["Kernel Name"])
        name = re.sub(r"\(.*$", "", name)
        name = name.replace("hmsim::<unnamed>::", "").replace("<unnamed>::", "")
        agg[name[:44]].append(us)
    total = sum(sum(v) for v in agg.values())
    n = sum(len(v) for v in agg.values())
    print(f"total {total / 1e3:.2f} ms over {n} launches ({path})")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:<44s} n={len(v):5d} mean={sum(v) / len(v):9.1f} us  max={max(v):9.1f}  share={100 * sum(v) / total:5.1f}%")


if __name__ == "__main__":
    main()
