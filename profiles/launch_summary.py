#!/usr/bin/env python
"""Per-kernel shares of an ncu launch list (``ncu --metrics gpu__time_duration.sum --clock-control none --csv``).

    python profiles/launch_summary.py profiles/launches_r1_bench_configC.csv > profiles/launches_r1_summary.txt
    python profiles/launch_summary.py LIST.csv --pass 2 --steps-per-pass 4

``--pass P``: only the launches of forward run number P (0-based; a forward run = ``--steps-per-pass`` simulator time
steps, each starting with ``k_tpfa_setup``, up to the ``k_gather_obs`` of its last step) - the timed pass of a bench
command without the prior sampling, the truth run and the update micro-benchmarks around it.
"""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--pass", dest="pass_", type=int, default=None)
    ap.add_argument("--steps-per-pass", type=int, default=40)
    args = ap.parse_args()
    path = args.path
    rows = [ln for ln in open(path, newline="") if ln.startswith('"')]
    rd = csv.DictReader(rows)
    launches = []
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        us = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v * (1e3 if r["Metric Unit"] in ("ms", "msecond") else 1.0)
        name = re.sub(r"^void ", "", r["Kernel Name"])
        name = re.sub(r"\(.*$", "", name)
        name = name.replace("hmsim::<unnamed>::", "").replace("<unnamed>::", "")
        launches.append((name[:44], us))
    what = path
    if args.pass_ is not None:
        starts = [i for i, (n, _) in enumerate(launches) if n.startswith("k_tpfa_setup")]
        first = starts[args.pass_ * args.steps_per_pass]
        last_step = starts[(args.pass_ + 1) * args.steps_per_pass - 1]
        end = next(i for i in range(last_step, len(launches)) if launches[i][0].startswith("k_gather_obs")) + 1
        launches = launches[first:end]
        what = f"{path}, forward run {args.pass_}: launches {first}..{end - 1}"
    agg = collections.defaultdict(list)
    for n, us in launches:
        agg[n].append(us)
    total = sum(sum(v) for v in agg.values())
    n = sum(len(v) for v in agg.values())
    print(f"total {total / 1e3:.2f} ms over {n} launches ({what})")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:<44s} n={len(v):5d} mean={sum(v) / len(v):9.1f} us  max={max(v):9.1f}  share={100 * sum(v) / total:5.1f}%")


if __name__ == "__main__":
    main()
