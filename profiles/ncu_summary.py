#!/usr/bin/env python
"""Condense an Nsight Compute report (``ncu --set full``) into the few numbers DESIGN.md / bench.py cite.

    python profiles/ncu_summary.py gpurun_out/foo.ncu-rep [--json out.json] > profiles/foo_summary.txt

Runs ``ncu -i <rep> --page raw --csv`` (works without a GPU) and prints, per captured launch: duration,
DRAM bytes read / written (the ``traffic`` of bench.py's roofline object), registers, the pipe utilisations
and the top warp-stall reasons.
"""

import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct_of_peak"),
    ("launch__registers_per_thread", "registers_per_thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_slot_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
]
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    rep = sys.argv[1]
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    summary = []
    for d in data:
        name = re.sub(r"\(.*", "", d[col["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        name = name.replace("hmsim::", "")
        rec = {"kernel": name}
        for key, short in KEEP:
            if key not in col or d[col[key]] == "":
                continue
            v = float(d[col[key]].replace(",", ""))
            u = units[col[key]]
            if short == "duration":
                rec["duration_us"] = v * UNIT_SCALE.get(u, 1.0)
            elif short in ("dram_read", "dram_write"):
                rec[short + "_MB"] = v * UNIT_SCALE.get(u, 1.0)
            else:
                rec[short] = v
        stalls = {}
        for h in hdr:
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
            if m and d[col[h]] != "":
                stalls[m.group(1)] = float(d[col[h]])
        rec["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        if "dram_read_MB" in rec:
            rec["traffic_MB"] = rec["dram_read_MB"] + rec.get("dram_write_MB", 0.0)
            if rec.get("duration_us"):
                rec["dram_GBps"] = rec["traffic_MB"] / rec["duration_us"] * 1e3
        summary.append(rec)
    for rec in summary:
        print(rec["kernel"])
        for k, v in rec.items():
            if k == "kernel":
                continue
            if isinstance(v, float):
                v = round(v, 3)
            print(f"    {k:28s} {v}")
    if out_json:
        json.dump(summary, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
