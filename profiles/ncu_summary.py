#!/usr/bin/env python
"""Compact text summary of an `ncu --set full` report (csv exports), for profiles/.
usage: ncu_summary.py <raw.csv> [<rep.ncu-rep> <kernel regex> ...]
  raw.csv        = ncu -i X.ncu-rep --page raw --csv
  per kernel regex: the 12 hottest source lines (stall samples), from the report."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit%"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]][:70]
        print(f"== {name}")
        for k, label in KEYS:
            if k in idx:
                print(f"   {label:>15}: {r[idx[k]]} {units[idx[k]]}")
        top = sorted(((float(r[idx[k]] or 0), k) for k in stall), reverse=True)[:4]
        print("   stalls/issue   : " + ", ".join(
            f"{k.split('issue_stalled_')[1].split('_per_')[0]} {v:.1f}" for v, k in top))
    args = sys.argv[2:]
    if len(args) >= 2:
        rep = args[0]
        for rx in args[1:]:
            out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                                  "cuda,sass", "--kernel-name", "regex:" + rx, "--launch-count", "1"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
            agg, tot, fname = {}, 0, None
            for r in csv.reader(out.splitlines()):
                if len(r) >= 2 and r[0] == "File Path":
                    fname = r[1].split("/")[-1]
                    continue
                if len(r) < 8 or r[0] in ("", "Line No") or r[2] != "-":
                    continue
                try:
                    ln, s = int(r[0]), int(r[6])
                except ValueError:
                    continue
                key = (fname, ln, r[1].strip()[:90])
                agg[key] = agg.get(key, 0) + s
                tot += s
            print(f"-- hottest source lines of {rx} (share of {tot} stall samples)")
            for (f, ln, src), s in sorted(agg.items(), key=lambda x: -x[1])[:12]:
                print(f"   {100 * s / max(tot, 1):5.1f}%  {f}:{ln}: {src}")


if __name__ == "__main__":
    main()
