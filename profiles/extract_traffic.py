#!/usr/bin/env python
"""ncu raw CSV (one stage-1 pass captured with --set full) -> DRAM bytes per
phase of bench.py (`profiles/traffic.json`), plus a per-kernel table.

  ncu -i X.ncu-rep --page raw --csv > X.csv ; python profiles/extract_traffic.py X.csv
"""
import csv
import json
import sys

PHASE = [  # first match wins
    ("SketchKernel", "sketch"), ("GatherReadOffsets", "sketch"),
    ("MicromizeKernel", "micromize"),
    ("NarrowKeys", "index_sort"), ("WidenKeys", "index_sort"),
    ("IndexTableKernel", "index_table"), ("FillLongGaps", "index_table"),
    ("ProbeSortedKernel", "probe"), ("ProbeSuffixKernel", "probe"), ("ProbeKernel", "probe"),
    ("UnpackProbe", "probe"), ("ExpandWarpKernel", "expand"), ("ExpandKernel", "expand"),
    ("GatherU64", "expand"),
    ("SplitKernel", "chain"), ("GroupChainKernel", "chain"), ("ChainKernelGlobal", "chain"),
    ("SizeClassStarts", "chain"), ("GatherOverlapsByIndex", "chain"),
    ("LocateReadOverlaps", "chain"), ("OverlapCounts", "chain"), ("ReorderOverlaps", "chain"),
    ("ScatterMarks", "pile"), ("ApplyCoverage", "pile"),
    ("RhsKeys", "gather"), ("ListTotals", "gather"), ("CopyOld", "gather"),
    ("PlaceRhs", "gather"), ("PlaceLhs", "gather"), ("Truncate", "gather"),
    ("Compact", "gather"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    need = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum"]
    units = dict(zip(hdr, rows[1]))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_kernel, per_phase = {}, {}
    last_sort_owner = "index_sort"
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]]
        t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale.get(
            units["dram__bytes_read.sum"], 1)
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale.get(
            units["dram__bytes_write.sum"], 1)
        phase = None
        for key, ph in PHASE:
            if key in name:
                phase = ph
                break
        if phase is None and "DeviceRadixSort" in name:
            # CUB sorts: (u32 key, u64 value) = the index; (u64, u32) = query order
            # of the probe; (u32, u32) = gather / chain bookkeeping
            if "unsigned int, unsigned long" in name:
                phase = "index_sort"
            elif "unsigned long, unsigned int" in name:
                phase = "probe"
            else:
                phase = "chain"
        if phase is None and ("Scan" in name or "Iota" in name):
            phase = "scan+misc"
        phase = phase or "other"
        short = name.split("(")[0][-60:]
        k = per_kernel.setdefault(short, dict(launches=0, ms=0.0, dram_read=0.0, dram_write=0.0,
                                              phase=phase))
        k["launches"] += 1
        k["ms"] += t if units["gpu__time_duration.sum"] == "ms" else t / 1e3
        k["dram_read"] += rd
        k["dram_write"] += wr
        per_phase[phase] = per_phase.get(phase, 0.0) + rd + wr
    out = {k: round(v) for k, v in sorted(per_phase.items())}
    print(json.dumps(out, indent=1))
    print()
    for name, k in sorted(per_kernel.items(), key=lambda x: -x[1]["ms"])[:25]:
        print(f"{k['ms']:9.3f} ms x{k['launches']:<3d} rd {k['dram_read']/1e9:7.2f} GB "
              f"wr {k['dram_write']/1e9:7.2f} GB  [{k['phase']}] {name}")
    return out


if __name__ == "__main__":
    res = main(sys.argv[1])
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], "w"), indent=1)
