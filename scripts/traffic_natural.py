#!/usr/bin/env python
"""Whole-step DRAM traffic with NATURAL cache state: sums dram__bytes_read/write over the kernels of one 64-frame batch call from
an `ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log (only two
counters, so every kernel runs once and the L2 is not flushed between kernels; ncu still serialises the launches).

    python scripts/traffic_natural.py gpurun_out/<tag>_traffic_p8c8.csv gpurun_out/<tag>_traffic_p1c64.csv > profiles/rNN_traffic_natural.json
"""
import collections
import csv
import json
import sys

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
out = {"_what": __doc__.strip().split("\n\n")[0]}
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault(int(r[ii]), {"k": r[ki]})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
    agg = collections.OrderedDict()
    for d in per.values():
        k = d["k"].split("(")[0].replace("orbx::", "").replace("void ", "").replace("_kernel", "").replace("<8>", "")
        a = agg.setdefault(k, {"launches": 0, "dram_read_MB": 0.0, "dram_write_MB": 0.0, "ncu_us": 0.0})
        a["launches"] += 1
        a["dram_read_MB"] += d["dram__bytes_read.sum"][0] * SCALE[d["dram__bytes_read.sum"][1]] / 1e6
        a["dram_write_MB"] += d["dram__bytes_write.sum"][0] * SCALE[d["dram__bytes_write.sum"][1]] / 1e6
        a["ncu_us"] += d["gpu__time_duration.sum"][0] * TIME[d["gpu__time_duration.sum"][1]]
    for a in agg.values():
        for k in ("dram_read_MB", "dram_write_MB", "ncu_us"):
            a[k] = round(a[k], 1)
    tot_r = sum(a["dram_read_MB"] for a in agg.values())
    tot_w = sum(a["dram_write_MB"] for a in agg.values())
    out[path.split("/")[-1]] = {"launches": len(per), "kernels": agg, "total_read_MB": round(tot_r, 1), "total_write_MB": round(tot_w, 1),
                                "total_MB_per_64_frames": round(tot_r + tot_w, 1)}
json.dump(out, sys.stdout, indent=1)
print()
