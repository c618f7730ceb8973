#!/usr/bin/env python
"""profiles/traffic.json from profiles/r1_final_all_kernels_full.txt: DRAM read+write bytes, ncu duration and issue
utilisation per kernel launch (64-frame launches), keyed by bench.py's stage names."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_all_kernels_full.txt")
txt = open(src).read()
names = {"pyramid_level0": "pyramid_level0", "pyramid_levels": "pyramid_levels", "fast_cells": "fast_cells", "quadtree": "quadtree", "orient_brief": "orient_brief", "frame_index": "frame_index",
         "stereo": "stereo_match", "grid": "grid", "area_match": "area_match", "serialize": "serialize"}
out = {}
for blk in txt.split("## ")[1:]:
    name = blk.split("(")[0].strip().replace("_kernel", "")
    name = re.sub(r"<.*>", "", name.split("::")[-1]).replace("void ", "").strip()
    m = re.search(r"traffic = dram read \+ write\s+([\d.]+) (\w+)", blk)
    t = re.search(r"gpu__time_duration.sum\s+([\d.]+) (\w+)", blk)
    i = re.search(r"smsp__issue_active.avg.pct_of_peak_sustained_active\s+([\d.]+)", blk)
    mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[m.group(2)]
    tm = {"us": 1e-3, "ms": 1, "ns": 1e-6}[t.group(2)]
    out[names.get(name, name)] = {"dram_bytes_per_launch": float(m.group(1)) * mult, "ncu_ms": float(t.group(1)) * tm, "issue_active_pct": float(i.group(1)),
                                  "frames_per_launch": 64}
out["_source"] = "profiles/" + os.path.basename(src) + " (ncu --set full --clock-control none, 64-frame launches, K/2000)"
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
for k, v in out.items():
    if k != "_source":
        print("%-14s %8.1f MB  %7.3f ms  issue %5.1f %%" % (k, v["dram_bytes_per_launch"] / 1e6, v["ncu_ms"], v["issue_active_pct"]))
