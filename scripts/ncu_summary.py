#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep     > profiles/rNN_<kernel>_full.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append((float(r[vi].replace(",", "")), r[gi], r[bi]))
    tot = sum(v[0] for vs in agg.values() for v in vs)
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold-cache: compare SHARES)")
    print(f"# {sum(len(v) for v in agg.values())} launches, total {tot / 1e6:.3f} ms")
    print(f"{'kernel':42s} {'n':>4s} {'mean_us':>10s} {'share':>7s}  grid / block")
    for k, vs in agg.items():
        t = sum(v[0] for v in vs)
        print(f"{k[:42]:42s} {len(vs):4d} {t / len(vs) / 1e3:10.1f} {t / tot:7.3f}  {vs[0][1]} / {vs[0][2]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on; source: {path}")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## {name}")
        for i, n in enumerate(hdr):
            if n in KEEP or ("issue_stalled" in n and n.endswith("_per_warp_active.pct")):
                print(f"{n:85s} {r[i]:>18s} {units[i]}")
        # ncu scales every COLUMN to its own unit (read may be in Mbyte while write is in byte): convert before adding
        dr, dw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = float(r[dr].replace(",", "")) * scale[units[dr]] + float(r[dw].replace(",", "")) * scale[units[dw]]
        print(f"{'traffic = dram read + write':85s} {total / 1e6:18.3f} Mbyte")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
