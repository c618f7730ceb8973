#!/usr/bin/env python
"""Per-CUDA-source-line instruction counts and stall samples from an .ncu-rep captured with --import-source on.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-regex> [top_n] [stalls]     ("stalls": order by stall samples)
"""
import csv
import re
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # blocks: a CUDA line row followed by its SASS rows; find the header
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    ii, si, wi = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    li = 0
    tot_i = tot_s = 0
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= max(ii, wi):
            continue
        if not re.match(r"^\d+$", r[li].strip()):
            continue  # SASS rows have hex addresses
        try:
            n = int(r[ii].replace(",", "") or 0)
            w = int(r[wi].replace(",", "") or 0)
        except ValueError:
            continue
        tot_i += n
        tot_s += w
        data.append((n, w, r[li], r[si].strip()[:105]))
    print(f"total warp instructions {tot_i}, stall samples {tot_s}")
    by_stalls = len(sys.argv) > 4 and sys.argv[4] == "stalls"
    for n, w, l, s in sorted(data, key=(lambda d: (d[1], d[0])) if by_stalls else None, reverse=True)[:top]:
        print(f"{n:12d} {100 * n / max(tot_i, 1):5.1f}%i {100 * w / max(tot_s, 1):5.1f}%s  L{l}: {s}")


if __name__ == "__main__":
    main()
