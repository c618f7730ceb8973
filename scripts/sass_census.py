#!/usr/bin/env python
"""SASS census of liborbx.so (no GPU needed): instructions per kernel that prove what each one uses -- UTMALDG / SYNCS (TMA + mbarrier),
LDG / LDS / STS / STG, VIMNMX (packed min / max), IDP (DP4A / DP2A), MATCH, REDUX, ATOMS, FP64, BAR.

    python scripts/sass_census.py > profiles/rNN_sass_census.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "orb_slam2_ros2_b200", "liborbx.so")], capture_output=True, text=True).stdout
cols = ["UTMALDG", "SYNCS", "LDG", "LDS", "STS", "STG", "VIMNMX", "IDP", "MATCH", "REDUX", "ATOMS", "FP64", "BAR"]
print("# cuobjdump -sass orb_slam2_ros2_b200/liborbx.so (sm_100a): static instruction counts per kernel")
print("%-44s %6s " % ("kernel", "SASS") + " ".join("%7s" % c for c in cols))
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0].strip()
    ops = collections.Counter(m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M))

    def cnt(*prefixes):
        return sum(v for k, v in ops.items() if any(k.startswith(p) for p in prefixes))

    dn = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dn = dn.replace("(anonymous namespace)::", "").replace("orbx::", "").replace("void ", "")
    dn = re.sub(r"\(.*", "", dn)
    vals = [cnt("UTMALDG"), cnt("SYNCS"), cnt("LDG"), cnt("LDS"), cnt("STS"), cnt("STG"), cnt("VIMNMX"), cnt("IDP"), cnt("MATCH"), cnt("REDUX"), cnt("ATOMS"),
            cnt("DMUL", "DADD", "DFMA", "DSETP", "F2F.F64", "I2F.F64", "F2F.F32.F64"), cnt("BAR")]
    print("%-44s %6d " % (dn[:44], sum(ops.values())) + " ".join("%7d" % v for v in vals))
print("# TMA box loads (cp.async.bulk.tensor.3d = UTMALDG.3D, completion on an mbarrier = SYNCS.*TRYWAIT) of fast_cells (cell patch), pyramid_levels (level-0 source rectangle) and orient_brief (blurred 37x37 patch):")
for line in txt.splitlines():
    if "UTMALDG" in line or "SYNCS" in line:
        print("#   " + line.strip()[:150])
