#!/usr/bin/env python
"""Per-level comparison of the quadtree kernel with the oracle on one configuration (debug aid)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle_py as O  # noqa: E402
from orb_slam2_ros2_b200 import api, synth  # noqa: E402

O.build()
w, h, nl, sc, nf, ini, mn, seed = 461, 667, 5, 1.1704192161560059, 2623, 29, 9, 1036
if len(sys.argv) > 1:
    w, h, nl, sc, nf, ini, mn, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7]), int(sys.argv[8])
img = synth.synth_image(h, w, seed)
ctx = api.Context(w, h, nf, nl, sc, ini, mn)
kps, desc = ctx.extract(img)
print("fast env", os.environ.get("ORBX_QT_FAST"), "stats", ctx.quadtree_stats(), "kps", len(kps))
for l in range(nl):
    lw, lh, sf, quota = ctx.level_info(l)
    c = ctx.level_corners(0, l)
    sel = ctx.level_selected(0, l)
    exp_idx, pops = O.quadtree_select(lw - 32, lh - 32, c[:, 0], c[:, 1], c[:, 2], quota)
    exp = np.stack([c[exp_idx, 0] + 16, c[exp_idx, 1] + 16, c[exp_idx, 2]], 1) if len(exp_idx) else np.zeros((0, 3), np.int32)
    same = sel.shape == exp.shape and np.array_equal(sel, exp)
    print(f"level {l}: {lw}x{lh} corners {len(c)} quota {quota} selected {len(sel)} expected {len(exp)} pops {pops} ->", "ok" if same else "MISMATCH")
    if not same:
        a = {tuple(r) for r in sel.tolist()}
        b = {tuple(r) for r in exp.tolist()}
        print("   only in kernel:", sorted(a - b)[:8], "only in oracle:", sorted(b - a)[:8], len(a - b), len(b - a))
ctx.close()
