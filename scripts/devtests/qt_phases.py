#!/usr/bin/env python
"""Cycle breakdown of the loop-free quadtree path (orbx_debug_quadtree_stats) on one image per configuration."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from orb_slam2_ros2_b200 import api, synth  # noqa: E402

NAMES = ["table nodes", "descent", "c*", "resp+ranks", "leaves", "surplus+flags", "PRE: cell scan + gather + keys + presort", "POST: ordered emit"]
for name, (h, w, nf, nl, seed) in {"K2000": (376, 1241, 2000, 8, 0), "K500": (376, 1241, 500, 8, 1), "K4000": (376, 1241, 4000, 8, 2), "T1000": (480, 640, 1000, 8, 3),
                                   "H5000": (1080, 1920, 5000, 12, 4)}.items():
    img = synth.synth_image(h, w, seed)
    ctx = api.Context(w, h, nf, nl, 1.2)
    reps = 5
    for _ in range(reps):
        ctx.extract(img)
    n_fast, ph = ctx.quadtree_phase_cycles()
    per = [c / max(n_fast, 1) for c in ph]
    print(f"{name}: {n_fast} loop-free problems; mean cycles per problem: " + ", ".join(f"{n}={c:.0f}" for n, c in zip(NAMES, per)) + f"; total {sum(per):.0f} cycles = {sum(per) / 1.965e3:.1f} us")
    ctx.close()
