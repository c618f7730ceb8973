#!/usr/bin/env python
"""Per-stage device times of the stereo path at several batch sizes (events between the kernels).

    python scripts/stage_times.py [batch ...]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_ros2_b200 import api, synth  # noqa: E402

c = synth.KITTI
batches = [int(a) for a in sys.argv[1:]] or [1, 8, 64]
P = max(batches)
lefts, rights = synth.synth_stereo_pool(c["height"], c["width"], min(P, 64), seed0=0)
reps = -(-P // lefts.shape[0])
dl = torch.from_numpy(np.concatenate([lefts] * reps)[:P]).cuda()
dr = torch.from_numpy(np.concatenate([rights] * reps)[:P]).cuda()
for B in batches:
    ctx = api.Context(c["width"], c["height"], 2000, 8, 1.2, camera=api.Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["bl"]), max_batch=B)
    acc = {}
    for r in range(12):
        t = ctx.profile_stereo_batch_device(B, dl.data_ptr(), dr.data_ptr(), c["width"], c["width"] * c["height"])
        if r >= 2:
            for k, v in t.items():
                acc.setdefault(k, []).append(v)
    med = {k: float(np.median(v)) for k, v in acc.items()}
    print(f"batch {B:3d}: " + "  ".join(f"{k}={v * 1e3:7.1f}us" for k, v in med.items()) + f"  total={sum(med.values()) * 1e3:8.1f}us")
    ctx.close()
