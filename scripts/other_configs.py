#!/usr/bin/env python
"""Throughput / latency of the non-headline configurations (BASELINE.json configs[1], [3], [4]) -- informational."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_ros2_b200 import api, synth  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


# configs[4]: 1920x1080 stereo, 5000 features, 12 levels
c = synth.HD
B = 16
l, r = synth.synth_stereo_pool(c["height"], c["width"], 4, seed0=0)
dl = torch.from_numpy(np.concatenate([l] * 4)).cuda()
dr = torch.from_numpy(np.concatenate([r] * 4)).cuda()
ctx = api.Context(c["width"], c["height"], 5000, 12, 1.2, camera=api.Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["bl"]), max_batch=B)
t = timed(lambda: ctx.stereo_batch_device(B, dl.data_ptr(), dr.data_ptr(), c["width"], c["width"] * c["height"]))
st = ctx.profile_stereo_batch_device(B, dl.data_ptr(), dr.data_ptr(), c["width"], c["width"] * c["height"])
print(f"HD 1920x1080/5000/12 stereo: {B / t:8.0f} frames/s (batch {B});  stages ms: " + ", ".join(f"{k}={v:.3f}" for k, v in st.items()))
one = api.Context(c["width"], c["height"], 5000, 12, 1.2, camera=api.Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["bl"]), max_batch=1)
t1 = timed(lambda: (one.stereo_batch_device(1, dl.data_ptr(), dr.data_ptr(), c["width"], c["width"] * c["height"]), one.synchronize()), n=30)
print(f"HD single pair device latency: {t1 * 1e3:.3f} ms")
ctx.close(); one.close()

# configs[1]: TUM-shaped RGB-D 640x480, 1000 features
c = synth.TUM
B = 64
g = np.stack([synth.synth_image(c["height"], c["width"], s) for s in range(8)] * 8)
d = np.stack([synth.synth_depth_u16(c["height"], c["width"], s, c["depth_scale"]) for s in range(8)] * 8)
dg, dd = torch.from_numpy(g).cuda(), torch.from_numpy(d.view(np.int16)).cuda()
cam = api.Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["bl"], tuple(c["dist"]), c["depth_scale"])
ctx = api.Context(c["width"], c["height"], 1000, 8, 1.2, camera=cam, max_batch=B)
fs = c["width"] * c["height"]
t = timed(lambda: ctx.rgbd_batch_device(B, dg.data_ptr(), c["width"], fs, dd.data_ptr(), 2 * c["width"], 2 * fs, api.DEPTH_U16))
print(f"TUM 640x480/1000 RGB-D: {B / t:8.0f} frames/s (batch {B}); algorithmic {ctx.algorithmic_bytes(False) + 18 * 1000} B/frame")
ctx.close()

# configs[3]: latency sweep on KITTI-shaped stereo
c = synth.KITTI
l, r = synth.synth_stereo_pair(c["height"], c["width"], 0, 17)
dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
for nf in (500, 1000, 2000, 4000):
    one = api.Context(c["width"], c["height"], nf, 8, 1.2, camera=api.Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["bl"]), max_batch=1)
    t1 = timed(lambda: (one.stereo_batch_device(1, dl.data_ptr(), dr.data_ptr(), c["width"], c["width"] * c["height"]), one.synchronize()), n=50)
    res = one.stereo_frame(l, r)
    print(f"KITTI stereo nFeatures={nf}: device latency {t1 * 1e3:.3f} ms, {len(res.kps_left)} kps, {res.n_matches} matches")
    one.close()
