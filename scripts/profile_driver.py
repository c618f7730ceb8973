#!/usr/bin/env python
"""Launch shape for ncu: `steps` x one 64-frame stereo batch (one launch per kernel: run with ORBX_PIPE=1 ORBX_CHUNK=64), then one
packed sequence chunk, the matchers / serialisation / bag-of-words kernels, a TUM-shaped RGB-D batch and a 1080p batch.

    ORBX_PIPE=1 ORBX_CHUNK=64 ncu ... python scripts/profile_driver.py [steps] [stereo|all]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_ros2_b200 import api, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
what = sys.argv[2] if len(sys.argv) > 2 else "stereo"
c = synth.KITTI
B = 64
W, H, N = c["width"], c["height"], c["n_features"]
lefts, rights = synth.synth_stereo_pool(H, W, B, seed0=0)
dl, dr = torch.from_numpy(lefts).cuda(), torch.from_numpy(rights).cuda()
cam = api.Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["bl"])
ctx = api.Context(W, H, N, 8, 1.2, camera=cam, max_batch=B)
for _ in range(steps):
    res = ctx.stereo_batch_device(B, dl.data_ptr(), dr.data_ptr(), W, W * H)
ctx.synchronize()
if what == "all":
    rs = ctx.record_layout().record_bytes
    d_rec = torch.zeros((B, rs), dtype=torch.uint8, device="cuda")
    ctx.sequence_stereo_ptr(B, dl.data_ptr(), dr.data_ptr(), W, W * H, None, True, d_rec.data_ptr(), rs, True)  # + pack_records
    res = ctx.stereo_batch_device(B, dl.data_ptr(), dr.data_ptr(), W, W * H)
    kps = ctx.read_device(res.kps_und, (2 * B, N), api.KP_DTYPE)[0::2]
    desc = ctx.read_device(res.desc, (2 * B, N, 32), np.uint8)[0::2]
    qs, qds, exs = np.zeros((B, N), api.AREA_QUERY_DTYPE), np.zeros((B, N, 32), np.uint8), np.zeros((B, N), np.uint8)
    for f in range(B):
        q, qd, ex, _ = synth.synth_area_queries(kps[f], desc[f], N, 500 + f, W, H, 8, 15.0)
        qs[f], qds[f], exs[f, : len(ex)] = q, qd, ex
    t_q, t_d, t_x = torch.from_numpy(qs.view(np.uint8).reshape(B, -1)).cuda(), torch.from_numpy(qds).cuda(), torch.from_numpy(exs).cuda()
    o = [torch.zeros((B, N), dtype=torch.int32, device="cuda") for _ in range(3)]
    o_ratio = torch.zeros((B, N), dtype=torch.float32, device="cuda")
    ctx.search_in_area_batch_device(B, N, t_q.data_ptr(), t_d.data_ptr(), 0, t_x.data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o_ratio.data_ptr(), o[2].data_ptr())
    cap = ctx.serialized_capacity()
    rec = torch.zeros((B, cap), dtype=torch.uint8, device="cuda")
    sizes = torch.zeros(B, dtype=torch.int64, device="cuda")
    ctx.serialize_keyframes_device(B, 1, rec.data_ptr(), cap, sizes.data_ptr())
    V = api.Vocabulary(ctx, **synth.synth_vocabulary(10, 6, 0))
    ctx.bow_transform_batch_device(V, B, 4)
    ctx.synchronize()
    t = synth.TUM
    g = np.stack([synth.synth_image(t["height"], t["width"], 3000 + i) for i in range(B)])
    d = np.stack([synth.synth_depth_u16(t["height"], t["width"], 3000 + i, t["depth_scale"]) for i in range(B)])
    dg, dd = torch.from_numpy(g).cuda(), torch.from_numpy(d.view(np.int16)).cuda()
    tc = api.Context(t["width"], t["height"], 1000, 8, 1.2, camera=api.Camera(t["fx"], t["fy"], t["cx"], t["cy"], t["bl"], tuple(t["dist"]), t["depth_scale"]), max_batch=B)
    fs = t["width"] * t["height"]
    tc.rgbd_batch_device(B, dg.data_ptr(), t["width"], fs, dd.data_ptr(), 2 * t["width"], 2 * fs, api.DEPTH_U16)
    tc.synchronize()
    tc.close()
print("profile driver done", ctx.launch_count)
ctx.close()
