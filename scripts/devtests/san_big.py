"""compute-sanitizer driver at the full KITTI / TUM / 1080p shapes (one frame each): run under memcheck / racecheck / initcheck"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from orb_slam2_ros2_b200 import api, synth  # noqa: E402

for name, c, nf, nl in (("K", synth.KITTI, 2000, 8), ("H", synth.HD, 5000, 12)):
    l, r = synth.synth_stereo_pair(c["height"], c["width"], 3, 17)
    ctx = api.Context(c["width"], c["height"], nf, nl, 1.2, camera=api.Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["bl"]), max_batch=2)
    res = ctx.stereo_frame(l, r)
    print(name, len(res.kps_left), res.n_matches)
    ctx.close()
t = synth.TUM
ctx = api.Context(t["width"], t["height"], 1000, 8, 1.2, camera=api.Camera(t["fx"], t["fy"], t["cx"], t["cy"], t["bl"], tuple(t["dist"]), t["depth_scale"]))
r = ctx.rgbd_frame(synth.synth_image(t["height"], t["width"], 5), synth.synth_depth_u16(t["height"], t["width"], 5, t["depth_scale"]))
print("T", len(r.kps))
ctx.close()
