#!/bin/bash
# compute-sanitizer passes over a small end-to-end run (smoke-sized inputs keep the 10-50x slowdown affordable)
set -u
mkdir -p gpurun_out
cat > /tmp/san_driver.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from orb_slam2_ros2_b200 import api, synth
l, r = synth.synth_stereo_pair(240, 320, 0, 9)
ctx = api.Context(320, 240, 500, 4, 1.2, camera=api.Camera(300.0, 300.0, 160.0, 120.0, 0.1), max_batch=2)
res = ctx.stereo_frame(l, r)
q, qd, ex, src = synth.synth_area_queries(res.kps_left, res.desc_left, 300, 5, 320, 240, n_levels=4, th=15.0)
m = ctx.search_in_area(q, qd, ex)
ok = m["best_idx"] >= 0
k2 = np.zeros(len(q), api.KP_DTYPE)
ctx.verify_angle(m["best_idx"][ok], np.nonzero(ok)[0], m["best_dist"][ok].astype(np.float32), res.kps_left, k2)
rec = ctx.serialize_keyframe(3)
voc = synth.synth_vocabulary(6, 3, 1)
V = api.Vocabulary(ctx, **voc)
bow = ctx.bow_transform(V)
V.close()
b = ctx.stereo_batch(np.stack([l, l]), np.stack([r, r]))
ctx4 = api.Context(320, 240, 500, 4, 1.2, camera=api.Camera(300.0, 300.0, 160.0, 120.0, 0.1), max_batch=4)
b4 = ctx4.stereo_batch(np.stack([l, l, l, l]), np.stack([r, r, r, r]))
assert int(b4.n_matches[3]) == res.n_matches
ctx4.close()
# frame-sharded sequence on two ranks sharing the device (peer-memory gather) + the CUDA-graph single-pair path
ranks = [api.Context(320, 240, 500, 4, 1.2, camera=api.Camera(300.0, 300.0, 160.0, 120.0, 0.1), max_batch=2) for _ in range(2)]
comms = api.Communicator.local(ranks, 5)
seq_l, seq_r = np.stack([l] * 5), np.stack([r] * 5)
for rk in range(2):
    blk = api.frame_range(5, rk, 2)
    rec, _, _ = ranks[rk].sequence_stereo(seq_l[blk.start:blk.stop], seq_r[blk.start:blk.stop], 5, comms[rk], gather_to_host=False)
    assert int(rec[0]["n_matches"]) == res.n_matches
for o in comms + ranks:
    o.close()
for _ in range(3):
    assert ctx.stereo_frame(l, r).n_matches == res.n_matches
noise = np.random.default_rng(1).integers(0, 256, (240, 320), dtype=np.uint8)
ctx.stereo_frame(noise, noise)
ctx.stereo_frame(np.zeros((240, 320), np.uint8), np.zeros((240, 320), np.uint8))
g = ctx.rgbd_frame(l, synth.synth_depth_u16(240, 320, 1, 5000.0)) if False else None
print("driver ok", res.n_matches, len(rec), int(b.n_matches.sum()), len(bow["bow_ids"]))
ctx.close()
PY
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 python /tmp/san_driver.py > gpurun_out/san_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|driver ok|Error|hazard" gpurun_out/san_$tool.log | head -8
done
