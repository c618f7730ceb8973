import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orb_slam2_ros2_b200 import api, synth
c=synth.KITTI
l,r=synth.synth_stereo_pair(c["height"],c["width"],0,17)
ctx=api.Context(c["width"],c["height"],2000,8,1.2)
for i in range(3):
    print('--- run',i, flush=True)
    ctx.stereo_frame(l,r)
