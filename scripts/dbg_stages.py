import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orb_slam2_ros2_b200 import api, synth
from oracle import oracle_py as O
img=synth.synth_image(376,1241,0)
ctx=api.Context(1241,376,2000,8,1.2)
kps,desc=ctx.extract(img)
e=O.extract(img)
for l in range(8):
    g=ctx.level_corners(0,l); o,nfb=O.fast_cells(e.pyr.level(l))
    print('level',l,'gpu',len(g),'oracle',len(o),'equal',np.array_equal(g,o))
    if not np.array_equal(g,o):
        n=min(len(g),len(o)); d=np.nonzero((g[:n]!=o[:n]).any(1))[0]
        print(' first diff at',d[:5], g[d[:3]] if len(d) else '', o[d[:3]] if len(d) else '')
        sg=set(map(tuple,g)); so=set(map(tuple,o)); print(' only gpu',len(sg-so),' only oracle',len(so-sg), list(sg-so)[:5], list(so-sg)[:5])
    w,h,_,q=ctx.level_info(l)
    sel=ctx.level_selected(0,l)
    idx,_=O.quadtree_select(w-32,h-32,g[:,0].astype(np.float32),g[:,1].astype(np.float32),g[:,2].astype(np.float32),q)
    exp=g[idx]+np.array([16,16,0])
    print('  quadtree on gpu corners: gpu',len(sel),'oracle',len(exp),'equal',np.array_equal(sel,exp))
