import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from orb_slam2_ros2_b200 import api, synth
from oracle import oracle_py as O
img=synth.synth_image(376,1241,0)
ctx=api.Context(1241,376,2000,8,1.2)
kps,desc=ctx.extract(img)
e=O.extract(img)
l=7
lvl=e.pyr.level(l)
g=ctx.level_corners(0,l); o,nfb=O.fast_cells(lvl)
print(len(g),len(o))
def m_at(x,y):
    sub=np.ascontiguousarray(lvl[y-3:y+4,x-3:x+4])
    return O.lib().oracle_fast9_arc_value(sub[3:,3:].ctypes.data_as(C.POINTER(C.c_uint8)), sub.strides[0])
w,h,_,_=ctx.level_info(l)
print('level size',w,h)
for (x,y,s) in g[:12]:
    X,Y=x+16,y+16
    print('gpu corner roi',(x,y,s),'oracle m at pos',m_at(X,Y),'nbrs',[m_at(X+dx,Y+dy) for dy in (-1,0,1) for dx in (-1,0,1)])
print('oracle first', o[:12])
X,Y=27+16,4+16
print('oracle row at corner', lvl[Y, X-3:X+4], 'col', lvl[Y-3:Y+4, X])
