"""Ad-hoc check: oracle area matchers vs the compiled reference (CPU only)."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import oracle_py as O
from orb_slam2_ros2_b200 import synth
rng=np.random.default_rng(0)
l,r=synth.synth_stereo_pair(376,1241,0)
tp='/tmp/tmpl.txt'; O.write_template_file(tp)
O.ref_reset(); O.ref_set_camera(718.856,718.856,607.19,185.2,0.537)
n,kps,desc,info=O.ref_extract(l,tp)
print(n)
b=(0.,0.,1241.,376.)
g1=O.grid_csr(kps,*b); g2=O.ref_grid_csr(kps,*b)
print(g1[0],g1[1], all(np.array_equal(a,c) for a,c in zip(g1[2:],g2[2:])))
g3=O.init_grid(kps,*b)
flat=np.concatenate([c for row in g3 for c in row]); print(np.array_equal(flat,g1[3]))
nq=1500
idx=rng.choice(n,nq,replace=False)
q=np.zeros(nq,O.AREA_QUERY_DTYPE)
q['x']=kps['x'][idx]+rng.normal(0,3,nq).astype(np.float32); q['y']=kps['y'][idx]+rng.normal(0,3,nq).astype(np.float32)
q['x']=np.clip(q['x'],0,1240); q['y']=np.clip(q['y'],0,375)
q['radius']=15; q['octave']=kps['octave'][idx]
q['min_level']=np.maximum(0,q['octave']-1); q['max_level']=np.minimum(7,q['octave']+1)
qd=desc[idx].copy()
flip=rng.integers(0,256,(nq,4))
for i in range(nq):
    for f in flip[i]: qd[i,f//8]^=1<<(f%8)
ex=(rng.random(n)<0.3).astype(np.uint8)
sf=info['sf']
for e in (None,ex):
    a=O.search_in_area(kps,desc,b,sf,q,qd,e); c=O.ref_search_in_area(kps,desc,b,sf,q,qd,e)
    print({k:bool(np.array_equal(a[k],c[k],equal_nan=True)) for k in a}, a['n_cand'].mean(), (a['best_idx']==idx).mean(), a['ratio'][:5])
ok=(a['best_idx']>=0)
qi=a['best_idx'][ok]; ti=np.nonzero(ok)[0].astype(np.int32); di=a['best_dist'][ok].astype(np.float32)
k2=np.zeros(nq,O.KP_DTYPE); k2['angle']=(kps['angle'][idx]+rng.normal(0,20,nq)).astype(np.float32)
r1=O.verify_angle(qi,ti,di,kps,k2); r2=O.ref_verify_angle(qi,ti,di,kps,k2)
print(len(qi),len(r1[0]), all(np.array_equal(x,y) for x,y in zip(r1,r2)))
