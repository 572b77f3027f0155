import sys, os, time
sys.path.insert(0,'cuda-raytracer_b200'); sys.path.insert(0,'oracle')
import numpy as np, b2rt
from b2rt.scene import random_soup
n=int(sys.argv[1]); tb=int(sys.argv[2]); nr=int(sys.argv[3])
sc=random_soup(n,size=0.02)
bvh=b2rt.BVHAccel(sc,treelet_bytes=tb)
print(bvh.stats()['bvh_levels'], bvh.stats()['bvh_subtrees'], flush=True)
rng=np.random.default_rng(21)
o=rng.random((nr,3)).astype(np.float32); d=rng.normal(size=(nr,3)); d/=np.linalg.norm(d,axis=1,keepdims=True)
t0=time.time(); t,p=bvh.intersect(o,d.astype(np.float32)); print('ok',time.time()-t0,(p!=0xFFFFFFFF).mean(), bvh.stats()['queue_pushes']/nr, flush=True)
