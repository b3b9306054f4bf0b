import sys; sys.path.insert(0,'.')
import numpy as np, vkresample_b200 as vb
w,h=3840,2160
with vb.Plan(w,h,2.0,0,0.2) as p:
    p.execute(5); pk=p.profile_kernels(20); print("zeros input:", {k: round(v*1e3,1) for k,v in pk.items()})
    x=np.random.default_rng(0).random((3,h,w),dtype=np.float32); p.upload(p.pack_input(x)); p.execute(5); pk=p.profile_kernels(20); print("random input:", {k: round(v*1e3,1) for k,v in pk.items()})
