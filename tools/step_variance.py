import sys, torch
sys.path.insert(0,'/root/repo')
import splashsurf_b200 as ss
from splashsurf_b200 import synthetic as syn, distributed as ssd
p=syn.dam_break_50m()
ctx=ss.Context(0); params=ss.make_params(particle_radius=0.01, smoothing_length=2.0, cube_size=0.5)
r=ssd.Runner(ctx, params, 1, 0, 0)
x=torch.from_numpy(p).cuda()
import time
for i in range(10):
    t=time.perf_counter(); res=r.step(x); dt=(time.perf_counter()-t)*1e3
    tm=res["timings"]; ssum=sum(tm[k] for k in ("aabb_and_grid","decomposition","density","binning","levelset","marching_cubes","stitching"))
    print(i, "wall %.1f total_device %.1f stages %.1f | ls %.1f mc %.1f dens %.1f dec %.1f" % (dt, tm["total_device"], ssum, tm["levelset"], tm["marching_cubes"], tm["density"], tm["decomposition"]))
