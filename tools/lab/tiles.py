import os, sys, numpy as np
ROOT="/root/repo"
sys.path[:0]=[os.path.join(ROOT,"volume-renderer_b200","python")]
import volren_b200 as vb
from volren_b200 import workloads
with vb.Context(1920,1080) as ctx:
    ctx.upload_synthetic((1024,1024,1024),2,4095,workloads.SEEDS["C4"])
    ctx.set_camera(workloads.camera_block("K2"))
    ctx.set_params(vb.default_params(alpha_scale=0.02,min_val=0,max_val=4095,filter=1))
    for world in (8,4):
        for T in (4,8,16,24,32,64):
            res=[]
            for rank in range(world):
                ctx.set_partition(rank,world,T)
                ctx.render_device(0); ctx.render_device(0)
                ms=min(ctx.render_device(0).kernel_ms for _ in range(7))
                res.append(ms)
            print(f"world {world} T {T:3d}: max {max(res):.3f} min {min(res):.3f} mean {np.mean(res):.3f}  per-rank {[round(x,3) for x in res]}", flush=True)
