"""Lab: launch order of the CTA tiles on small grids (VR_LPT: 0 = row-major, 1 = longest rays first whenever the grid
has at most 65536 tiles, unset = the product's rule).  The variable is read once per process, so each mode runs in
its own process.  usage: python tools/lab/lpt.py <mode|auto>   (prints one line per partition / camera)"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "volume-renderer_b200", "python")]
if sys.argv[1] != "auto":
    os.environ["VR_LPT"] = sys.argv[1]
import volren_b200 as vb
from volren_b200 import workloads

def t(ctx, reps=9):
    ctx.render_device(0); ctx.render_device(0)
    return min(ctx.render_device(0).kernel_ms for _ in range(reps))

out = []
with vb.Context(1920, 1080) as ctx:
    ctx.upload_synthetic((1024, 1024, 1024), 2, 4095, workloads.SEEDS["C4"])
    for name, kw in (("headline", dict(alpha_scale=0.02, min_val=0, max_val=4095, filter=1)),
                     ("win[1000,3000]", dict(alpha_scale=0.05, min_val=1000, max_val=3000, filter=1)),
                     ("nearest", dict(alpha_scale=0.02, min_val=0, max_val=4095, filter=0))):
        for cam in ("K2", "K0", "K1"):
            ctx.set_camera(workloads.camera_block(cam))
            for world in (1, 2, 4, 8):
                ranks = (0, world // 2, world - 1) if world == 8 else (0,)
                ms = []
                for rank in ranks:
                    ctx.set_partition(rank, world, 8)
                    ctx.set_params(vb.default_params(kernel=vb.KERNEL_DIRECT, **kw)); ctx.render_device(0); ref = ctx.read_frame().view(np.uint32).copy()
                    dirty = dict(kw); dirty["alpha_scale"] = kw["alpha_scale"] * 3.0       # overwrite the owned rows with something else first
                    ctx.set_params(vb.default_params(**dirty)); ctx.render_device(0)
                    ctx.set_params(vb.default_params(**kw))
                    ms.append(t(ctx))
                    if not np.array_equal(ctx.read_frame().view(np.uint32), ref):
                        ms[-1] = float("nan")
                out.append(f"{name} {cam} 1/{world}: " + "/".join(f"{m:.3f}" for m in ms))
            if name != "headline":
                break
print(f"VR_LPT={sys.argv[1]}: " + " | ".join(out))
