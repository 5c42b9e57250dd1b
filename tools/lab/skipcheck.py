"""Lab: checkpoint period of the skipping forms (VR_SKIP_CHECK = passes between checkpoints; read once per process,
so every period runs in its own process).  usage: python tools/lab/skipcheck.py <period>"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "volume-renderer_b200", "python")]
os.environ["VR_SKIP_CHECK"] = sys.argv[1]
import volren_b200 as vb
from volren_b200 import workloads

def t(ctx, reps=7):
    ctx.render_device(0); ctx.render_device(0)
    return min(ctx.render_device(0).kernel_ms for _ in range(reps))

out = []
with vb.Context(1920, 1080) as ctx:
    ctx.upload_synthetic((1024, 1024, 1024), 2, 4095, workloads.SEEDS["C4"])
    for cam in ("K2", "K0"):
        ctx.set_camera(workloads.camera_block(cam))
        for name, kw in (("win[1000,3000]", dict(alpha_scale=0.05, min_val=1000, max_val=3000)), ("win[2000,4000]", dict(alpha_scale=0.05, min_val=2000, max_val=4000)),
                         ("full(forced)", dict(alpha_scale=0.02, min_val=0, max_val=4095))):
            for filt in (1, 0):
                ctx.set_params(vb.default_params(filter=filt, kernel=vb.KERNEL_DIRECT, **kw)); ctx.render_device(0); ref = ctx.read_frame().view(np.uint32).copy()
                ctx.set_params(vb.default_params(filter=filt, empty_skip=vb.SKIP_ON, **kw))
                ms = t(ctx); ok = np.array_equal(ctx.read_frame().view(np.uint32), ref)
                out.append(f"C4 {cam} {name} f{filt}: {ms:.3f}{'' if ok else ' MISMATCH'}")
with vb.Context(1920, 1080) as ctx:
    ctx.upload_synthetic((512, 512, 512), 2, 4095, workloads.SEEDS["C3"])
    ctx.set_camera(workloads.camera_block("K2"))
    ctx.set_params(vb.default_params(filter=1, empty_skip=vb.SKIP_ON, alpha_scale=0.05, min_val=1000, max_val=3000, step_scale=0.5))
    out.append(f"C3 K2: {t(ctx):.3f}")
with vb.Context(1920, 1080) as ctx:
    ctx.upload_synthetic((512, 512, 300), 2, 4095, 0x5EED0011, voxel_size=(0.7, 0.7, 1.5))
    ctx.set_camera(workloads.camera_block("K2"))
    ctx.set_params(vb.default_params(filter=1, empty_skip=vb.SKIP_ON, alpha_scale=0.03, min_val=1200, max_val=3500))
    out.append(f"CT K2 [1200,3500]: {t(ctx):.3f}")
print(f"period {sys.argv[1]}: " + " | ".join(out))
