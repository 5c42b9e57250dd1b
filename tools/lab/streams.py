"""Lab: does creating extra CUDA streams before the band streams slow the pipelined host path on a small partition
(hardware-queue aliasing, CUDA_DEVICE_MAX_CONNECTIONS)?  usage: python tools/lab/streams.py <extra streams>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "volume-renderer_b200", "python")]
import torch
import volren_b200 as vb
from volren_b200 import workloads
extra = [torch.cuda.Stream() for _ in range(int(sys.argv[1]))]
W, H = 1920, 1080
with vb.Context(W, H) as ctx:
    ctx.upload_synthetic((1024, 1024, 1024), 2, 4095, workloads.SEEDS["C4"])
    ctx.set_camera(workloads.camera_block("K2"))
    ctx.set_params(vb.default_params(alpha_scale=0.02, min_val=0, max_val=4095, filter=1))
    ctx.set_partition(0, 8, 8)
    bufs = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    for s in extra:                                  # put some work on them, as bench.py's device-timed loop does
        with torch.cuda.stream(s):
            torch.zeros(16, device="cuda")
    torch.cuda.synchronize()
    res = []
    for rep in range(3):
        t = {}
        for f in range(1, 64):
            if f == 4:
                ctx.render_wait(t.pop(3)); torch.cuda.synchronize(); t0 = time.perf_counter()
            t[f] = ctx.render_submit(bufs[f % 2].data_ptr())
            if f - 1 in t:
                ctx.render_wait(t.pop(f - 1))
        ctx.render_wait(t.pop(63))
        res.append((time.perf_counter() - t0) * 1e3 / 60)
    print(f"extra streams {sys.argv[1]}, CUDA_DEVICE_MAX_CONNECTIONS={os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS')}: pipelined 1/8 partition {min(res):.3f} ms/frame ({[round(r, 3) for r in res]})")
