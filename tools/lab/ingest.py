"""Lab: volume ingest (vr_upload_volume_device) on a device-resident source: wall time of the whole call (pad + min/max
+ histogram + cell table; the layered arrays are built lazily and are not part of it), fused vs separate passes
(VR_INGEST_FUSED=0), and agreement of the two paths (cell table, stats, padded copy through a rendered frame).
usage: python tools/lab/ingest.py <fused:0|1>"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "volume-renderer_b200", "python")]
os.environ["VR_INGEST_FUSED"] = sys.argv[1]
import torch
import volren_b200 as vb
from volren_b200 import workloads

out = []
for dims, bpv in (((1024, 1024, 1024), 2), ((1024, 1024, 1024), 1), ((512, 512, 512), 2), ((2048, 2048, 1024), 2)):
    n = dims[0] * dims[1] * dims[2]
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    src = torch.randint(0, 4096 if bpv == 2 else 256, (n,), device="cuda", dtype=torch.int32, generator=g).to(torch.uint16 if bpv == 2 else torch.uint8)
    torch.cuda.synchronize()
    with vb.Context(640, 360) as ctx:
        ts = []
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ctx.upload_volume_device(src.data_ptr(), dims, bpv)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        t = ctx.cell_table(); mn, mx, hist = ctx.volume_stats()
        ctx.set_camera(workloads.camera_block("K1"))
        ctx.set_params(vb.default_params(alpha_scale=0.05, min_val=0, max_val=4095 if bpv == 2 else 255, filter=1, kernel=vb.KERNEL_DIRECT))
        img, _ = ctx.render()
    sig = (int(t["mins"].astype(np.int64).sum()), int(t["maxs"].astype(np.int64).sum()), mn, mx, float(np.asarray(hist).sum()), float(img.astype(np.float64).sum()))
    gb = n * bpv / 1e9
    out.append(f"{dims} u{8*bpv}: upload {min(ts):.2f} ms ({gb / (min(ts) * 1e-3):.0f} GB/s of source) sig {sig}")
    del src
print(f"fused={sys.argv[1]}:\n  " + "\n  ".join(out))
