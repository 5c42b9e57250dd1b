import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python"), os.path.join(ROOT, "tests")]
import numpy as np
import volren_b200 as vb
import scenarios
cid = sys.argv[1] if len(sys.argv) > 1 else "smooth_trilinear_k1"
_, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
if len(sys.argv) > 3: W, H = int(sys.argv[2]), int(sys.argv[3])
vox, dims, bpv, vs = scenarios.volume(vname)
cam = scenarios.camera(cname)
_, vkw = scenarios.split_kwargs(kw)
imgs = {}
for k in (vb.KERNEL_DIRECT, vb.KERNEL_WINDOWED):
    with vb.Context(W, H) as ctx:
        ctx.upload_volume(vox, dims, vs); ctx.set_camera(cam); ctx.set_params(vb.default_params(kernel=k, **vkw))
        imgs[k], st = ctx.render()
        print("kernel", k, "used", st.kernel_used, "ms", st.kernel_ms)
bad = (imgs[1].view(np.uint32) != imgs[2].view(np.uint32)).any(axis=2)
ys, xs = np.nonzero(bad)
print("bad pixels", bad.sum(), "of", W * H)
if bad.sum():
    print("x range", xs.min(), xs.max(), "y range", ys.min(), ys.max())
    tiles = {}
    for y, x in zip(ys, xs): tiles[(x // 16, y // 16)] = tiles.get((x // 16, y // 16), 0) + 1
    print("tiles:", sorted(tiles.items())[:40])
    y, x = ys[0], xs[0]
    print("first", x, y, imgs[1][y, x], imgs[2][y, x])
