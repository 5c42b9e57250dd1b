"""Kernel lab (development tool, runs on a B200 through gpurun): times pipeline-depth / occupancy variants of the
trilinear march (-DVR_LAB build, tools/lab/build_lab.sh) and the empty-space-skipping forms on BASELINE's
workloads, checking every frame bit for bit against the generic DIRECT kernel (which pytest checks against the
oracle).  Prints one line per measurement; JSON copy in gpurun_out/lab_variants.json.

    VOLREN_B200_LIB=volume-renderer_b200/lib/libvolren_b200_lab.so python tools/lab/variants.py [quick]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "volume-renderer_b200", "python")]
os.environ.setdefault("VOLREN_B200_LIB", os.path.join(ROOT, "volume-renderer_b200", "lib", "libvolren_b200_lab.so"))

import volren_b200 as vb                      # noqa: E402
from volren_b200 import workloads             # noqa: E402

RESULTS = []


def timed(ctx, reps=7):
    ms = []
    for _ in range(reps):
        st = ctx.render_device(0)
        ms.append(st.kernel_ms)
    return float(np.min(ms)), float(np.median(ms)), st


def frame(ctx):
    return ctx.read_frame()


def run_case(tag, ctx, base_kw, variants, cams=("K2",), partition=None):
    for cam in cams:
        ctx.set_camera(workloads.camera_block(cam))
        if partition:
            ctx.set_partition(*partition)
        os.environ.pop("VR_LAB_TP", None); os.environ.pop("VR_LAB_NN", None)
        ctx.set_params(vb.default_params(kernel=vb.KERNEL_DIRECT, **base_kw))
        ctx.render_device(0)
        ref = frame(ctx).view(np.uint32).copy()
        for name, env, kw in variants:
            for var in ("VR_LAB_TP", "VR_LAB_NN"):
                if env:
                    os.environ[var] = env
                else:
                    os.environ.pop(var, None)
            p = dict(base_kw); p.update(kw)
            ctx.set_params(vb.default_params(**p))
            ctx.render_device(0); ctx.render_device(0)
            lo, med, st = timed(ctx)
            exact = bool(np.array_equal(frame(ctx).view(np.uint32), ref))
            rec = dict(case=tag, cam=cam, variant=name, ms_min=lo, ms_med=med, kernel=st.kernel_used, skip=st.skip_used, exact=exact,
                       partition=partition)
            RESULTS.append(rec)
            print(f"{tag:28s} {cam} {name:26s} min {lo:7.3f} ms  med {med:7.3f} ms  kernel {st.kernel_used} skip {st.skip_used} "
                  f"{'bit-exact' if exact else 'MISMATCH'}", flush=True)
        os.environ.pop("VR_LAB_TP", None); os.environ.pop("VR_LAB_NN", None)


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    tri = dict(filter=vb.FILTER_TRILINEAR)
    nn = dict(filter=vb.FILTER_NEAREST)
    off, on = dict(empty_skip=vb.SKIP_OFF), dict(empty_skip=vb.SKIP_ON)
    tpv = [("default", None, off)] + [(f"depth{d} minb{m}", f"{d},{m}", off) for d, m in ((3, 4), (3, 5), (2, 6), (4, 4))]
    skv = [("skip off", None, off), ("skip on (default)", None, on)] + [(f"skip d{d}m{m}", f"{d},{m}", on) for d, m in ((2, 4), (2, 5), (3, 4), (2, 6))]
    nnv = [("default", None, off)] + [(f"depth{d} minb{m}", f"{d},{m}", off) for d, m in ((2, 8), (2, 6), (3, 8), (3, 6), (3, 5), (4, 6))]
    nskv = [("skip off", None, off), ("skip on (default)", None, on)] + [(f"skip d{d}m{m}", f"{d},{m}", on) for d, m in ((2, 5), (2, 6), (2, 4), (3, 5))]
    n = 256 if quick else 1024
    with vb.Context(1920, 1080) as ctx:
        ctx.upload_synthetic((n, n, n), 2, 4095, workloads.SEEDS["C4"])
        run_case(f"C4 {n}^3 a0.02 full", ctx, dict(alpha_scale=0.02, min_val=0, max_val=4095, **tri), tpv, cams=("K2", "K0"))
        run_case(f"C4 {n}^3 a0.02 nearest", ctx, dict(alpha_scale=0.02, min_val=0, max_val=4095, **nn), nnv, cams=("K2", "K0"))
        run_case(f"C4 {n}^3 win[1000,3000] a0.05", ctx, dict(alpha_scale=0.05, min_val=1000, max_val=3000, **tri), skv, cams=("K2", "K0", "K1"))
        print("empty cells", ctx.cell_table()["empty_cells"], "of", np.prod(ctx.cell_table()["cells"]), flush=True)
        run_case(f"C4 {n}^3 win[1000,3000] nearest", ctx, dict(alpha_scale=0.05, min_val=1000, max_val=3000, **nn), nskv, cams=("K2", "K0"))
        run_case(f"C4 {n}^3 win[2000,4000] a0.05", ctx, dict(alpha_scale=0.05, min_val=2000, max_val=4000, **tri), skv[:3], cams=("K2",))
        print("empty cells", ctx.cell_table()["empty_cells"], "of", np.prod(ctx.cell_table()["cells"]), flush=True)
        # a frame with NO empty cell: what the skipping form costs when it cannot skip
        run_case(f"C4 {n}^3 full window, skip forced", ctx, dict(alpha_scale=0.02, min_val=0, max_val=4095, **tri), skv[:3], cams=("K2",))
    n3 = 128 if quick else 512
    with vb.Context(1920, 1080) as ctx:
        ctx.upload_synthetic((n3, n3, n3), 2, 4095, workloads.SEEDS["C3"])
        kw = dict(alpha_scale=0.05, min_val=1000, max_val=3000, step_scale=0.5, **tri)
        run_case(f"C3 {n3}^3 win a0.05", ctx, kw, skv, cams=("K2", "K0"))
        print("C3 empty cells", ctx.cell_table()["empty_cells"], "of", np.prod(ctx.cell_table()["cells"]), "shift", ctx.cell_table()["shift"], flush=True)
    with vb.Context(1920, 1080) as ctx:
        dims = (128, 128, 75) if quick else (512, 512, 300)
        ctx.upload_synthetic(dims, 2, 4095, 0x5EED0011, voxel_size=(0.7, 0.7, 1.5))
        run_case(f"CT {dims} aniso [200,3500]", ctx, dict(alpha_scale=0.03, min_val=200, max_val=3500, **tri), tpv[:3] + skv[1:3], cams=("K2",))
        run_case(f"CT {dims} aniso [1200,3500]", ctx, dict(alpha_scale=0.03, min_val=1200, max_val=3500, **tri), skv[:4], cams=("K2", "K1"))
        print("CT empty cells", ctx.cell_table()["empty_cells"], "of", np.prod(ctx.cell_table()["cells"]), "shift", ctx.cell_table()["shift"], flush=True)
    with vb.Context(1024, 1024) as ctx:
        n2 = 128 if quick else 256
        ctx.upload_synthetic((n2, n2, n2), 1, 255, workloads.SEEDS["C2"])
        run_case(f"C2 {n2}^3 u8", ctx, dict(alpha_scale=0.05, min_val=0, max_val=255, step_scale=0.5, **tri), tpv, cams=("K2",))
        run_case(f"C2 {n2}^3 u8 [30,180]", ctx, dict(alpha_scale=0.05, min_val=30, max_val=180, step_scale=0.5, **tri), skv[:3], cams=("K2",))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(RESULTS, open(os.path.join(ROOT, "gpurun_out", "lab_variants2.json"), "w"), indent=1)
    bad = [r for r in RESULTS if not r["exact"]]
    print(f"{len(RESULTS)} measurements, {len(bad)} mismatches")


if __name__ == "__main__":
    main()
