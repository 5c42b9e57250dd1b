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
    """CTA shape / resident-warp variants (VR_LAB_TP = depth, resident warps per SM, warps per CTA)."""
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    tri = dict(filter=vb.FILTER_TRILINEAR)
    off = dict(empty_skip=vb.SKIP_OFF)
    on = dict(empty_skip=vb.SKIP_ON)
    combos = ((4, 32, 8), (4, 32, 4), (4, 32, 42), (4, 32, 41), (4, 32, 82), (4, 32, 21), (4, 32, 2), (3, 36, 42))
    if os.environ.get("LAB_COMBOS"):                       # e.g. LAB_COMBOS="4,32,42;4,24,42;3,24,42"
        combos = tuple(tuple(int(v) for v in c.split(",")) for c in os.environ["LAB_COMBOS"].split(";"))
    tpv = [("default", None, off)] + [(f"d{d} w{w} cta{c}", f"{d},{w},{c}", off) for d, w, c in combos]
    skv = [("skip default", None, on)] + [(f"skip d{d} w{w} cta{c}", f"{d},{w},{c}", on) for d, w, c in
                                          (combos if os.environ.get("LAB_COMBOS") else ((4, 32, 8), (4, 32, 4), (4, 32, 42), (4, 32, 82)))]
    n = 256 if quick else 1024
    with vb.Context(1920, 1080) as ctx:
        ctx.upload_synthetic((n, n, n), 2, 4095, workloads.SEEDS["C4"])
        kw = dict(alpha_scale=0.02, min_val=0, max_val=4095, **tri)
        run_case(f"C4 {n}^3 full", ctx, kw, tpv, cams=("K2", "K0", "K1"))
        run_case(f"C4 {n}^3 rank0/8 T8", ctx, kw, tpv, partition=(0, 8, 8))
        run_case(f"C4 {n}^3 rank3/8 T8", ctx, kw, tpv[:5], partition=(3, 8, 8))
        run_case(f"C4 {n}^3 rank0/4 T8", ctx, kw, tpv[:5], partition=(0, 4, 8))
        ctx.set_partition(0, 1, 8)
        run_case(f"C4 {n}^3 win[1000,3000]", ctx, dict(alpha_scale=0.05, min_val=1000, max_val=3000, **tri), skv, cams=("K2", "K0"))
        run_case(f"C4 {n}^3 win[1000,3000] rank0/8", ctx, dict(alpha_scale=0.05, min_val=1000, max_val=3000, **tri), skv, partition=(0, 8, 8))
        ctx.set_partition(0, 1, 8)
    with vb.Context(1920, 1080) as ctx:
        dims = (128, 128, 75) if quick else (512, 512, 300)
        ctx.upload_synthetic(dims, 2, 4095, 0x5EED0011, voxel_size=(0.7, 0.7, 1.5))
        run_case(f"CT {dims} aniso", ctx, dict(alpha_scale=0.03, min_val=200, max_val=3500, **tri), tpv[:8], cams=("K2",))
    with vb.Context(1024, 1024) as ctx:
        n2 = 128 if quick else 256
        ctx.upload_synthetic((n2, n2, n2), 1, 255, workloads.SEEDS["C2"])
        run_case(f"C2 {n2}^3 u8", ctx, dict(alpha_scale=0.05, min_val=0, max_val=255, step_scale=0.5, **tri), tpv[:8], cams=("K2",))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(RESULTS, open(os.path.join(ROOT, "gpurun_out", "lab_variants3.json"), "w"), indent=1)
    bad = [r for r in RESULTS if not r["exact"]]
    print(f"{len(RESULTS)} measurements, {len(bad)} mismatches")


if __name__ == "__main__":
    main()
