"""Generates tests/golden/ref_*.npz with the REFERENCE'S OWN SHADER: /root/reference/VolumeRenderer.cs compiled
for the CPU into oracle/_ref/libshader_ref.so (oracle/Makefile; the GLSL text is compiled where it lies through
oracle/shim/glsl_compat.h).  Must run in the build container (where /root/reference exists); the fixtures travel.

    python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python"), os.path.dirname(HERE)]

import scenarios                      # noqa: E402
from oracle import orc                # noqa: E402

REF_GOLDEN = ["c1_nearest_ref_step", "ragged_trilinear", "u16_aniso_trilinear", "mip_nearest", "view_top",
              "window_min_gt_max", "iteration_cap_reference"]


def main():
    assert orc.ref_shader_lib() is not None, "oracle/_ref/libshader_ref.so missing: run `make -C oracle` with /root/reference present"
    for cid in REF_GOLDEN:
        _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
        vox, dims, bpv, vs = scenarios.volume(vname)
        cam = scenarios.camera(cname)
        okw, _ = scenarios.split_kwargs(kw)
        p = orc.make_params(W, H, dims, bpv, cam, voxel_size=vs, **okw)
        img, cnt = orc.ref_render(p, vox, nthreads=1)
        fn = os.path.join(HERE, f"ref_{cid}.npz")
        np.savez_compressed(fn, rgba=img, cam=cam.astype(np.float32), samples=np.uint64(cnt["samples"]))
        print(cid, img.shape, cnt, os.path.getsize(fn))


if __name__ == "__main__":
    main()
