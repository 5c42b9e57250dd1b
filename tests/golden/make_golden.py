"""Generates tests/golden/*.npz: expected images for a few seeded parity scenarios.

The reference cannot be executed (no GL stack) and ships no golden images, so
these fixtures are outputs of the CPU oracle (oracle/march_oracle.c), i.e. "parity unpinned"
against a running reference; they pin the oracle itself against regressions and give the GPU
tests an expected image that does not depend on running the oracle at test time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python"), os.path.dirname(HERE)]

import scenarios                      # noqa: E402
from oracle import orc                # noqa: E402

GOLDEN_CASES = ["c1_trilinear_128steps", "ragged_nearest", "u16_aniso_trilinear", "mip_nearest"]


def voxel_crc(vox):
    return int(np.bitwise_xor.reduce(vox.astype(np.uint32) * np.arange(1, vox.size + 1, dtype=np.uint32)))


def main():
    for cid in GOLDEN_CASES:
        _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
        vox, dims, bpv, vs = scenarios.volume(vname)
        cam = scenarios.camera(cname)
        okw, _ = scenarios.split_kwargs(kw)
        p = orc.make_params(W, H, dims, bpv, cam, voxel_size=vs, **okw)
        img, cnt, _ = orc.render(p, vox, nthreads=1)
        np.savez_compressed(os.path.join(HERE, f"{cid}.npz"), rgba=img, cam=cam.astype(np.float32),
                            voxel_crc=np.uint32(voxel_crc(vox)),
                            samples=np.uint64(cnt["samples"]), rays_hit=np.uint64(cnt["rays_hit"]))
        print(cid, img.shape, cnt, os.path.getsize(os.path.join(HERE, f"{cid}.npz")))


if __name__ == "__main__":
    main()
