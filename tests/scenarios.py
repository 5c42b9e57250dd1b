"""Seeded parity scenarios shared by the GPU parity tests and the golden-fixture generator.
Sizes are chosen so the CPU oracle finishes each case in well under a second."""
from __future__ import annotations

import math

import numpy as np

from volren_b200 import workloads

DEFAULT_KNOTS = [(0, 0.0), (141, 0.759), (149, 0.45), (255, 1.0)]   # AlphaControlSplineWidget.cpp:56-59


def volume(name):
    """-> (flat voxel array, dims, bytes per voxel, voxel_size)"""
    if name == "mix64_u8":          # BASELINE config 1
        return workloads.mix_volume((64, 64, 64), 255, workloads.SEEDS["C1"]), (64, 64, 64), 1, (1.0, 1.0, 1.0)
    if name == "smooth64_u8":       # same without the hash term
        return workloads.mix_volume((64, 64, 64), 255, workloads.SEEDS["C1"], with_hash=False), (64, 64, 64), 1, (1.0, 1.0, 1.0)
    if name == "rand_48x40x36_u8":  # ragged, every tex-coord divisor is a non power of two
        rng = np.random.default_rng(7)
        return rng.integers(0, 256, 48 * 40 * 36, dtype=np.uint8), (48, 40, 36), 1, (1.0, 1.0, 1.0)
    if name == "mix_64x64x32_u16":  # anisotropic spacing, power-of-two divisors (1, 1, 1)
        return workloads.mix_volume((64, 64, 32), 4095, workloads.SEEDS["C3"]), (64, 64, 32), 2, (1.0, 1.0, 2.0)
    if name == "rand_40x56x33_u16":
        rng = np.random.default_rng(11)
        return rng.integers(0, 4096, 40 * 56 * 33, dtype=np.uint16), (40, 56, 33), 2, (1.0, 0.8, 1.7)
    if name == "full_u16":          # uses the whole 16-bit range
        rng = np.random.default_rng(13)
        return rng.integers(0, 65536, 32 * 32 * 32, dtype=np.uint16), (32, 32, 32), 2, (1.0, 1.0, 1.0)
    if name == "const_u8":
        return np.full(16 * 16 * 16, 100, dtype=np.uint8), (16, 16, 16), 1, (1.0, 1.0, 1.0)
    if name == "one_voxel_u8":
        return np.array([200], dtype=np.uint8), (1, 1, 1), 1, (1.0, 1.0, 1.0)
    if name == "rod_1x1x12000_u8":  # |vol_size| = 12000 > 10000 samples along z at the REFERENCE step (no step override)
        return np.full(12000, 3, dtype=np.uint8), (1, 1, 12000), 1, (2000.0, 2000.0, 1.0)
    raise KeyError(name)


def camera(name):
    from volren_b200.host import Camera
    if name in ("K0", "K1", "K2"):
        return workloads.camera_block(name)
    cam = Camera(30.0)
    if name == "inside":            # eye inside the box: t_min < 0 (VolumeRenderer.cs:237 quirk)
        cam.setSpherical(0.3, math.radians(80.0), math.radians(10.0))
    elif name == "pole":            # zenith 0: the side vector comes from the rotated +x branch
        cam.setSpherical(2.5, 0.0, math.radians(30.0))
    elif name == "orbit":           # a few mouse drags + zooms through the reference API
        for dz, da in ((0.06, 0.0), (0.06, -0.06), (-0.06, 0.06), (0.0, 0.06)) * 5:
            cam.setOrientation(0, dz, da)
        cam.setOrientation(1, 0, 0)
    else:
        raise KeyError(name)
    return cam.ubo()


def tf_lut():
    from volren_b200.host import CubicSpline
    return CubicSpline(DEFAULT_KNOTS).bakeAlphaLUT()


# (id, volume, camera, (W,H), kwargs common to oracle.make_params and the vr_params fields)
CASES = [
    ("c1_nearest_ref_step", "mix64_u8", "K0", (256, 256), dict(alpha_scale=0.05, min_val=0, max_val=255, filter=0)),
    ("c1_trilinear_128steps", "mix64_u8", "K0", (256, 256), dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1, step_scale=0.5)),
    ("c1_trilinear_window_ert", "mix64_u8", "K1", (256, 256), dict(alpha_scale=1.0, min_val=40, max_val=200, filter=1)),
    ("c1_nearest_window_ert", "mix64_u8", "K2", (256, 256), dict(alpha_scale=1.0, min_val=40, max_val=200, filter=0)),
    ("c1_trilinear_k2_dense", "mix64_u8", "K2", (256, 256), dict(alpha_scale=0.02, min_val=0, max_val=255, filter=1)),
    ("smooth_trilinear_k1", "smooth64_u8", "K1", (200, 136), dict(alpha_scale=0.1, min_val=0, max_val=255, filter=1)),
    ("ragged_nearest", "rand_48x40x36_u8", "K1", (250, 131), dict(alpha_scale=0.08, min_val=10, max_val=250, filter=0)),
    ("ragged_trilinear", "rand_48x40x36_u8", "K2", (250, 131), dict(alpha_scale=0.08, min_val=10, max_val=250, filter=1)),
    ("u16_aniso_trilinear", "mix_64x64x32_u16", "K1", (256, 144), dict(alpha_scale=0.05, min_val=1000, max_val=3000, filter=1)),
    ("u16_aniso_nearest", "mix_64x64x32_u16", "K0", (256, 144), dict(alpha_scale=0.5, min_val=1000, max_val=3000, filter=0)),
    ("u16_ragged_trilinear", "rand_40x56x33_u16", "K2", (240, 135), dict(alpha_scale=0.03, min_val=0, max_val=4095, filter=1)),
    ("u16_ragged_nearest", "rand_40x56x33_u16", "orbit", (240, 135), dict(alpha_scale=0.03, min_val=0, max_val=4095, filter=0)),
    ("u16_full_range", "full_u16", "K1", (128, 128), dict(alpha_scale=0.04, min_val=0, max_val=65535, filter=1)),
    ("eye_inside", "mix64_u8", "inside", (160, 90), dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1)),
    ("pole_camera", "mix64_u8", "pole", (160, 90), dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1)),
    ("mip_nearest", "mix64_u8", "K1", (256, 256), dict(alpha_scale=0.9, min_val=0, max_val=255, filter=0, is_mip=1)),
    ("mip_trilinear_u16", "mix_64x64x32_u16", "K2", (256, 144), dict(alpha_scale=1.0, min_val=1000, max_val=3000, filter=1, is_mip=1)),
    ("view_top", "rand_48x40x36_u8", "K1", (200, 120), dict(alpha_scale=0.08, min_val=0, max_val=255, filter=1, view_top=1)),
    ("view_bottom", "rand_40x56x33_u16", "K0", (200, 120), dict(alpha_scale=0.05, min_val=0, max_val=4095, filter=0, view_bottom=1)),
    ("tf_default_knots", "mix64_u8", "K1", (256, 256), dict(alpha_scale=0.3, min_val=0, max_val=255, filter=1, tf=True)),
    ("tf_mip", "mix64_u8", "K0", (128, 128), dict(alpha_scale=0.9, min_val=0, max_val=255, filter=0, tf=True, is_mip=1)),
    ("window_min_eq_max_nan", "mix64_u8", "K0", (64, 64), dict(alpha_scale=0.5, min_val=100, max_val=100, filter=0)),
    ("window_min_gt_max", "mix64_u8", "K0", (64, 64), dict(alpha_scale=0.001, min_val=200, max_val=100, filter=1)),
    ("const_closed_form", "const_u8", "K0", (96, 96), dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1)),
    ("one_voxel", "one_voxel_u8", "K0", (64, 64), dict(alpha_scale=0.7, min_val=0, max_val=255, filter=1)),
    # VolumeRenderer.cs:115: the loop stops after 10000 samples (16 voxels / (step 1/16 * 0.001) = 16000 > 10000)
    ("iteration_cap_10000", "const_u8", "K0", (48, 48), dict(alpha_scale=0.00001, min_val=0, max_val=255, filter=1, step_scale=0.001)),
    # the same cap reached WITHOUT the step override (so the reference's own shader can run it): a 1x1x12000 rod seen end-on
    ("iteration_cap_reference", "rod_1x1x12000_u8", "K0", (40, 40), dict(alpha_scale=0.001, min_val=0, max_val=255, filter=0)),
    ("half_step_opacity_corrected", "smooth64_u8", "K1", (128, 128), dict(alpha_scale=0.1, min_val=0, max_val=255, filter=1, step_scale=0.5, opacity_correction=1)),
]

# cases whose result may legitimately differ from the oracle in the last bits (double pow on
# the GPU vs glibc); they are held to TOL only
TOLERANCE_ONLY = {"half_step_opacity_corrected"}


def is_reference_semantics(kw):
    """True when the case uses nothing VolumeRenderer.cs does not have (step override, TF, opacity correction):
    such a case can be run through the reference's own shader (oracle/_ref/libshader_ref.so)."""
    return kw.get("step_scale", 1.0) == 1.0 and not kw.get("tf", False) and not kw.get("opacity_correction", 0)


def case_by_id(cid):
    for c in CASES:
        if c[0] == cid:
            return c
    raise KeyError(cid)


def split_kwargs(kw):
    """-> (oracle kwargs, vr_params kwargs)"""
    kw = dict(kw)
    use_tf = kw.pop("tf", False)
    lut = tf_lut() if use_tf else None
    o = dict(kw)
    o["tf_lut"] = lut
    v = dict(kw)
    v["tf_lut"] = lut
    return o, v
