"""CPU: pins the march oracle (oracle/march_oracle.c) -- against hand-derived known answers
from the shader text, against an independent numpy restatement (tests/pyref.py)
and against the committed golden fixtures."""
import math
import os

import numpy as np
import pytest

from oracle import orc

import pyref
import scenarios
from util import oracle_frame

K0 = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 3, 1, 0, 0, 3, 1, 3.7320508], dtype=np.float32)


def const_volume(n, value, dtype=np.uint8):
    return np.full(n * n * n, value, dtype=dtype)


def test_constant_volume_closed_form():
    """VolumeRenderer.cs:130-132 closed form: A_n = 1-(1-v*a)^n, C_n = v*A_n."""
    n, s, alpha = 32, 100, 0.05
    W = H = 64
    img, cnt = oracle_frame(K0, const_volume(n, s), (n, n, n), 1, W, H, nthreads=1,
                            alpha_scale=alpha, min_val=0, max_val=255, filter=0)
    v = s / 255.0
    centre = img[H // 2, W // 2]
    # a central ray crosses the unit cube front to back: ~n samples (step = 1/n, :109)
    k = round(math.log(1 - centre[3]) / math.log(1 - v * alpha))
    assert abs(k - n) <= 1
    assert centre[3] == pytest.approx(1 - (1 - v * alpha) ** k, rel=1e-5)
    assert centre[0] == pytest.approx(v * centre[3], rel=1e-5)
    assert centre[0] == centre[1] == centre[2]


def test_reference_step_is_one_voxel_for_cubes():
    """step = |bbox diagonal| / |vol_size| = 1/N for an N^3 unit-spacing volume (:109)."""
    for n in (16, 40):
        W = H = 2           # four near-axial rays through the middle of the cube
        _, cnt = oracle_frame(K0, const_volume(n, 1), (n, n, n), 1, W, H, nthreads=1,
                              alpha_scale=0.0, min_val=0, max_val=255, filter=0)
        per_ray = cnt["samples"] / cnt["rays_hit"]
        assert n * 0.98 <= per_ray <= n * 1.08


def test_default_camera_coverage_and_miss_pixels():
    """K0: unit cube front face at depth 2.5, view_plane_dist 3.732 -> the cube spans 0.746 of
    the image height: ~31 % of a 16:9 frame is hit; misses store vec4(0) (:98-101)."""
    W, H = 320, 180
    img, cnt = oracle_frame(K0, const_volume(8, 255), (8, 8, 8), 1, W, H, nthreads=1,
                            alpha_scale=1.0, min_val=0, max_val=255, filter=0)
    assert 0.29 < cnt["rays_hit"] / cnt["rays"] < 0.33
    assert (img[0] == 0).all() and (img[:, 0] == 0).all()
    rows_hit = (img[..., 3] > 0).any(axis=1).sum()
    assert abs(rows_hit / H - 0.746) < 0.02


def test_ert_thresholds():
    """dest.a >= 0.95 (top of loop, :118) is the effective threshold; v*alpha = 1 -> one sample."""
    n = 16
    img, cnt = oracle_frame(K0, const_volume(n, 255), (n, n, n), 1, 32, 32, nthreads=1,
                            alpha_scale=1.0, min_val=0, max_val=255, filter=0)
    assert cnt["samples"] == cnt["rays_hit"]
    assert img[16, 16, 3] == 1.0 and img[16, 16, 0] == 1.0
    # a = 0.5 per sample: A = .5, .75, .875, .9375, .96875 -> 5 samples, stops at the 0.95 test
    img, cnt = oracle_frame(K0, const_volume(n, 255), (n, n, n), 1, 32, 32, nthreads=1,
                            alpha_scale=0.5, min_val=0, max_val=255, filter=0)
    assert img[16, 16, 3] == 0.96875
    assert 0.85 * 5 * cnt["rays_hit"] < cnt["samples"] <= 5 * cnt["rays_hit"]   # corner-clipping rays exit earlier


def test_mip_is_max_of_window_times_alpha_with_ert_quirk():
    n = 16
    vol = np.zeros((n, n, n), np.uint8)
    vol[:, :, :] = 50
    vol[5, :, :] = 200          # a bright slab somewhere along z
    img, _ = oracle_frame(K0, vol.reshape(-1), (n, n, n), 1, 32, 32, nthreads=1,
                          alpha_scale=0.5, min_val=0, max_val=255, filter=0, is_mip=1)
    assert img[16, 16, 3] == np.float32(np.float32(200) / np.float32(255)) * np.float32(0.5)
    assert (img[16, 16, :3] == img[16, 16, 3]).all()
    # with alpha 1 the MIP loop also stops once dest.a >= 0.95 (inherited test, :156)
    vol[:] = 250
    img, cnt = oracle_frame(K0, vol.reshape(-1), (n, n, n), 1, 32, 32, nthreads=1,
                            alpha_scale=1.0, min_val=0, max_val=255, filter=0, is_mip=1)
    assert cnt["samples"] == cnt["rays_hit"]


def test_window_equal_bounds_gives_nan_like_the_shader():
    n = 8
    img, _ = oracle_frame(K0, const_volume(n, 7), (n, n, n), 1, 16, 16, nthreads=1,
                          alpha_scale=1.0, min_val=7, max_val=7, filter=0)
    assert np.isnan(img[8, 8]).all()      # 0/0 at :124, no guard in the shader


def test_file_slice_zero_faces_the_default_camera():
    """point.z = 1 - point.z (:183): file slice k = 0 is the one nearest the default camera."""
    n = 8
    vol = np.zeros((n, n, n), np.uint8)
    vol[0] = 255
    img, _ = oracle_frame(K0, vol.reshape(-1), (n, n, n), 1, 32, 32, nthreads=1,
                          alpha_scale=1.0, min_val=0, max_val=255, filter=0)
    assert img[16, 16, 3] == 1.0          # first sample already opaque
    vol[:] = 0
    vol[n - 1] = 255
    _, cnt = oracle_frame(K0, vol.reshape(-1), (n, n, n), 1, 32, 32, nthreads=1,
                          alpha_scale=1.0, min_val=0, max_val=255, filter=0)
    assert cnt["samples"] > 6 * cnt["rays_hit"]     # has to march to the back


def test_trilinear_equals_nearest_on_constant_data_and_interpolates_ramps():
    n = 16
    a, _ = oracle_frame(K0, const_volume(n, 90), (n, n, n), 1, 48, 48, nthreads=1, alpha_scale=0.1, min_val=0, max_val=255, filter=0)
    b, _ = oracle_frame(K0, const_volume(n, 90), (n, n, n), 1, 48, 48, nthreads=1, alpha_scale=0.1, min_val=0, max_val=255, filter=1)
    assert np.array_equal(a, b)
    ramp = np.tile(np.arange(n, dtype=np.uint8) * 10, n * n)       # value = 10*x
    t, _ = oracle_frame(K0, ramp, (n, n, n), 1, 64, 64, nthreads=1, alpha_scale=0.02, min_val=0, max_val=255, filter=1)
    row = t[32, 18:46, 0]      # columns whose rays cross the full depth of the cube
    assert (np.diff(row) >= -1e-7).all() and row[-1] > row[0]      # smooth and monotone in x


def test_multithreaded_equals_scalar():
    vox, dims, bpv, vs = scenarios.volume("rand_48x40x36_u8")
    cam = scenarios.camera("K1")
    kw = dict(alpha_scale=0.08, min_val=10, max_val=250, filter=1, voxel_size=vs)
    a, ca = oracle_frame(cam, vox, dims, bpv, 120, 70, nthreads=1, **kw)
    b, cb = oracle_frame(cam, vox, dims, bpv, 120, 70, nthreads=5, **kw)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and ca == cb


def test_row_subset_rendering():
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K1")
    full, _ = oracle_frame(cam, vox, dims, bpv, 64, 48, nthreads=2, alpha_scale=0.05, min_val=0, max_val=255, filter=1)
    p = orc.make_params(64, 48, dims, bpv, cam, alpha_scale=0.05, min_val=0, max_val=255, filter=1, row_begin=3, row_stride=7)
    part, _, _ = orc.render(p, vox, nthreads=2)
    rows = np.arange(3, 48, 7)
    assert np.array_equal(part[rows], full[rows])
    mask = np.ones(48, bool); mask[rows] = False
    assert (part[mask] == 0).all()


def test_touch_bitmap_counts_distinct_voxels():
    n = 8
    p = orc.make_params(64, 64, (n, n, n), 1, K0, alpha_scale=0.0, min_val=0, max_val=255, filter=1)
    _, cnt, tb = orc.render(p, const_volume(n, 1), nthreads=1, touch=True)
    assert orc.popcount(tb, n ** 3) == n ** 3          # every voxel is seen at this resolution
    p = orc.make_params(2, 2, (n, n, n), 1, K0, alpha_scale=0.0, min_val=0, max_val=255, filter=0)
    _, cnt, tb = orc.render(p, const_volume(n, 1), nthreads=1, touch=True)
    assert 0 < orc.popcount(tb, n ** 3) <= cnt["samples"]


PYREF_CASES = [
    ("c1_nearest_ref_step", [(128, 128), (70, 90), (40, 200), (3, 3)]),
    ("ragged_trilinear", [(125, 65), (60, 40), (200, 100)]),
    ("u16_ragged_trilinear", [(120, 67), (30, 20)]),
    ("mip_nearest", [(128, 128), (100, 60)]),
    ("view_top", [(100, 60), (80, 30)]),
    ("view_bottom", [(100, 60), (120, 90)]),
    ("tf_default_knots", [(128, 128), (90, 170)]),
    ("window_min_gt_max", [(32, 32)]),
    ("eye_inside", [(80, 45), (10, 10)]),
    ("u16_aniso_trilinear", [(128, 72), (60, 100)]),
    ("mip_trilinear_u16", [(128, 72), (200, 40)]),
    ("pole_camera", [(80, 45), (30, 70)]),
    ("c1_trilinear_window_ert", [(128, 128), (180, 90)]),
    ("u16_full_range", [(64, 64)]),
    ("iteration_cap_10000", [(24, 24)]),
]


@pytest.mark.parametrize("cid,pixels", PYREF_CASES)
def test_c_oracle_equals_independent_numpy_restatement(cid, pixels):
    _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    okw, _ = scenarios.split_kwargs(kw)
    ref, _ = oracle_frame(cam, vox, dims, bpv, W, H, nthreads=4, voxel_size=vs, **okw)
    vox3d = vox.reshape(dims[2], dims[1], dims[0])
    for (px, py) in pixels:
        got = pyref.shade_pixel(px, py, W, H, dims, vox3d, cam, voxel_size=vs,
                                alpha_scale=okw.get("alpha_scale", 1.0), min_val=okw["min_val"], max_val=okw["max_val"],
                                is_mip=okw.get("is_mip", 0), view_top=okw.get("view_top", 0), view_bottom=okw.get("view_bottom", 0),
                                trilinear=okw.get("filter", 0) == 1, step_scale=okw.get("step_scale", 1.0), tf_lut=okw.get("tf_lut"))
        exp = ref[py, px]
        assert np.array_equal(np.array(got, np.float32).view(np.uint32), exp.view(np.uint32)), (cid, px, py, got, exp)


@pytest.mark.parametrize("cid", ["c1_trilinear_128steps", "ragged_nearest", "u16_aniso_trilinear", "mip_nearest"])
def test_oracle_reproduces_golden_fixtures(cid, golden_dir):
    g = np.load(os.path.join(golden_dir, f"{cid}.npz"))
    _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    assert np.array_equal(cam.view(np.uint32), g["cam"].view(np.uint32))     # host Camera is stable too
    okw, _ = scenarios.split_kwargs(kw)
    img, cnt = oracle_frame(cam, vox, dims, bpv, W, H, nthreads=4, voxel_size=vs, **okw)
    assert np.array_equal(img.view(np.uint32), g["rgba"].view(np.uint32))
    assert cnt["samples"] == int(g["samples"]) and cnt["rays_hit"] == int(g["rays_hit"])
