"""Pins the oracle (and the host mirror) to the REFERENCE'S OWN SOURCES, compiled where they lie under
/root/reference into oracle/_ref/ (oracle/Makefile):

  libshader_ref.so  VolumeRenderer.cs -- the GLSL compute shader text itself, compiled by g++ through
                    oracle/shim/glsl_compat.h and run once per pixel like glDispatchCompute does
  libhost_ref.so    src/Camera.cpp + src/CubicSpline.cpp, unmodified, against the GLM stand-in oracle/shim/glm

Every comparison below is bit for bit.  The .so files travel to the GPU box with the snapshot; when
oracle/_ref was never built (a checkout without the reference tree) the tests that need it skip, and the
committed fixtures tests/golden/ref_*.npz (written by tests/golden/make_ref_golden.py from libshader_ref.so)
still pin the oracle."""
import math
import os

import numpy as np
import pytest

from oracle import orc
from volren_b200 import host

import scenarios

needs_shader_ref = pytest.mark.skipif(orc.ref_shader_lib() is None, reason="oracle/_ref/libshader_ref.so not built")
needs_host_ref = pytest.mark.skipif(orc.ref_host_lib() is None, reason="oracle/_ref/libhost_ref.so not built")


def bits_equal(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


# ------------------------------------------------------------------ march: VolumeRenderer.cs

@needs_shader_ref
@pytest.mark.parametrize("cid", [c[0] for c in scenarios.CASES if scenarios.is_reference_semantics(c[4])])
def test_oracle_equals_the_reference_shader_on_every_scenario(cid):
    _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    okw, _ = scenarios.split_kwargs(kw)
    p = orc.make_params(W, H, dims, bpv, cam, voxel_size=vs, **okw)
    exp, rcnt = orc.ref_render(p, vox, nthreads=4)
    got, cnt, _ = orc.render(p, vox, nthreads=4)
    assert bits_equal(got, exp), f"{cid}: {(got.view(np.uint32) != exp.view(np.uint32)).sum()} values differ"
    assert cnt["samples"] == rcnt["samples"] and cnt["rays"] == rcnt["rays"]


@needs_shader_ref
def test_extension_scenarios_are_refused_by_the_reference_shader():
    """step override / TF / opacity correction do not exist in VolumeRenderer.cs: they are pinned only through
    the identity (step_scale 1, no LUT) cases above."""
    ext = [c for c in scenarios.CASES if not scenarios.is_reference_semantics(c[4])]
    assert {c[0] for c in ext} == {"c1_trilinear_128steps", "tf_default_knots", "tf_mip", "iteration_cap_10000",
                                   "half_step_opacity_corrected"}
    _, vname, cname, (W, H), kw = ext[0]
    vox, dims, bpv, vs = scenarios.volume(vname)
    okw, _ = scenarios.split_kwargs(kw)
    with pytest.raises(ValueError):
        orc.ref_render(orc.make_params(W, H, dims, bpv, scenarios.camera(cname), voxel_size=vs, **okw), vox)


@needs_shader_ref
def test_oracle_equals_the_reference_shader_on_random_frames():
    """60 random volumes / spacings / orbit cameras / windows / modes, 40x28 pixels each."""
    rng = np.random.default_rng(2026)
    for trial in range(60):
        dims = tuple(int(v) for v in rng.integers(1, 40, 3))
        bpv = int(rng.integers(1, 3))
        vmax = 255 if bpv == 1 else int(rng.choice([4095, 65535]))
        vox = rng.integers(0, vmax + 1, dims[0] * dims[1] * dims[2]).astype(np.uint8 if bpv == 1 else np.uint16)
        vs = tuple(float(np.float32(v)) for v in rng.uniform(0.4, 2.5, 3)) if trial % 3 else (1.0, 1.0, 1.0)
        cam = host.Camera(30.0)
        cam.setSpherical(float(rng.uniform(0.2, 4.0)), float(rng.uniform(0.0, math.pi)), float(rng.uniform(0.0, 2 * math.pi)))
        lo, hi = sorted(int(v) for v in rng.integers(0, vmax + 1, 2))
        if trial % 7 == 0:
            lo, hi = hi, lo                      # unordered window (VolumeRenderer.cs:123 false)
        if trial % 11 == 0:
            hi = lo                              # 0/0
        view = int(rng.integers(0, 3))
        p = orc.make_params(40, 28, dims, bpv, cam.ubo(), voxel_size=vs,
                            alpha_scale=float(np.float32(rng.choice([0.01, 0.1, 1.0, 3.0]))), min_val=lo, max_val=hi,
                            is_mip=int(rng.integers(0, 2)), view_top=int(view == 1), view_bottom=int(view == 2),
                            filter=int(rng.integers(0, 2)))
        exp, rcnt = orc.ref_render(p, vox)
        got, cnt, _ = orc.render(p, vox)
        assert bits_equal(got, exp), (trial, dims, bpv, vs, lo, hi, view)
        assert cnt["samples"] == rcnt["samples"]


@needs_shader_ref
def test_multithreaded_reference_equals_scalar_and_row_subsets():
    _, vname, cname, (W, H), kw = scenarios.case_by_id("ragged_trilinear")
    vox, dims, bpv, vs = scenarios.volume(vname)
    okw, _ = scenarios.split_kwargs(kw)
    cam = scenarios.camera(cname)
    a, _ = orc.ref_render(orc.make_params(W, H, dims, bpv, cam, voxel_size=vs, **okw), vox, nthreads=1)
    b, _ = orc.ref_render(orc.make_params(W, H, dims, bpv, cam, voxel_size=vs, **okw), vox, nthreads=5)
    assert bits_equal(a, b)
    c, cnt = orc.ref_render(orc.make_params(W, H, dims, bpv, cam, voxel_size=vs, row_begin=3, row_end=90, row_stride=7, **okw), vox, nthreads=3)
    rows = list(range(3, 90, 7))
    assert bits_equal(c[rows], a[rows]) and cnt["rays"] == len(rows) * W
    mask = np.ones(H, bool); mask[rows] = False
    assert not c[mask].any()


REF_GOLDEN = ["c1_nearest_ref_step", "ragged_trilinear", "u16_aniso_trilinear", "mip_nearest", "view_top",
              "window_min_gt_max", "iteration_cap_reference"]


@pytest.mark.parametrize("cid", REF_GOLDEN)
def test_oracle_reproduces_fixtures_written_by_the_reference_shader(cid, golden_dir):
    """tests/golden/ref_<cid>.npz: output of the reference's own shader (libshader_ref.so), committed so the pin
    also holds where oracle/_ref does not exist."""
    g = np.load(os.path.join(golden_dir, f"ref_{cid}.npz"))
    _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    assert np.array_equal(cam.view(np.uint32), g["cam"].view(np.uint32))
    okw, _ = scenarios.split_kwargs(kw)
    img, cnt, _ = orc.render(orc.make_params(W, H, dims, bpv, cam, voxel_size=vs, **okw), vox, nthreads=4)
    assert bits_equal(img, g["rgba"])
    assert cnt["samples"] == int(g["samples"])


# ------------------------------------------------------------------ Camera: src/Camera.cpp

@needs_host_ref
def test_reference_camera_reset_ubo_and_flags():
    r = orc.RefCamera(30.0)
    ubo, eye, side, up, look, before, after = r.state()
    assert ubo.tolist()[:20] == [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 3, 1, 0, 0, 3, 1]
    assert ubo[20] == pytest.approx(3.7320508, rel=1e-6)
    assert before and after                   # Camera.cpp:63-72 returns before :79 clears is_changed
    assert bits_equal(ubo, orc.OracleCamera(30.0).ubo()) and bits_equal(ubo, host.Camera(30.0).ubo())


@needs_host_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_and_host_camera_equal_the_reference_camera_over_random_walks(seed):
    rng = np.random.default_rng(seed)
    r, o, h = orc.RefCamera(30.0), orc.OracleCamera(30.0), host.Camera(30.0)
    for i in range(600):
        u = rng.random()
        if u < 0.12:
            z, a, b = (1 if rng.random() < 0.5 else -1), 0.0, 0.0
        elif u < 0.16:
            r.reset(); o.reset(); h.resetCamera(); continue
        elif u < 0.22:
            z, a, b = 0, float(rng.choice([-3.0, 3.0])), float(rng.uniform(-0.4, 0.4))     # drives zenith into the clamp -> pole branch
        else:
            z, a, b = 0, float(rng.uniform(-0.4, 0.4)), float(rng.uniform(-9.5, 9.5) if u > 0.9 else rng.uniform(-0.4, 0.4))
        r.set_orientation(z, a, b); o.set_orientation(z, a, b); h.setOrientation(z, a, b)
        ubo, eye, side, up, look, _, changed = r.state()
        assert bits_equal(o.ubo(), ubo), (seed, i)
        assert bits_equal(h.ubo(), ubo), (seed, i)
        assert bits_equal(np.array(o.c.eye[:]), eye) and bits_equal(np.array(o.c.side[:]), side)
        assert bits_equal(np.array(o.c.up[:]), up) and bits_equal(np.array(o.c.look_at[:]), look)
        assert changed


# ------------------------------------------------------------------ CubicSpline: src/CubicSpline.cpp

@needs_host_ref
def test_reference_spline_known_answers_of_the_default_knots():
    s = orc.RefSpline(scenarios.DEFAULT_KNOTS)
    exp = {1: 0.007808861, 64: 0.46778646, 128: 0.74364215, 140: 0.7584175, 145: 0.620175, 200: 0.6007686, 254: 0.9919789}
    for iso, a in exp.items():
        assert s.eval_iso(iso)[3] == pytest.approx(a, rel=2e-6)


@needs_host_ref
def test_oracle_and_host_spline_equal_the_reference_spline():
    rng = np.random.default_rng(5)
    for trial in range(30):
        n = int(rng.integers(2, 10))
        isos = sorted(rng.choice(np.arange(1, 255), n - 2, replace=False).tolist()) if n > 2 else []
        knots = ([(0, tuple(rng.random(4).tolist()))] + [(int(i), tuple(rng.uniform(-0.5, 1.5, 4).tolist())) for i in isos] +
                 [(255, tuple(rng.random(4).tolist()))])
        r, o, h = orc.RefSpline(knots), orc.OracleSpline(knots), host.CubicSpline(knots)
        for iso in range(256):
            e = r.eval_iso(iso)
            assert bits_equal(o.eval_iso(iso), e) and bits_equal(h.getPointOnSpline(iso), e), (trial, iso)
        for seg in range(n - 1):
            for t in (0.0, 0.1, 0.5, 0.99, 1.0):
                e = r.eval_t(t, seg)
                assert bits_equal(o.eval_t(t, seg), e) and bits_equal(h.getPointOnSplineT(t, seg), e)
        # the 256-entry opacity LUT the kernel reads = clamp(reference spline .w, 0, 1) (AlphaControlSplineWidget.cpp:247)
        lut = np.array([min(max(r.eval_iso(i)[3], np.float32(0)), np.float32(1)) for i in range(256)], np.float32)
        assert bits_equal(h.bakeAlphaLUT(), lut) and bits_equal(o.alpha_lut(), lut)
