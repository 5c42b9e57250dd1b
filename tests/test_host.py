"""CPU: the C++ host mirror of the reference surface (volume-renderer_b200/host/): Camera,
CubicSpline, .raw.inf sidecar, image writers -- against hand-derived known answers
 and, bit for bit, against the oracle restatements."""
import math
import os
import struct
import zlib

import numpy as np
import pytest

from oracle import orc
from volren_b200 import host

DEFAULT_KNOTS = [(0, 0.0), (141, 0.759), (149, 0.45), (255, 1.0)]


# ------------------------------------------------------------------ Camera

def test_reset_camera_ubo_known_answer():
    cam = host.Camera(30.0)
    # Camera.cpp:34-38,56,59-80,19: [side|up|-look|eye], eye xyz1, 1/tan(15 deg)
    exp = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 3, 1, 0, 0, 3, 1, 3.7320508]
    assert cam.ubo() == pytest.approx(exp, abs=1e-6)
    assert cam.ubo()[20] == pytest.approx(1.0 / math.tan(math.radians(15.0)), rel=3e-7)


def test_zoom_moves_eye_along_look_at_in_unit_steps():
    cam = host.Camera(30.0)
    cam.setOrientation(1, 0, 0)        # scroll up: eye += look_at
    u = cam.ubo()
    assert u[16:19] == pytest.approx([0, 0, 2]) and u[12:15] == pytest.approx([0, 0, 2])
    cam.setOrientation(-1, 0, 0)
    cam.setOrientation(-1, 0, 0)
    assert cam.ubo()[16:19] == pytest.approx([0, 0, 4])


def test_orbit_matches_spherical_formulas():
    cam = host.Camera(30.0)
    dz, da = -0.3, 0.5
    cam.setOrientation(0, dz, da)      # rotation speed 0.7 (Camera.h:14)
    zen = np.float32(np.float32(np.pi) / 2) + np.float32(dz) * np.float32(0.7)
    az = np.float32(da) * np.float32(0.7)
    eye = 3 * np.array([math.sin(zen) * math.sin(az), math.cos(zen), math.sin(zen) * math.cos(az)])
    u = cam.ubo()
    assert u[16:19] == pytest.approx(eye, abs=1e-5)
    side, up, back = u[0:3], u[4:7], u[8:11]
    assert np.dot(side, up) == pytest.approx(0, abs=1e-6) and np.linalg.norm(side) == pytest.approx(1, abs=1e-6)
    assert back == pytest.approx(eye / 3, abs=1e-6)          # third column is -look_at = eye/|eye|
    assert side[1] == 0                                       # side stays horizontal


def test_azimuth_wrap_quirk_and_zenith_clamp():
    cam = host.Camera(30.0)
    o = orc.OracleCamera(30.0)
    for c in (cam, o):
        (c.setOrientation if c is cam else c.set_orientation)(0, 0.0, -0.1)     # negative azimuth -> 2*pi - a (sic)
    assert np.array_equal(cam.ubo().view(np.uint32), o.ubo().view(np.uint32))
    az = np.float32(2 * np.float32(np.pi)) - np.float32(np.float32(-0.1) * np.float32(0.7))
    assert cam.ubo()[16] == pytest.approx(3 * math.sin(az), abs=1e-5)
    for _ in range(40):
        cam.setOrientation(0, -0.2, 0.0)                    # zenith clamps at 0 -> pole branch
        o.set_orientation(0, -0.2, 0.0)
    assert cam.ubo()[17] == pytest.approx(3.0)
    assert np.array_equal(cam.ubo().view(np.uint32), o.ubo().view(np.uint32))


def test_camera_bitwise_equals_oracle_over_random_walk():
    rng = np.random.default_rng(1)
    cam, o = host.Camera(30.0), orc.OracleCamera(30.0)
    for i in range(500):
        r = rng.random()
        if r < 0.15:
            z, a, b = (1 if rng.random() < 0.5 else -1), 0.0, 0.0
        elif r < 0.2:
            cam.resetCamera(); o.reset(); continue
        else:
            z, a, b = 0, float(rng.uniform(-0.4, 0.4)), float(rng.uniform(-0.4, 0.4))
        cam.setOrientation(z, a, b); o.set_orientation(z, a, b)
        a, b = cam.ubo(), o.ubo()
        # zooming the eye onto the origin yields NaNs in the reference too (normalize(0)); the
        # sign of a NaN is not part of the contract
        assert (((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all()), i


def test_is_changed_is_never_cleared_like_the_reference():
    cam = host.Camera(30.0)
    assert cam.is_changed
    cam.ubo()
    assert cam.is_changed          # Camera.cpp:63-72 returns before :79


# ------------------------------------------------------------------ CubicSpline

def test_default_knots_known_answers():
    """SURVEY.md 8c, fp32 evaluation of CubicSpline.cpp:50-115 on AlphaControlSplineWidget.cpp:56-59."""
    s = host.CubicSpline(DEFAULT_KNOTS)
    exp = {1: 0.007808861, 64: 0.46778646, 128: 0.74364215, 140: 0.7584175, 145: 0.620175, 200: 0.6007686, 254: 0.9919789}
    for iso, a in exp.items():
        assert s.getPointOnSpline(iso)[3] == pytest.approx(a, rel=2e-6)
    for iso, a in DEFAULT_KNOTS:
        assert s.getPointOnSpline(iso)[3] == np.float32(a)     # exact knots return the knot (:27-28)
    o = orc.OracleSpline(DEFAULT_KNOTS)
    assert o.field("coeffs")[:, 3] == pytest.approx([0.5, 0.2857143, 0.26923078, 0.5777778], rel=1e-6)
    assert o.field("deriv")[:, 3] == pytest.approx([1.1010666, 0.07486665, -0.050533354, 0.85026675], rel=1e-5)
    assert o.field("d")[:3, 3] == pytest.approx([-0.34206676, 0.6423333, -0.30026668], rel=1e-5)


def test_spline_bitwise_equals_oracle():
    rng = np.random.default_rng(2)
    for trial in range(20):
        n = int(rng.integers(2, 9))
        isos = sorted(rng.choice(np.arange(1, 255), n - 2, replace=False).tolist()) if n > 2 else []
        knots = [(0, tuple(rng.random(4).tolist()))] + [(int(i), tuple(rng.random(4).tolist())) for i in isos] + [(255, tuple(rng.random(4).tolist()))]
        s, o = host.CubicSpline(knots), orc.OracleSpline(knots)
        for iso in range(256):
            assert np.array_equal(s.getPointOnSpline(iso).view(np.uint32), o.eval_iso(iso).view(np.uint32))
        for seg in range(n - 1):
            for t in (0.0, 0.25, 0.5, 1.0):
                assert np.array_equal(s.getPointOnSplineT(t, seg).view(np.uint32), o.eval_t(t, seg).view(np.uint32))
        assert np.array_equal(s.bakeAlphaLUT().view(np.uint32), o.alpha_lut().view(np.uint32))


def test_alpha_lut_clamps_and_covers_partial_knot_ranges():
    lut = host.CubicSpline([(20, 0.2), (100, 1.4), (200, -0.3)]).bakeAlphaLUT()
    assert lut.min() >= 0 and lut.max() <= 1
    assert (lut[:21] == np.float32(0.2)).all()       # below the first knot: clamped to it
    assert (lut[200:] == 0).all()                    # above the last knot: clamped (-0.3 -> 0)


# ------------------------------------------------------------------ .raw.inf

def test_rawinf_text_format_and_round_trip(tmp_path):
    fn = str(tmp_path / "vol.raw")
    assert host.rawinf_write(fn, (64, 48, 40), (1.0, 1.5, 2.0))
    # RendererCore.cpp:311-315
    assert open(fn + ".inf").read() == "#dimensions\n64 48 40\n\n#voxel-spacing\n1 1.5 2\n"
    rc, dims, sp, _, _ = host.rawinf_read(fn)
    assert rc == 1 and dims == (64, 48, 40) and sp == (1.0, 1.5, 2.0)
    assert host.rawinf_read(str(tmp_path / "missing.raw"))[0] == -1


@pytest.mark.parametrize("text,title,frag", [
    ("#dimensions\n\n#voxel-spacing\n1 1 1\n", "Invalid .raw.inf file!", "Dimensions for Volume Data not provided"),
    ("#dimensions\n4 4 4\n#voxel-spacing\n\n", "Invalid .raw.inf file!", "Aspect Ratio for Volume Data not provided"),
    ("#voxel-spacing\n1 1 1\n", "Invalid .raw.inf file!", "#dimesnsions"),
    ("#dimensions\n4 4 4\n", "Invalid .raw.inf file!", "#voxel-spacing"),
])
def test_rawinf_errors_use_the_reference_messages(tmp_path, text, title, frag):
    fn = str(tmp_path / "v.raw")
    open(fn + ".inf", "w").write(text)
    rc, _, _, t, m = host.rawinf_read(fn)
    assert rc == 0 and t == title and frag in m          # RendererCore.cpp:264-301


# ------------------------------------------------------------------ image writers

def _png_pixels(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + body) == crc
        if typ == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 2)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    data = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    assert (data[:, 0] == 0).all()
    return data[:, 1:].reshape(h, w, 3)


def test_png_bmp_ppm_writers(tmp_path):
    rng = np.random.default_rng(4)
    for (h, w) in ((5, 7), (300, 251)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        p = str(tmp_path / "a.png")
        assert host.write_image(p, ".png", img)
        assert np.array_equal(_png_pixels(p), img)
        p = str(tmp_path / "a.ppm")
        assert host.write_image(p, ".ppm", img)
        raw = open(p, "rb").read()
        hdr = f"P6\n{w} {h}\n255\n".encode()
        assert raw.startswith(hdr) and raw[len(hdr):] == img.tobytes()
        p = str(tmp_path / "a.bmp")
        assert host.write_image(p, ".bmp", img)
        raw = open(p, "rb").read()
        assert raw[:2] == b"BM" and struct.unpack("<ii", raw[18:26]) == (w, h)
        stride = (w * 3 + 3) & ~3
        rows = np.frombuffer(raw[54:], np.uint8).reshape(h, stride)[:, :w * 3].reshape(h, w, 3)
        assert np.array_equal(rows[::-1, :, ::-1], img)      # bottom-up, BGR
    assert not host.write_image(str(tmp_path / "a.tga"), ".tga", img)     # not a format of RendererCore::saveImage


def test_headless_example_compiles_against_the_host_mirror(tmp_path):
    """examples/headless_render.cpp drives RendererCore / Camera with the reference GUI's call sequence;
    it must compile and link with the host sources built into the application.
    Without a GPU it fails the way the reference does when its framebuffer is incomplete: a runtime_error
    caught in main, exit status 1."""
    import subprocess
    import volren_b200 as vb
    root = vb.REPO_ROOT
    h = os.path.join(root, "volume-renderer_b200", "host")
    exe = str(tmp_path / "headless_render")
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I" + os.path.join(root, "include"), "-I" + h,
           os.path.join(root, "examples", "headless_render.cpp")] + \
          [os.path.join(h, f) for f in ("RendererCore.cpp", "Camera.cpp", "CubicSpline.cpp", "VolumeIO.cpp", "ImageIO.cpp")] + \
          ["-L" + vb.LIB_DIR, "-lvolren_b200", "-Wl,-rpath," + vb.LIB_DIR, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 2 and "usage:" in run.stderr


def test_c_example_compiles_against_the_c_abi(tmp_path):
    """examples/pipelined_frames.c: the drop-in boundary is a C ABI -- the header must compile as plain C11 (no C++
    constructs), and the pipelined entry points (vr_render_submit / vr_render_wait) must link.  Without a GPU the
    program fails loudly on its first compute call: status 1 and the library's error text, never a silent CPU path."""
    import subprocess
    import volren_b200 as vb
    root = vb.REPO_ROOT
    exe = str(tmp_path / "pipelined_frames")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"),
           os.path.join(root, "examples", "pipelined_frames.c"), "-L" + vb.LIB_DIR, "-lvolren_b200",
           "-Wl,-rpath," + vb.LIB_DIR, "-lm", "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    run = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert run.returncode == 2 and "usage:" in run.stderr
    import torch
    if not torch.cuda.is_available():
        run = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=120)
        assert run.returncode == 1 and "vr_create" in run.stderr


@pytest.mark.parametrize("w,h,kind", [(64, 48, "smooth"), (131, 67, "smooth"), (17, 9, "noise"), (1, 1, "noise"), (250, 131, "noise")])
def test_jpg_writer_decodes_with_independent_decoders(tmp_path, w, h, kind):
    """saveImage(".jpg") (RendererCore.cpp:175-176: stb quality 100): baseline JFIF, 4:4:4, unit quantiser,
    image-optimised Huffman tables.  Pillow and OpenCV (two independent libjpeg-class decoders) must both
    reproduce the pixels to within the rounding of an 8-bit DCT round trip."""
    PIL_Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(w * 1000 + h)
    if kind == "noise":
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    else:
        y, x = np.mgrid[0:h, 0:w]
        img = np.stack([127 + 120 * np.sin(x / 9.0), 127 + 120 * np.cos(y / 7.0), 127 + 100 * np.sin((x + y) / 11.0)], -1).astype(np.uint8)
    fn = str(tmp_path / "t.jpg")
    assert host.write_image(fn, ".jpg", np.ascontiguousarray(img))
    got = np.asarray(PIL_Image.open(fn).convert("RGB"))
    assert got.shape == img.shape
    assert np.abs(got.astype(int) - img.astype(int)).max() <= 4
    try:
        import cv2
    except ImportError:
        return
    got2 = cv2.cvtColor(cv2.imread(fn), cv2.COLOR_BGR2RGB)
    assert np.abs(got2.astype(int) - img.astype(int)).max() <= 4


def test_frame_constants_match_the_oracle_bit_for_bit():
    """The per-frame constants (bounding box :62-83, tex-coord denominators :179, step :109/:146, window
    floats :122-124) are evaluated on the HOST in the product (csrc/frame.h) and once more in the oracle:
    every one of them must carry the same bits for arbitrary dims / spacings / view flags / windows /
    step scales, or no kernel could be bit-exact for that volume."""
    import volren_b200 as vb
    rng = np.random.default_rng(2026)
    cam = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 3, 1, 0, 0, 3, 1, 3.7320508], dtype=np.float32)
    for it in range(400):
        if it % 4 == 0:
            dims = tuple(int(2 ** rng.integers(0, 12)) for _ in range(3))
            vs = tuple(float(2.0 ** rng.integers(-2, 3)) for _ in range(3))
        else:
            dims = tuple(int(x) for x in rng.integers(1, 2049, 3))
            vs = tuple(float(np.float32(x)) for x in rng.uniform(0.2, 4.0, 3))
        kw = dict(alpha_scale=float(np.float32(rng.uniform(0, 1))), min_val=int(rng.integers(-50, 3000)),
                  max_val=int(rng.integers(-50, 70000)), is_mip=int(rng.integers(0, 2)),
                  view_top=int(rng.integers(0, 2)), view_bottom=int(rng.integers(0, 2)),
                  step_scale=float(np.float32(rng.choice([1.0, 0.5, 0.25, 2.0, rng.uniform(0.05, 3.0)]))))
        mine = host.frame_consts(640, 480, dims, vs, cam, vb.default_params(filter=1, **kw))
        ref = orc.frame_consts(orc.make_params(640, 480, dims, 2, cam, voxel_size=vs, filter=1, **kw))
        what = f"dims {dims} spacing {vs} {kw}"
        assert np.array_equal(mine[:12].view(np.uint32), ref[:12].view(np.uint32)), what      # pmin, pmax, half_len, denom
        step_ref = ref[13] if kw["is_mip"] else ref[12]
        assert mine[12:13].view(np.uint32)[0] == np.float32(step_ref).view(np.uint32), what
        assert np.array_equal(mine[13:16].view(np.uint32), ref[14:17].view(np.uint32)), what   # fmin, fmax, frange
        # derived values the kernels rely on: correctly rounded reciprocals, division mode
        assert np.array_equal(mine[16:19], (np.float32(1.0) / mine[9:12]).astype(np.float32)), what
        pow2 = all(float(d) > 0 and np.frexp(float(d))[0] == 0.5 for d in mine[9:12])
        assert int(mine[20]) == (0 if pow2 else 1), what


# ------------------------------------------------------------------ .raw mmap path, 64-bit sizes, untrusted headers

def test_raw_payload_is_memory_mapped_with_64_bit_offsets(tmp_path):
    """the reference reads a .raw through `int len` (RendererCore.cpp:327) and fails above 2^31 voxels;
    here the payload is mapped read-only.  A sparse 5 GiB file costs no disk space; bytes beyond 4 GiB are reachable."""
    fn = str(tmp_path / "big.raw")
    size = 5 * (1 << 30) + 123
    with open(fn, "wb") as f:
        f.truncate(size)
        f.seek((1 << 32) + 17); f.write(b"\xAB")
        f.seek(size - 1); f.write(b"\xCD")
    m = host.MappedRaw(fn)
    assert m.size == size
    assert m.byte((1 << 32) + 17) == 0xAB and m.byte(size - 1) == 0xCD and m.byte(12345) == 0
    assert m.byte(size) == -1
    with pytest.raises(OSError):
        host.MappedRaw(str(tmp_path / "missing.raw"))


def test_volume_byte_counts_are_overflow_checked():
    """untrusted dimensions are range checked before they are multiplied."""
    assert host.checked_volume_bytes((2048, 2048, 1024), 2) == 8 << 30
    assert host.checked_volume_bytes((16384, 16384, 16384), 2) == 1 << 43
    for bad in ((0, 4, 4), (16385, 1, 1), (1 << 31, 1 << 31, 4), (2 ** 32 - 1, 2 ** 32 - 1, 2 ** 32 - 1)):
        assert host.checked_volume_bytes(bad, 2) is None
    assert host.checked_volume_bytes((4, 4, 4), 0) is None


def test_pvm_header_with_hostile_dimensions_is_rejected():
    """A PVM header whose width*height*depth*components wraps 64 bits must not pass the payload-length check."""
    for dims in (b"4294967295 4294967295 4294967295", b"2147483648 2147483648 4", b"16385 1 1"):
        blob = b"PVM\n" + dims + b"\n1\n" + b"\x00" * 64
        p = host.pvm_decode(data=blob)
        assert not p["ok"], dims
    assert "16384" in host.pvm_decode(data=b"PVM\n16385 1 1\n1\n" + b"\x00" * 64)["error"]


def test_read_volume_data_reports_bad_inf_dimensions_without_throwing(tmp_path):
    """A .raw.inf with absurd dimensions must end in a popup (title/msg), not in bad_alloc through the C boundary."""
    fn = str(tmp_path / "v.raw")
    open(fn, "wb").write(b"\x00" * 64)
    open(fn + ".inf", "w").write("#dimensions\n2000000000 2000000000 2000000000\n\n#voxel-spacing\n1 1 1\n")
    core = host.RendererCore(64, 64)
    core.set_datasize_bytes(1)
    core.readVolumeData(fn)                  # no GPU needed: fails before any upload
    st = core.strings()
    assert st["title"] == "Invalid Data Size!" and "16384" in st["msg"]


def test_stored_camera_blocks_equal_the_host_camera():
    """workloads.camera_block_const (used by bench.py --impl reference, which must not load the product libraries)."""
    from volren_b200 import workloads
    for k in ("K0", "K1", "K2"):
        assert np.array_equal(workloads.camera_block_const(k).view(np.uint32), workloads.camera_block(k).view(np.uint32))
