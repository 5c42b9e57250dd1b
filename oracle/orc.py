"""ctypes bindings for oracle/liboracle.so and the reference's own sources compiled into oracle/_ref/
(libddsbase_ref.so = src/ddsbase.cpp, libhost_ref.so = src/Camera.cpp + src/CubicSpline.cpp,
libshader_ref.so = VolumeRenderer.cs compiled as C++ through oracle/shim/glsl_compat.h).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): the checker, never the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_LIB_PATH = os.path.join(HERE, "_ref", "libddsbase_ref.so")
REF_HOST_LIB_PATH = os.path.join(HERE, "_ref", "libhost_ref.so")
REF_SHADER_LIB_PATH = os.path.join(HERE, "_ref", "libshader_ref.so")

FILTER_NEAREST = 0
FILTER_TRILINEAR = 1


def build(force: bool = False) -> None:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref)."""
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h", ".cpp"))]
    shim = os.path.join(HERE, "shim")
    for root, _, files in os.walk(shim):
        srcs += [os.path.join(root, f) for f in files]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    need_ref = os.path.exists("/root/reference/src/ddsbase.cpp") and not all(
        os.path.exists(q) for q in (REF_LIB_PATH, REF_HOST_LIB_PATH, REF_SHADER_LIB_PATH))
    if stale or need_ref:
        subprocess.run(["make", "-C", HERE, "-s"] + (["-B"] if force else []), check=True)


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("dim", C.c_int32 * 3),
        ("bytes_per_voxel", C.c_int32),
        ("voxel_size", C.c_float * 3),
        ("cam", C.c_float * 21),
        ("alpha_scale", C.c_float),
        ("min_val", C.c_int32), ("max_val", C.c_int32),
        ("is_mip", C.c_int32), ("view_top", C.c_int32), ("view_bottom", C.c_int32),
        ("filter", C.c_int32),
        ("step_scale", C.c_float),
        ("opacity_correction", C.c_int32),
        ("use_tf", C.c_int32),
        ("tf_lut", C.c_float * 256),
        ("row_begin", C.c_int32), ("row_end", C.c_int32), ("row_stride", C.c_int32),
    ]


class Counters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("rays_hit", C.c_uint64), ("samples", C.c_uint64)]


class Camera(C.Structure):
    _fields_ = [
        ("eye", C.c_float * 4), ("side", C.c_float * 4), ("up", C.c_float * 4), ("look_at", C.c_float * 4),
        ("view2world", C.c_float * 16),
        ("view_plane_dist", C.c_float), ("y_fov", C.c_float),
        ("rotation_speed", C.c_float), ("mov_speed", C.c_float),
        ("zenith", C.c_float), ("azimuth", C.c_float), ("radius", C.c_float),
        ("is_changed", C.c_int32),
    ]


SPLINE_MAX = 64


class Spline(C.Structure):
    _fields_ = [
        ("n_points", C.c_int32),
        ("iso", C.c_int32 * SPLINE_MAX),
        ("color", (C.c_float * 4) * SPLINE_MAX),
        ("coeffs", (C.c_float * 4) * SPLINE_MAX),
        ("deriv", (C.c_float * 4) * SPLINE_MAX),
        ("a", (C.c_float * 4) * SPLINE_MAX), ("b", (C.c_float * 4) * SPLINE_MAX),
        ("c", (C.c_float * 4) * SPLINE_MAX), ("d", (C.c_float * 4) * SPLINE_MAX),
    ]


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_render.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(Counters), C.c_int]
        L.orc_render.restype = C.c_int
        L.orc_popcount.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_popcount.restype = C.c_uint64
        L.orc_camera_init.argtypes = [C.POINTER(Camera), C.c_float, C.c_float, C.c_float]
        L.orc_camera_reset.argtypes = [C.POINTER(Camera)]
        L.orc_camera_set_orientation.argtypes = [C.POINTER(Camera), C.c_float, C.c_float, C.c_float]
        L.orc_camera_ubo.argtypes = [C.POINTER(Camera), C.POINTER(C.c_float * 21)]
        L.orc_spline_calc.argtypes = [C.POINTER(Spline), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.orc_spline_calc.restype = C.c_int
        L.orc_spline_eval_iso.argtypes = [C.POINTER(Spline), C.c_int, C.POINTER(C.c_float * 4)]
        L.orc_spline_eval_t.argtypes = [C.POINTER(Spline), C.c_float, C.c_int, C.POINTER(C.c_float * 4)]
        L.orc_spline_bake_alpha_lut.argtypes = [C.POINTER(Spline), C.POINTER(C.c_float * 256)]
        L.orc_pvm_decode.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_float * 3), C.POINTER(C.c_uint64)]
        L.orc_pvm_decode.restype = C.c_void_p
        L.orc_checksum.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_checksum.restype = C.c_uint32
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_synth_mix.argtypes = [C.c_void_p, C.POINTER(C.c_int32 * 3), C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.orc_synth_mix.restype = C.c_int
        _lib = L
    return _lib


def ref_lib():
    """The reference's own ddsbase.cpp (oracle/_ref); None when it was never built."""
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(REF_LIB_PATH):
            return None
        R = C.CDLL(REF_LIB_PATH)
        R.ref_readPVMvolume.argtypes = [C.c_char_p] + [C.POINTER(C.c_uint)] * 4 + [C.POINTER(C.c_float)] * 3
        R.ref_readPVMvolume.restype = C.c_void_p
        R.ref_writePVMvolume.argtypes = [C.c_char_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint,
                                         C.c_float, C.c_float, C.c_float,
                                         C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
        R.ref_writePVMvolume.restype = None
        R.ref_checksum.argtypes = [C.c_void_p, C.c_uint]
        R.ref_checksum.restype = C.c_uint
        R.ref_free.argtypes = [C.c_void_p]
        _ref = R
    return _ref


_ref_shader = None
_ref_host = None


def ref_shader_lib():
    """The reference's own VolumeRenderer.cs compiled for the CPU (oracle/_ref); None when never built."""
    global _ref_shader
    if _ref_shader is None:
        build()
        if not os.path.exists(REF_SHADER_LIB_PATH):
            return None
        S = C.CDLL(REF_SHADER_LIB_PATH)
        S.shader_ref_render.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, C.POINTER(Counters), C.c_int]
        S.shader_ref_render.restype = C.c_int
        _ref_shader = S
    return _ref_shader


def ref_host_lib():
    """The reference's own Camera.cpp + CubicSpline.cpp (oracle/_ref); None when never built."""
    global _ref_host
    if _ref_host is None:
        build()
        if not os.path.exists(REF_HOST_LIB_PATH):
            return None
        Hh = C.CDLL(REF_HOST_LIB_PATH)
        Hh.ref_camera_new.argtypes = [C.c_float, C.c_float, C.c_float]
        Hh.ref_camera_new.restype = C.c_void_p
        Hh.ref_camera_delete.argtypes = [C.c_void_p]
        Hh.ref_camera_reset.argtypes = [C.c_void_p]
        Hh.ref_camera_set_orientation.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        Hh.ref_camera_state.argtypes = [C.c_void_p, C.POINTER(C.c_float * 37), C.POINTER(C.c_int * 2)]
        Hh.ref_camera_state.restype = C.c_int
        Hh.ref_spline_new.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        Hh.ref_spline_new.restype = C.c_void_p
        Hh.ref_spline_delete.argtypes = [C.c_void_p]
        Hh.ref_spline_eval_iso.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float * 4)]
        Hh.ref_spline_eval_t.argtypes = [C.c_void_p, C.c_float, C.c_int, C.POINTER(C.c_float * 4)]
        _ref_host = Hh
    return _ref_host


# --------------------------------------------------------------------------- march

def make_params(width, height, dim, bytes_per_voxel, cam21, *, voxel_size=(1.0, 1.0, 1.0),
                alpha_scale=1.0, min_val=0, max_val=255, is_mip=0, view_top=0, view_bottom=0,
                filter=FILTER_NEAREST, step_scale=1.0, opacity_correction=0, tf_lut=None,
                row_begin=0, row_end=None, row_stride=1) -> Params:
    p = Params()
    p.width, p.height = int(width), int(height)
    p.dim[:] = [int(d) for d in dim]
    p.bytes_per_voxel = int(bytes_per_voxel)
    p.voxel_size[:] = [float(v) for v in voxel_size]
    p.cam[:] = [float(v) for v in np.asarray(cam21, dtype=np.float32)]
    p.alpha_scale = float(alpha_scale)
    p.min_val, p.max_val = int(min_val), int(max_val)
    p.is_mip, p.view_top, p.view_bottom = int(is_mip), int(view_top), int(view_bottom)
    p.filter = int(filter)
    p.step_scale = float(step_scale)
    p.opacity_correction = int(opacity_correction)
    if tf_lut is not None:
        p.use_tf = 1
        p.tf_lut[:] = [float(v) for v in np.asarray(tf_lut, dtype=np.float32)]
    p.row_begin = int(row_begin)
    p.row_end = int(height if row_end is None else row_end)
    p.row_stride = int(row_stride)
    return p


def render(p: Params, voxels: np.ndarray, *, nthreads: int = 1, touch: bool = False, out: np.ndarray = None):
    """Returns (rgba[H,W,4] float32 with row 0 = bottom, counters dict, touch bitmap or None)."""
    vox = np.ascontiguousarray(voxels)
    assert vox.dtype == (np.uint8 if p.bytes_per_voxel == 1 else np.uint16)
    assert vox.size == p.dim[0] * p.dim[1] * p.dim[2]
    if out is None:
        out = np.zeros((p.height, p.width, 4), dtype=np.float32)
    tb = None
    if touch:
        tb = np.zeros((vox.size + 7) // 8, dtype=np.uint8)
    cnt = Counters()
    rc = lib().orc_render(C.byref(p), vox.ctypes.data, out.ctypes.data,
                          tb.ctypes.data if tb is not None else None, C.byref(cnt), int(nthreads))
    if rc != 0:
        raise ValueError("orc_render: bad arguments")
    return out, {"rays": cnt.rays, "rays_hit": cnt.rays_hit, "samples": cnt.samples}, tb


def ref_render(p: Params, voxels: np.ndarray, *, nthreads: int = 1, out: np.ndarray = None):
    """The REFERENCE's own shader (VolumeRenderer.cs compiled for the CPU, oracle/_ref/libshader_ref.so) on
    the same parameter block.  Only reference semantics: raises for step_scale != 1, TF, opacity correction.
    Returns (rgba[H,W,4], counters dict with rays / samples) or None when oracle/_ref was never built."""
    S = ref_shader_lib()
    if S is None:
        return None
    vox = np.ascontiguousarray(voxels)
    assert vox.dtype == (np.uint8 if p.bytes_per_voxel == 1 else np.uint16)
    assert vox.size == p.dim[0] * p.dim[1] * p.dim[2]
    if out is None:
        out = np.zeros((p.height, p.width, 4), dtype=np.float32)
    cnt = Counters()
    rc = S.shader_ref_render(C.byref(p), vox.ctypes.data, out.ctypes.data, C.byref(cnt), int(nthreads))
    if rc == -2:
        raise ValueError("the reference shader has no step override / transfer function / opacity correction")
    if rc != 0:
        raise ValueError("shader_ref_render: bad arguments")
    return out, {"rays": cnt.rays, "samples": cnt.samples}


def synth_mix(dims, bytes_per_voxel: int, vmax: int, seed: int, with_hash: bool = True, nthreads: int = 0) -> np.ndarray:
    """The `mix` volume of SURVEY.md 8(d) generated on the host cores (flat array, x fastest)."""
    out = np.empty(int(dims[0]) * int(dims[1]) * int(dims[2]), dtype=np.uint8 if bytes_per_voxel == 1 else np.uint16)
    d = (C.c_int32 * 3)(*[int(x) for x in dims])
    if lib().orc_synth_mix(out.ctypes.data, C.byref(d), int(bytes_per_voxel), int(vmax), int(seed) & 0xFFFFFFFF,
                           1 if with_hash else 0, int(nthreads) or (os.cpu_count() or 1)) != 0:
        raise ValueError("orc_synth_mix: bad arguments")
    return out


def frame_consts(p: Params) -> np.ndarray:
    """pmin[3], pmax[3], half_len[3], denom[3], step_dvr, step_mip, fmin, fmax, frange (float32)."""
    out = np.zeros(17, dtype=np.float32)
    lib().orc_frame_consts(C.byref(p), out.ctypes.data_as(C.c_void_p))
    return out


def popcount(tb: np.ndarray, nvoxels: int) -> int:
    return int(lib().orc_popcount(tb.ctypes.data, int(nvoxels)))


# --------------------------------------------------------------------------- camera

class OracleCamera:
    def __init__(self, y_fov=30.0, rot_speed=0.7, mov_speed=0.3):
        self.c = Camera()
        lib().orc_camera_init(C.byref(self.c), y_fov, rot_speed, mov_speed)

    def reset(self):
        lib().orc_camera_reset(C.byref(self.c))

    def set_orientation(self, zoom, zenith, azimuth):
        lib().orc_camera_set_orientation(C.byref(self.c), zoom, zenith, azimuth)

    def ubo(self) -> np.ndarray:
        out = (C.c_float * 21)()
        lib().orc_camera_ubo(C.byref(self.c), C.byref(out))
        return np.array(out[:], dtype=np.float32)


class RefCamera:
    """The reference's own Camera class (src/Camera.cpp compiled into oracle/_ref/libhost_ref.so)."""

    def __init__(self, y_fov=30.0, rot_speed=0.7, mov_speed=0.3):
        self.h = ref_host_lib().ref_camera_new(y_fov, rot_speed, mov_speed)

    def __del__(self):
        if getattr(self, "h", None):
            ref_host_lib().ref_camera_delete(self.h)
            self.h = None

    def reset(self):
        ref_host_lib().ref_camera_reset(self.h)

    def set_orientation(self, zoom, zenith, azimuth):
        ref_host_lib().ref_camera_set_orientation(self.h, zoom, zenith, azimuth)

    def state(self):
        """-> (ubo[21], eye[4], side[4], up[4], look_at[4], is_changed before setUBO, after setUBO)"""
        out = (C.c_float * 37)()
        ch = (C.c_int * 2)()
        if ref_host_lib().ref_camera_state(self.h, C.byref(out), C.byref(ch)) != 0:
            raise RuntimeError("Camera::setUBO did not emit 21 floats")
        a = np.array(out[:], dtype=np.float32)
        return a[:21], a[21:25], a[25:29], a[29:33], a[33:37], bool(ch[0]), bool(ch[1])

    def ubo(self) -> np.ndarray:
        return self.state()[0]


class RefSpline:
    """The reference's own CubicSpline class (src/CubicSpline.cpp compiled into oracle/_ref/libhost_ref.so)."""

    def __init__(self, knots):
        n = len(knots)
        iso = (C.c_int * n)(*[int(k[0]) for k in knots])
        col = (C.c_float * (4 * n))()
        for i, k in enumerate(knots):
            c4 = k[1] if isinstance(k[1], (tuple, list)) else (0.0, 0.0, 0.0, k[1])
            col[i * 4:(i + 1) * 4] = [float(v) for v in c4]
        self.h = ref_host_lib().ref_spline_new(n, iso, col)

    def __del__(self):
        if getattr(self, "h", None):
            ref_host_lib().ref_spline_delete(self.h)
            self.h = None

    def eval_iso(self, iso) -> np.ndarray:
        out = (C.c_float * 4)()
        ref_host_lib().ref_spline_eval_iso(self.h, int(iso), C.byref(out))
        return np.array(out[:], dtype=np.float32)

    def eval_t(self, t, seg) -> np.ndarray:
        out = (C.c_float * 4)()
        ref_host_lib().ref_spline_eval_t(self.h, float(t), int(seg), C.byref(out))
        return np.array(out[:], dtype=np.float32)


# --------------------------------------------------------------------------- spline

DEFAULT_ALPHA_KNOTS = [(0, 0.0), (141, 0.759), (149, 0.45), (255, 1.0)]   # AlphaControlSplineWidget.cpp:56-59


class OracleSpline:
    def __init__(self, knots=DEFAULT_ALPHA_KNOTS):
        self.s = Spline()
        n = len(knots)
        iso = (C.c_int32 * n)(*[int(k[0]) for k in knots])
        col = (C.c_float * (4 * n))()
        for i, k in enumerate(knots):
            c4 = k[1] if isinstance(k[1], (tuple, list)) else (0.0, 0.0, 0.0, k[1])
            col[i * 4:(i + 1) * 4] = [float(v) for v in c4]
        if lib().orc_spline_calc(C.byref(self.s), n, iso, col) != 0:
            raise ValueError("bad knots")

    def eval_iso(self, iso) -> np.ndarray:
        out = (C.c_float * 4)()
        lib().orc_spline_eval_iso(C.byref(self.s), int(iso), C.byref(out))
        return np.array(out[:], dtype=np.float32)

    def eval_t(self, t, seg) -> np.ndarray:
        out = (C.c_float * 4)()
        lib().orc_spline_eval_t(C.byref(self.s), float(t), int(seg), C.byref(out))
        return np.array(out[:], dtype=np.float32)

    def alpha_lut(self) -> np.ndarray:
        out = (C.c_float * 256)()
        lib().orc_spline_bake_alpha_lut(C.byref(self.s), C.byref(out))
        return np.array(out[:], dtype=np.float32)

    def field(self, name) -> np.ndarray:
        n = self.s.n_points
        return np.array([list(getattr(self.s, name)[i]) for i in range(n)], dtype=np.float32)


# --------------------------------------------------------------------------- codec

def pvm_decode(file_bytes: bytes):
    """Restated decoder.  Returns dict(dims, components, scale, payload bytes) or None."""
    buf = np.frombuffer(file_bytes, dtype=np.uint8)
    w, h, d, comps = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    scale = (C.c_float * 3)()
    nbytes = C.c_uint64()
    ptr = lib().orc_pvm_decode(buf.ctypes.data, buf.size, C.byref(w), C.byref(h), C.byref(d),
                               C.byref(comps), C.byref(scale), C.byref(nbytes))
    if not ptr:
        return None
    payload = C.string_at(ptr, nbytes.value)
    lib().orc_free(ptr)
    return {"dims": (w.value, h.value, d.value), "components": comps.value,
            "scale": tuple(scale[:]), "payload": payload}


def checksum(data: bytes) -> int:
    buf = np.frombuffer(data, dtype=np.uint8)
    return int(lib().orc_checksum(buf.ctypes.data, buf.size))


def ref_read_pvm(path: str):
    R = ref_lib()
    if R is None:
        return None
    w, h, d, comps = C.c_uint(), C.c_uint(), C.c_uint(), C.c_uint()
    sx, sy, sz = C.c_float(), C.c_float(), C.c_float()
    ptr = R.ref_readPVMvolume(path.encode(), C.byref(w), C.byref(h), C.byref(d), C.byref(comps),
                              C.byref(sx), C.byref(sy), C.byref(sz))
    if not ptr:
        return None
    n = w.value * h.value * d.value * comps.value
    payload = C.string_at(ptr, n)
    R.ref_free(ptr)
    return {"dims": (w.value, h.value, d.value), "components": comps.value,
            "scale": (sx.value, sy.value, sz.value), "payload": payload}


def ref_write_pvm(path: str, volume: np.ndarray, dims, components=1, scale=(1.0, 1.0, 1.0),
                  description=None, courtesy=None, parameter=None, comment=None) -> None:
    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref not built")
    vol = np.ascontiguousarray(volume).view(np.uint8).copy()
    assert vol.size == dims[0] * dims[1] * dims[2] * components
    enc = lambda s: None if s is None else s.encode()
    R.ref_writePVMvolume(path.encode(), vol.ctypes.data, dims[0], dims[1], dims[2], components,
                         scale[0], scale[1], scale[2],
                         enc(description), enc(courtesy), enc(parameter), enc(comment))


def ref_checksum(data: bytes) -> int:
    R = ref_lib()
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    return int(R.ref_checksum(buf.ctypes.data, buf.size))
