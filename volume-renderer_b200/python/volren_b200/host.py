"""ctypes view of libvolren_host.so: the C++ host mirror of the reference's Camera /
CubicSpline / loaders / RendererCore (volume-renderer_b200/host/), driven the way the
reference's RendererGUI drives RendererCore."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import HOST_LIB_PATH, LIB_PATH, Params, VolrenError

_h = None


class CoreState(C.Structure):
    _fields_ = [("tex3D_dim", C.c_int * 3), ("voxel_size", C.c_float * 3),
                ("datasize_bytes", C.c_int), ("min_val", C.c_int), ("max_val", C.c_int),
                ("min_dataset_val", C.c_int), ("max_dataset_val", C.c_int),
                ("workgroups_x", C.c_int), ("workgroups_y", C.c_int),
                ("alpha_scale", C.c_float), ("kerneltime_sum", C.c_float),
                ("last_kernel_ms", C.c_float), ("last_kernel_used", C.c_uint)]


def hlib():
    global _h
    if _h is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise VolrenError(-2, f"{HOST_LIB_PATH} is missing: run volume-renderer_b200/build.sh")
        C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        H = C.CDLL(HOST_LIB_PATH)
        vp, f, i = C.c_void_p, C.c_float, C.c_int
        sig = {
            "vrh_camera_new": ([f, f, f], vp), "vrh_camera_free": ([vp], None), "vrh_camera_reset": ([vp], None),
            "vrh_camera_set_orientation": ([vp, f, f, f], None), "vrh_camera_set_spherical": ([vp, f, f, f], None),
            "vrh_camera_is_changed": ([vp], i), "vrh_camera_ubo": ([vp, C.POINTER(f * 21)], None),
            "vrh_spline_new": ([i, C.POINTER(C.c_int), C.POINTER(f)], vp), "vrh_spline_free": ([vp], None),
            "vrh_spline_eval_iso": ([vp, i, C.POINTER(f * 4)], None), "vrh_spline_eval_t": ([vp, f, i, C.POINTER(f * 4)], None),
            "vrh_spline_bake_alpha_lut": ([vp, C.POINTER(f * 256)], None),
            "vrh_pvm_decode": ([vp, C.c_uint64], vp), "vrh_pvm_read": ([C.c_char_p], vp), "vrh_pvm_ok": ([vp], i),
            "vrh_pvm_error": ([vp], C.c_char_p), "vrh_pvm_header": ([vp, C.POINTER(C.c_uint32 * 5), C.POINTER(f * 3)], None),
            "vrh_pvm_payload_bytes": ([vp], C.c_uint64), "vrh_pvm_payload": ([vp], vp), "vrh_pvm_string": ([vp, i], C.c_char_p),
            "vrh_pvm_free": ([vp], None), "vrh_dds_checksum": ([vp, C.c_uint64], C.c_uint32),
            "vrh_raw_map": ([C.c_char_p], vp), "vrh_raw_size": ([vp], C.c_uint64), "vrh_raw_byte": ([vp, C.c_uint64], i),
            "vrh_raw_free": ([vp], None),
            "vrh_checked_volume_bytes": ([C.POINTER(C.c_uint64 * 3), C.c_uint64, C.POINTER(C.c_uint64)], i),
            "vrh_rawinf_write": ([C.c_char_p, C.POINTER(C.c_int * 3), C.POINTER(f * 3)], i),
            "vrh_rawinf_read": ([C.c_char_p, C.POINTER(C.c_int * 3), C.POINTER(f * 3), C.c_char_p, C.c_char_p, i], i),
            "vrh_write_image": ([C.c_char_p, C.c_char_p, i, i, vp], i),
            "vrh_core_new": ([i, i, i], vp), "vrh_core_free": ([vp], None), "vrh_core_setup": ([vp, C.c_char_p, i], i),
            "vrh_core_load_shader": ([vp, C.c_char_p], i), "vrh_core_set_datasize": ([vp, i], None),
            "vrh_core_set_raw_info": ([vp, C.POINTER(C.c_int * 3), C.POINTER(f * 3)], None),
            "vrh_core_check_raw_inf": ([vp, C.c_char_p], i), "vrh_core_read_volume": ([vp, C.c_char_p], None),
            "vrh_core_render": ([vp], None), "vrh_core_read_frame": ([vp, vp], i),
            "vrh_core_save_image": ([vp, C.c_char_p, C.c_char_p], i),
            "vrh_core_camera_orient": ([vp, f, f, f], None), "vrh_core_camera_reset": ([vp], None),
            "vrh_core_camera_ubo": ([vp, C.POINTER(f * 21)], None),
            "vrh_core_gui_alpha": ([vp, f], None), "vrh_core_gui_mip": ([vp, i], None),
            "vrh_core_gui_min": ([vp, i], None), "vrh_core_gui_max": ([vp, i], None), "vrh_core_gui_view": ([vp, i, i], None),
            "vrh_core_ext_filter": ([vp, i], None), "vrh_core_ext_step": ([vp, f, i], None),
            "vrh_core_ext_tf": ([vp, vp], None), "vrh_core_ext_kernel": ([vp, i], None),
            "vrh_core_get_params": ([vp, C.POINTER(Params)], None),
            "vrh_core_state_get": ([vp, C.POINTER(CoreState)], None), "vrh_core_reset_kerneltime": ([vp], None),
            "vrh_core_strings": ([vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, i], None),
            "vrh_core_clear_popup": ([vp], None), "vrh_core_histogram": ([vp, C.POINTER(f * 256)], None),
        }
        for name, (args, res) in sig.items():
            fn = getattr(H, name)
            fn.argtypes, fn.restype = args, res
        _h = H
    return _h


class Camera:
    """host/Camera.cpp (reference surface: include/Camera.h:9-33)."""

    def __init__(self, y_fov=30.0, rot_speed=0.7, mov_speed=0.3):
        self._p = hlib().vrh_camera_new(y_fov, rot_speed, mov_speed)

    def __del__(self):
        if getattr(self, "_p", None):
            hlib().vrh_camera_free(self._p)
            self._p = None

    def resetCamera(self):
        hlib().vrh_camera_reset(self._p)

    def setOrientation(self, zoom, zenith, azimuth):
        hlib().vrh_camera_set_orientation(self._p, zoom, zenith, azimuth)

    def setSpherical(self, radius, zenith, azimuth):
        hlib().vrh_camera_set_spherical(self._p, radius, zenith, azimuth)

    @property
    def is_changed(self):
        return bool(hlib().vrh_camera_is_changed(self._p))

    def ubo(self) -> np.ndarray:
        out = (C.c_float * 21)()
        hlib().vrh_camera_ubo(self._p, C.byref(out))
        return np.array(out[:], dtype=np.float32)


class CubicSpline:
    """host/CubicSpline.cpp (reference surface: include/CubicSpline.h:7-31)."""

    def __init__(self, knots):
        n = len(knots)
        iso = (C.c_int * n)(*[int(k[0]) for k in knots])
        col = (C.c_float * (4 * n))()
        for j, k in enumerate(knots):
            c4 = k[1] if isinstance(k[1], (tuple, list)) else (0.0, 0.0, 0.0, k[1])
            col[j * 4:(j + 1) * 4] = [float(v) for v in c4]
        self._p = hlib().vrh_spline_new(n, iso, col)

    def __del__(self):
        if getattr(self, "_p", None):
            hlib().vrh_spline_free(self._p)
            self._p = None

    def getPointOnSpline(self, iso: int) -> np.ndarray:
        out = (C.c_float * 4)()
        hlib().vrh_spline_eval_iso(self._p, int(iso), C.byref(out))
        return np.array(out[:], dtype=np.float32)

    def getPointOnSplineT(self, t: float, seg: int) -> np.ndarray:
        out = (C.c_float * 4)()
        hlib().vrh_spline_eval_t(self._p, float(t), int(seg), C.byref(out))
        return np.array(out[:], dtype=np.float32)

    def bakeAlphaLUT(self) -> np.ndarray:
        out = (C.c_float * 256)()
        hlib().vrh_spline_bake_alpha_lut(self._p, C.byref(out))
        return np.array(out[:], dtype=np.float32)


def pvm_decode(data: bytes = None, path: str = None):
    H = hlib()
    if path is not None:
        p = H.vrh_pvm_read(path.encode())
    else:
        buf = np.frombuffer(data, dtype=np.uint8)
        p = H.vrh_pvm_decode(buf.ctypes.data, buf.size)
    try:
        if not H.vrh_pvm_ok(p):
            return {"ok": False, "error": H.vrh_pvm_error(p).decode()}
        hdr = (C.c_uint32 * 5)()
        sc = (C.c_float * 3)()
        H.vrh_pvm_header(p, C.byref(hdr), C.byref(sc))
        n = H.vrh_pvm_payload_bytes(p)
        payload = C.string_at(H.vrh_pvm_payload(p), n)
        return {"ok": True, "dims": (hdr[0], hdr[1], hdr[2]), "components": hdr[3], "version": hdr[4],
                "scale": tuple(sc[:]), "payload": payload,
                "strings": [H.vrh_pvm_string(p, k).decode(errors="replace") for k in range(4)]}
    finally:
        H.vrh_pvm_free(p)


def dds_checksum(data: bytes) -> int:
    buf = np.frombuffer(data, dtype=np.uint8)
    return int(hlib().vrh_dds_checksum(buf.ctypes.data, buf.size))


class MappedRaw:
    """Read-only memory map of a .raw payload (host/VolumeIO.h MappedFile): 64-bit sizes, no host copy."""

    def __init__(self, fn: str):
        self._p = hlib().vrh_raw_map(fn.encode())
        if not self._p:
            raise OSError("cannot map " + fn)

    def __del__(self):
        if getattr(self, "_p", None):
            hlib().vrh_raw_free(self._p)
            self._p = None

    @property
    def size(self) -> int:
        return int(hlib().vrh_raw_size(self._p))

    def byte(self, offset: int) -> int:
        return int(hlib().vrh_raw_byte(self._p, int(offset)))


def checked_volume_bytes(dims, bytes_per_voxel: int):
    """-> byte count, or None when a dimension is outside [1,16384] (overflow-safe)."""
    d = (C.c_uint64 * 3)(*[int(x) & 0xFFFFFFFFFFFFFFFF for x in dims])
    out = C.c_uint64()
    return int(out.value) if hlib().vrh_checked_volume_bytes(C.byref(d), int(bytes_per_voxel), C.byref(out)) else None


def rawinf_write(raw_fn: str, dims, spacing) -> bool:
    d = (C.c_int * 3)(*[int(x) for x in dims])
    s = (C.c_float * 3)(*[float(x) for x in spacing])
    return bool(hlib().vrh_rawinf_write(raw_fn.encode(), C.byref(d), C.byref(s)))


def rawinf_read(raw_fn: str):
    d = (C.c_int * 3)()
    s = (C.c_float * 3)()
    t = C.create_string_buffer(512)
    m = C.create_string_buffer(512)
    rc = hlib().vrh_rawinf_read(raw_fn.encode(), C.byref(d), C.byref(s), t, m, 512)
    return rc, tuple(d[:]), tuple(s[:]), t.value.decode(), m.value.decode()


def write_image(fn: str, ext: str, rgb: np.ndarray) -> bool:
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w = rgb.shape[:2]
    return bool(hlib().vrh_write_image(fn.encode(), ext.encode(), w, h, rgb.ctypes.data))


class RendererCore:
    """host/RendererCore.cpp driven like RendererGUI drives the reference's RendererCore."""

    def __init__(self, width, height, device=0):
        self.width, self.height = width, height
        self._p = hlib().vrh_core_new(device, width, height)

    def __del__(self):
        if getattr(self, "_p", None):
            hlib().vrh_core_free(self._p)
            self._p = None

    def setup(self):
        err = C.create_string_buffer(512)
        if not hlib().vrh_core_setup(self._p, err, 512):
            raise RuntimeError(err.value.decode())

    def loadShader(self, fn="VolumeRenderer.cs"):
        return bool(hlib().vrh_core_load_shader(self._p, fn.encode()))

    def set_datasize_bytes(self, n):
        hlib().vrh_core_set_datasize(self._p, n)

    def set_raw_info(self, dims, spacing):
        d = (C.c_int * 3)(*[int(x) for x in dims])
        s = (C.c_float * 3)(*[float(x) for x in spacing])
        hlib().vrh_core_set_raw_info(self._p, C.byref(d), C.byref(s))

    def checkRawInfFile(self, fn):
        return bool(hlib().vrh_core_check_raw_inf(self._p, fn.encode()))

    def readVolumeData(self, fn):
        hlib().vrh_core_read_volume(self._p, fn.encode())

    def render(self):
        hlib().vrh_core_render(self._p)

    def readFrame(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        if not hlib().vrh_core_read_frame(self._p, out.ctypes.data):
            raise RuntimeError("readFrame failed: " + self.strings()["msg"])
        return out

    def saveImage(self, fn, ext):
        return bool(hlib().vrh_core_save_image(self._p, fn.encode(), ext.encode()))

    def camera_setOrientation(self, zoom, zenith, azimuth):
        hlib().vrh_core_camera_orient(self._p, zoom, zenith, azimuth)

    def camera_reset(self):
        hlib().vrh_core_camera_reset(self._p)

    def camera_ubo(self):
        out = (C.c_float * 21)()
        hlib().vrh_core_camera_ubo(self._p, C.byref(out))
        return np.array(out[:], dtype=np.float32)

    def gui_alpha(self, a): hlib().vrh_core_gui_alpha(self._p, a)
    def gui_mip(self, on): hlib().vrh_core_gui_mip(self._p, int(on))
    def gui_min(self, v): hlib().vrh_core_gui_min(self._p, int(v))
    def gui_max(self, v): hlib().vrh_core_gui_max(self._p, int(v))
    def gui_view(self, top, bottom): hlib().vrh_core_gui_view(self._p, int(top), int(bottom))
    def ext_filter(self, f): hlib().vrh_core_ext_filter(self._p, int(f))
    def ext_step(self, s, oc=False): hlib().vrh_core_ext_step(self._p, float(s), int(oc))
    def ext_kernel(self, k): hlib().vrh_core_ext_kernel(self._p, int(k))

    def ext_tf(self, lut):
        if lut is None:
            hlib().vrh_core_ext_tf(self._p, None)
        else:
            a = np.ascontiguousarray(lut, dtype=np.float32)
            hlib().vrh_core_ext_tf(self._p, a.ctypes.data)

    def params(self) -> Params:
        p = Params()
        hlib().vrh_core_get_params(self._p, C.byref(p))
        return p

    def state(self) -> CoreState:
        s = CoreState()
        hlib().vrh_core_state_get(self._p, C.byref(s))
        return s

    def reset_kerneltime(self):
        hlib().vrh_core_reset_kerneltime(self._p)

    def strings(self):
        bufs = [C.create_string_buffer(1024) for _ in range(4)]
        hlib().vrh_core_strings(self._p, *bufs, 1024)
        return dict(zip(("title", "msg", "loaded_dataset", "loaded_shader"), [b.value.decode() for b in bufs]))

    def clear_popup(self):
        hlib().vrh_core_clear_popup(self._p)

    def histogram(self):
        out = (C.c_float * 256)()
        hlib().vrh_core_histogram(self._p, C.byref(out))
        return np.array(out[:], dtype=np.float32)


def frame_consts(width, height, dims, voxel_size, cam21, params) -> np.ndarray:
    """The product's host-side per-frame constants (csrc/frame.h): pmin[3], pmax[3], half_len[3], denom[3],
    step, fmin, fmax, frange, inv_denom[3], inv_frange, tc_div_mode, win_div_mode."""
    H = hlib()
    out = np.zeros(22, dtype=np.float32)
    d = (C.c_int32 * 3)(*[int(x) for x in dims])
    vs = (C.c_float * 3)(*[float(x) for x in voxel_size])
    cam = (C.c_float * 21)(*[float(x) for x in cam21])
    H.vrh_frame_consts(int(width), int(height), d, vs, cam, C.byref(params), out.ctypes.data_as(C.c_void_p))
    return out
