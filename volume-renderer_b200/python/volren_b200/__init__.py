"""volren_b200 -- thin ctypes harness over the C-ABI of libvolren_b200.so.

This package is test / bench plumbing: every call goes straight through include/volren_b200.h.
There is no Python or CPU implementation of the ray march here; if the CUDA library is
missing or no GPU is usable, calls raise VolrenError.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # volume-renderer_b200/
REPO_ROOT = os.path.dirname(PKG_ROOT)
LIB_DIR = os.path.join(PKG_ROOT, "lib")
LIB_PATH = os.environ.get("VOLREN_B200_LIB") or os.path.join(LIB_DIR, "libvolren_b200.so")   # env override: lab builds
HOST_LIB_PATH = os.path.join(LIB_DIR, "libvolren_host.so")

FILTER_NEAREST, FILTER_TRILINEAR = 0, 1
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TEXPAIR_PIPE, KERNEL_NEAREST_TEX = 0, 1, 7, 10
KERNEL_NAMES = {KERNEL_DIRECT: "march_direct_kernel", KERNEL_TEXPAIR_PIPE: "march_texpair_kernel", KERNEL_NEAREST_TEX: "march_nearest_kernel"}
SKIP_AUTO, SKIP_ON, SKIP_OFF = 0, 1, 2

VR_OK = 0

# every symbol include/volren_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "vr_version", "vr_last_error", "vr_params_default", "vr_device_count",
    "vr_create", "vr_destroy", "vr_resize", "vr_image_size",
    "vr_upload_volume", "vr_upload_volume_device", "vr_set_voxel_size", "vr_volume_stats_get",
    "vr_cell_table_get", "vr_memory_info_get", "vr_render_peer", "vr_peer_frame_reset", "vr_peer_kernel_ms", "vr_peer_frame_wait_arrivals",
    "vr_set_camera", "vr_set_params", "vr_get_params", "vr_set_partition", "vr_owned_rows",
    "vr_render", "vr_read_frame", "vr_render_device", "vr_render_owned_to_host", "vr_render_submit", "vr_render_wait", "vr_assemble_tiles", "vr_read_rgb8", "vr_count_frame",
    "vr_frame_device_ptr", "vr_frame_export_ipc", "vr_frame_open_ipc", "vr_frame_close_ipc",
    "vr_peer_frame_arrive", "vr_peer_frame_release", "vr_peer_frame_status",
    "vr_upload_synthetic", "vr_synthetic_to_host",
]


class VolrenError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"volren_b200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("alpha_scale", C.c_float),
        ("min_val", C.c_int32), ("max_val", C.c_int32),
        ("is_mip", C.c_int32), ("view_top", C.c_int32), ("view_bottom", C.c_int32),
        ("filter", C.c_int32),
        ("step_scale", C.c_float),
        ("opacity_correction", C.c_int32),
        ("use_tf", C.c_int32),
        ("tf_lut", C.c_float * 256),
        ("kernel", C.c_int32),
        ("empty_skip", C.c_int32),
    ]


class RenderStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_float), ("total_ms", C.c_float),
                ("kernel_launches", C.c_uint32), ("kernel_used", C.c_uint32), ("skip_used", C.c_uint32)]


class MemoryInfo(C.Structure):
    _fields_ = [("linear_bytes", C.c_uint64), ("array_bytes", C.c_uint64), ("zpair_array_bytes", C.c_uint64),
                ("cell_table_bytes", C.c_uint64), ("frame_bytes", C.c_uint64)]


class VolumeStats(C.Structure):
    _fields_ = [("min_value", C.c_int32), ("max_value", C.c_int32), ("histogram", C.c_float * 256)]


def build(verbose: bool = False) -> None:
    """Compile the CUDA library (and the C++ host mirror) in-tree for sm_100a."""
    res = subprocess.run(["bash", os.path.join(PKG_ROOT, "build.sh")], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("volren_b200 build failed")


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VolrenError(-2, f"{LIB_PATH} is missing: run volume-renderer_b200/build.sh "
                                  "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.vr_version.restype = C.c_char_p
        L.vr_last_error.restype = C.c_char_p
        L.vr_params_default.argtypes = [C.POINTER(Params)]
        L.vr_params_default.restype = None
        L.vr_device_count.argtypes = [C.POINTER(C.c_int)]
        L.vr_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.vr_destroy.argtypes = [C.c_void_p]
        L.vr_destroy.restype = None
        L.vr_resize.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.vr_image_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.vr_upload_volume.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64 * 3), C.c_int, C.POINTER(C.c_float * 3)]
        L.vr_upload_volume_device.argtypes = L.vr_upload_volume.argtypes
        L.vr_set_voxel_size.argtypes = [C.c_void_p, C.POINTER(C.c_float * 3)]
        L.vr_volume_stats_get.argtypes = [C.c_void_p, C.POINTER(VolumeStats)]
        L.vr_set_camera.argtypes = [C.c_void_p, C.POINTER(C.c_float * 21)]
        L.vr_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.vr_get_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.vr_set_partition.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.vr_owned_rows.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.vr_render.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RenderStats)]
        L.vr_read_frame.argtypes = [C.c_void_p, C.c_void_p]
        L.vr_render_owned_to_host.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RenderStats)]
        L.vr_render_submit.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        L.vr_render_wait.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(RenderStats)]
        L.vr_render_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(RenderStats)]
        L.vr_assemble_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vr_read_rgb8.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.vr_frame_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.vr_frame_export_ipc.argtypes = [C.c_void_p, C.c_void_p]
        L.vr_frame_open_ipc.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.vr_frame_close_ipc.argtypes = [C.c_void_p, C.c_void_p]
        L.vr_peer_frame_arrive.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        L.vr_peer_frame_release.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.vr_peer_frame_status.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.vr_render_peer.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.POINTER(RenderStats)]
        L.vr_peer_kernel_ms.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
        L.vr_peer_frame_wait_arrivals.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.vr_peer_frame_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.vr_cell_table_get.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int * 3), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.vr_memory_info_get.argtypes = [C.c_void_p, C.POINTER(MemoryInfo)]
        L.vr_count_frame.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.vr_upload_synthetic.argtypes = [C.c_void_p, C.POINTER(C.c_uint64 * 3), C.c_int, C.POINTER(C.c_float * 3),
                                          C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.vr_synthetic_to_host.argtypes = [C.c_int, C.POINTER(C.c_uint64 * 3), C.c_int, C.c_uint32, C.c_uint32,
                                           C.c_int, C.c_void_p]
        for name in ABI_SYMBOLS:
            fn = getattr(L, name)
            if fn.restype is C.c_int:
                pass
        _lib = L
    return _lib


def _check(rc):
    if rc != VR_OK:
        raise VolrenError(rc, lib().vr_last_error().decode(errors="replace"))


def default_params(**kw) -> Params:
    p = Params()
    lib().vr_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "tf_lut":
            if v is not None:
                p.use_tf = 1
                p.tf_lut[:] = [float(x) for x in np.asarray(v, dtype=np.float32)]
        else:
            setattr(p, k, v)
    return p


def device_count() -> int:
    n = C.c_int()
    _check(lib().vr_device_count(C.byref(n)))
    return n.value


class Context:
    """One render target + volume on one GPU (the RendererCore's GL state, on CUDA)."""

    def __init__(self, width: int, height: int, device: int = 0):
        self._h = C.c_void_p()
        _check(lib().vr_create(device, width, height, C.byref(self._h)))
        self.width, self.height, self.device = width, height, device

    def close(self):
        if self._h:
            lib().vr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- volume
    def upload_volume(self, voxels: np.ndarray, dims, voxel_size=(1.0, 1.0, 1.0)):
        vox = np.ascontiguousarray(voxels)
        if vox.dtype not in (np.uint8, np.uint16):
            raise TypeError("voxels must be uint8 or uint16")
        assert vox.size == dims[0] * dims[1] * dims[2]
        d = (C.c_uint64 * 3)(*[int(x) for x in dims])
        vs = (C.c_float * 3)(*[float(x) for x in voxel_size])
        _check(lib().vr_upload_volume(self._h, vox.ctypes.data, C.byref(d), vox.dtype.itemsize, C.byref(vs)))

    def upload_volume_device(self, dptr: int, dims, bytes_per_voxel: int, voxel_size=(1.0, 1.0, 1.0)):
        d = (C.c_uint64 * 3)(*[int(x) for x in dims])
        vs = (C.c_float * 3)(*[float(x) for x in voxel_size])
        _check(lib().vr_upload_volume_device(self._h, C.c_void_p(dptr), C.byref(d), bytes_per_voxel, C.byref(vs)))

    def upload_synthetic(self, dims, bytes_per_voxel, vmax, seed, with_hash=True,
                         voxel_size=(1.0, 1.0, 1.0), copy_out_dptr: int = 0):
        d = (C.c_uint64 * 3)(*[int(x) for x in dims])
        vs = (C.c_float * 3)(*[float(x) for x in voxel_size])
        _check(lib().vr_upload_synthetic(self._h, C.byref(d), bytes_per_voxel, C.byref(vs), vmax, seed,
                                         1 if with_hash else 0, C.c_void_p(copy_out_dptr) if copy_out_dptr else None))

    def set_voxel_size(self, voxel_size):
        vs = (C.c_float * 3)(*[float(x) for x in voxel_size])
        _check(lib().vr_set_voxel_size(self._h, C.byref(vs)))

    def volume_stats(self):
        st = VolumeStats()
        _check(lib().vr_volume_stats_get(self._h, C.byref(st)))
        return st.min_value, st.max_value, np.array(st.histogram[:], dtype=np.float32)

    # -- per-frame state
    def set_camera(self, cam21):
        cam = (C.c_float * 21)(*[float(x) for x in np.asarray(cam21, dtype=np.float32)])
        _check(lib().vr_set_camera(self._h, C.byref(cam)))

    def set_params(self, p: Params):
        _check(lib().vr_set_params(self._h, C.byref(p)))

    def get_params(self) -> Params:
        p = Params()
        _check(lib().vr_get_params(self._h, C.byref(p)))
        return p

    def set_partition(self, rank: int, world: int, tile_rows: int = 8):
        _check(lib().vr_set_partition(self._h, rank, world, tile_rows))

    def owned_rows(self) -> int:
        n = C.c_int()
        _check(lib().vr_owned_rows(self._h, C.byref(n)))
        return n.value

    # -- render
    def render(self, out: np.ndarray = None):
        """Full frame to host memory.  Returns (rgba[H,W,4], RenderStats)."""
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == self.width * self.height * 4
        st = RenderStats()
        _check(lib().vr_render(self._h, out.ctypes.data, C.byref(st)))
        return out, st

    def read_frame(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        _check(lib().vr_read_frame(self._h, out.ctypes.data))
        return out

    def read_frame_into(self, host_ptr: int):
        _check(lib().vr_read_frame(self._h, C.c_void_p(host_ptr)))

    def render_to_host_ptr(self, host_ptr: int):
        st = RenderStats()
        _check(lib().vr_render(self._h, C.c_void_p(host_ptr), C.byref(st)))
        return st

    def render_owned_to_host_ptr(self, host_full_frame_ptr: int):
        """This rank's row tiles into their rows of a full host frame (shared by all ranks)."""
        st = RenderStats()
        _check(lib().vr_render_owned_to_host(self._h, C.c_void_p(host_full_frame_ptr), C.byref(st)))
        return st

    def render_submit(self, host_frame_ptr: int) -> int:
        """Pipelined form of render_owned_to_host_ptr (the whole frame when world == 1): enqueue and return a ticket;
        at most two frames in flight, each into its own page-locked host buffer."""
        t = C.c_uint32(0)
        _check(lib().vr_render_submit(self._h, C.c_void_p(host_frame_ptr), C.byref(t)))
        return t.value

    def render_wait(self, ticket: int):
        """Block until the frame of `ticket` is complete in host memory."""
        st = RenderStats()
        _check(lib().vr_render_wait(self._h, C.c_uint32(ticket), C.byref(st)))
        return st

    def render_device(self, dptr: int, compact: bool = False, stream: int = 0):
        st = RenderStats()
        _check(lib().vr_render_device(self._h, C.c_void_p(dptr), 1 if compact else 0,
                                      C.c_void_p(stream) if stream else None, C.byref(st)))
        return st

    def assemble_tiles(self, gathered_dptr: int, frame_dptr: int, world: int, tile_rows: int, stream: int = 0):
        _check(lib().vr_assemble_tiles(self._h, C.c_void_p(gathered_dptr), C.c_void_p(frame_dptr), world, tile_rows,
                                       C.c_void_p(stream) if stream else None))

    # -- fused multi-GPU hand-off (peer stores into rank 0's frame)
    def frame_device_ptr(self) -> int:
        p = C.c_void_p()
        _check(lib().vr_frame_device_ptr(self._h, C.byref(p)))
        return p.value

    def frame_export_ipc(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        _check(lib().vr_frame_export_ipc(self._h, buf))
        return bytes(buf)

    def frame_open_ipc(self, handle: bytes) -> int:
        buf = (C.c_ubyte * 64)(*handle)
        p = C.c_void_p()
        _check(lib().vr_frame_open_ipc(self._h, buf, C.byref(p)))
        return p.value

    def frame_close_ipc(self, ptr: int):
        _check(lib().vr_frame_close_ipc(self._h, C.c_void_p(ptr)))

    # -- frame barrier in the same peer memory (replaces the per-frame NCCL barrier)
    def peer_frame_arrive(self, target_ptr: int, frame_no: int, world: int, is_owner: bool, stream: int = 0):
        _check(lib().vr_peer_frame_arrive(self._h, C.c_void_p(target_ptr), frame_no & 0xffffffff, world, 1 if is_owner else 0,
                                          C.c_void_p(stream)))

    def peer_frame_release(self, target_ptr: int, frame_no: int, is_owner: bool, stream: int = 0):
        _check(lib().vr_peer_frame_release(self._h, C.c_void_p(target_ptr), frame_no & 0xffffffff, 1 if is_owner else 0,
                                           C.c_void_p(stream)))

    def peer_frame_status(self, target_ptr: int):
        a, r, t = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().vr_peer_frame_status(self._h, C.c_void_p(target_ptr), C.byref(a), C.byref(r), C.byref(t)))
        return {"arrivals": a.value, "released": r.value, "timed_out": t.value}

    def render_peer(self, target_ptr: int, frame_no: int, world: int, is_owner: bool, stream: int = 0, wait: bool = True):
        """Render this rank's tiles into the (own or peer) frame; the kernel's last CTA signals the arrival.
        wait=False: asynchronous (returns None; peer_kernel_ms(frame_no) reads the march time later)."""
        if not wait:
            _check(lib().vr_render_peer(self._h, C.c_void_p(target_ptr), frame_no & 0xffffffff, world, 1 if is_owner else 0,
                                        C.c_void_p(stream) if stream else None, None))
            return None
        st = RenderStats()
        _check(lib().vr_render_peer(self._h, C.c_void_p(target_ptr), frame_no & 0xffffffff, world, 1 if is_owner else 0,
                                    C.c_void_p(stream) if stream else None, C.byref(st)))
        return st

    def peer_kernel_ms(self, frame_no: int) -> float:
        ms = C.c_float(0.0)
        _check(lib().vr_peer_kernel_ms(self._h, frame_no & 0xffffffff, C.byref(ms)))
        return ms.value

    def peer_frame_wait_arrivals(self, target_ptr: int, frame_no: int, world: int, stream: int = 0):
        """Owner side, on a stream of its choice: wait until every rank has arrived for `frame_no`."""
        _check(lib().vr_peer_frame_wait_arrivals(self._h, C.c_void_p(target_ptr), frame_no & 0xffffffff, world,
                                                 C.c_void_p(stream) if stream else None))

    def peer_frame_reset(self, target_ptr: int = 0):
        _check(lib().vr_peer_frame_reset(self._h, C.c_void_p(target_ptr) if target_ptr else None))

    def cell_table(self):
        """-> dict(shift, cells (cx,cy,cz), mins[cz,cy,cx], maxs[cz,cy,cx], empty_cells under the current min_val)"""
        sh, cells, ne = C.c_int(), (C.c_int * 3)(), C.c_uint64()
        _check(lib().vr_cell_table_get(self._h, C.byref(sh), C.byref(cells), None, None, None))
        n = cells[0] * cells[1] * cells[2]
        mins, maxs = np.empty(n, np.uint16), np.empty(n, np.uint16)
        _check(lib().vr_cell_table_get(self._h, C.byref(sh), C.byref(cells), mins.ctypes.data, maxs.ctypes.data, C.byref(ne)))
        shape = (cells[2], cells[1], cells[0])
        return {"shift": sh.value, "cells": tuple(cells[:]), "mins": mins.reshape(shape), "maxs": maxs.reshape(shape),
                "empty_cells": ne.value}

    def memory_info(self):
        m = MemoryInfo()
        _check(lib().vr_memory_info_get(self._h, C.byref(m)))
        return {k: getattr(m, k) for k, _ in MemoryInfo._fields_}

    def read_rgb8(self, flip_vertical: bool = True) -> np.ndarray:
        out = np.empty((self.height, self.width, 3), dtype=np.uint8)
        _check(lib().vr_read_rgb8(self._h, out.ctypes.data, 1 if flip_vertical else 0))
        return out

    def count_frame(self):
        dv, ns, rh = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().vr_count_frame(self._h, C.byref(dv), C.byref(ns), C.byref(rh)))
        return {"distinct_voxels": dv.value, "samples": ns.value, "rays_hit": rh.value}


def synthetic_to_host(dims, bytes_per_voxel, vmax, seed, with_hash=True, device=0) -> np.ndarray:
    out = np.empty(dims[0] * dims[1] * dims[2], dtype=np.uint8 if bytes_per_voxel == 1 else np.uint16)
    d = (C.c_uint64 * 3)(*[int(x) for x in dims])
    _check(lib().vr_synthetic_to_host(device, C.byref(d), bytes_per_voxel, vmax, seed, 1 if with_hash else 0,
                                      out.ctypes.data))
    return out
