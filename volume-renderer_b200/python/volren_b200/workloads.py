"""Deterministic synthetic workloads of SURVEY.md 8(d): the `mix` volume, the cameras K0/K1/K2
and the five BASELINE.json configurations.  Pure numpy + the product's own host Camera; no
oracle imports here (bench.py's product arm uses this module)."""
from __future__ import annotations

import math

import numpy as np

SEEDS = {"C1": 0x5EED0001, "C2": 0x5EED0002, "C3": 0x5EED0003, "C4": 0x5EED0004, "C5": 0x5EED0005}


def mix_volume(dims, vmax: int, seed: int, with_hash: bool = True, dtype=None) -> np.ndarray:
    """voxel(i,j,k) of SURVEY.md 8(d), float64 arithmetic, x fastest.  Returns a flat array."""
    nx, ny, nz = dims
    if dtype is None:
        dtype = np.uint8 if vmax <= 255 else np.uint16
    out = np.empty((nz, ny, nx), dtype=dtype)
    i = np.arange(nx, dtype=np.uint32)[None, :]
    j = np.arange(ny, dtype=np.uint32)[:, None]
    px = (np.arange(nx, dtype=np.float64)[None, :] + 0.5) / nx
    py = (np.arange(ny, dtype=np.float64)[:, None] + 0.5) / ny
    two_pi = 6.283185307179586476925286766559
    sx = np.sin(two_pi * (3.0 * px + 0.1))
    sy = np.sin(two_pi * (2.0 * py + 0.2))
    hx = i * np.uint32(73856093)
    hy = j * np.uint32(19349663)
    for k in range(nz):
        pz = (k + 0.5) / nz
        r = np.sqrt((px - 0.5) ** 2 + (py - 0.5) ** 2 + (pz - 0.5) ** 2)
        s1 = (r - 0.30) / 0.04
        s2 = (r - 0.15) / 0.03
        f = 0.70 * (0.6 * np.exp(-(s1 * s1)) + 0.4 * np.exp(-(s2 * s2))) \
            + 0.25 * (0.5 + 0.5 * sx * sy * math.sin(two_pi * (5.0 * pz + 0.3)))
        if with_hash:
            h = hx ^ hy ^ np.uint32((k * 83492791) & 0xFFFFFFFF) ^ np.uint32(seed)
            h = h * np.uint32(2654435761)
            f = f + 0.05 * (h.astype(np.float64) / 4294967296.0)
        f = np.clip(f, 0.0, 1.0)
        out[k] = np.floor(vmax * f + 0.5).astype(dtype)
    return out.reshape(-1)


# The three camera blocks as bit patterns (what camera_block() returns; tests/test_host.py keeps them in sync), so
# that a CPU-only consumer -- bench.py --impl reference -- needs neither the host library nor the CUDA library.
_CAMERA_BITS = {
    "K0": [0x3f800000, 0x0, 0x0, 0x0, 0x0, 0x3f800000, 0x0, 0x0, 0x80000000, 0x80000000, 0x3f800000, 0x80000000,
           0x0, 0x0, 0x40400000, 0x3f800000, 0x0, 0x0, 0x40400000, 0x3f800000, 0x406ed9ec],
    "K1": [0x3f51b3f3, 0x0, 0xbf12d5e7, 0x0, 0xbe92d5e6, 0x3f5db3d7, 0xbed1b3f2, 0x0, 0x3efe53a0, 0x3efffffe, 0x3f359baa,
           0x80000000, 0x3fbebeb9, 0x3fbfffff, 0x400834c0, 0x3f800000, 0x3fbebeb9, 0x3fbfffff, 0x400834c0, 0x3f800000, 0x406ed9ec],
    "K2": [0xbf708fb3, 0x0, 0x3eaf1d44, 0x0, 0x3def920c, 0x3f708fb2, 0x3ea48dbb, 0x0, 0xbea48dba, 0x3eaf1d43, 0xbf620dbe,
           0x80000000, 0xbf03a496, 0x3f0c176a, 0xbfb4d7cc, 0x3f800000, 0xbf03a496, 0x3f0c176a, 0xbfb4d7cc, 0x3f800000, 0x406ed9ec],
}


def camera_block_const(kind: str) -> np.ndarray:
    """camera_block(kind) from the stored bit patterns: no native library involved."""
    return np.array(_CAMERA_BITS[kind], dtype=np.uint32).view(np.float32).copy()


def camera_block(kind: str) -> np.ndarray:
    """The 21-float camera block for K0 (reset), K1 (zenith 60, azimuth 35, r 3), K2 (zenith 70,
    azimuth 200, r 1.6), built by the product's host Camera."""
    from .host import Camera
    cam = Camera(30.0)
    if kind == "K0":
        pass
    elif kind == "K1":
        cam.setSpherical(3.0, math.radians(60.0), math.radians(35.0))
    elif kind == "K2":
        cam.setSpherical(1.6, math.radians(70.0), math.radians(200.0))
    else:
        raise ValueError(kind)
    return cam.ubo()


CONFIGS = {
    # name: dims, bytes/voxel, vmax, image, step_scale (reference step * scale), notes
    # `window` = the uniform values of SURVEY.md 8(d)'s table (C3: GUI 0..2000 + the reference's +1000 rule)
    "C1": dict(dims=(64, 64, 64), bpv=1, vmax=255, image=(256, 256), step_scale=0.5, window=(0, 255)),
    "C2": dict(dims=(256, 256, 256), bpv=1, vmax=255, image=(1024, 1024), step_scale=0.5, window=(0, 255)),
    "C3": dict(dims=(512, 512, 512), bpv=2, vmax=4095, image=(1920, 1080), step_scale=0.5, window=(1000, 3000)),
    "C4": dict(dims=(1024, 1024, 1024), bpv=2, vmax=4095, image=(1920, 1080), step_scale=1.0, window=(0, 4095)),
    "C5": dict(dims=(2048, 2048, 1024), bpv=2, vmax=4095, image=(3840, 2160), step_scale=1.0, window=(0, 4095)),
}
