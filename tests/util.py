"""Shared helpers for the tests: scenario tables and the bit-exact comparison."""
from __future__ import annotations

import numpy as np

from oracle import orc

# The tolerance BASELINE.json's north_star states for this floating-point path.  The CUDA path
# is engineered to be bit-identical to the oracle, so the tests check BOTH: max |diff| <= TOL
# (the contract) and exact bit equality (the stronger property we rely on for N-GPU == 1-GPU).
TOL = 1e-4


def compare(img: np.ndarray, ref: np.ndarray, what: str = "", exact: bool = True):
    assert img.shape == ref.shape and img.dtype == np.float32 and ref.dtype == np.float32
    both_nan = np.isnan(img) & np.isnan(ref)
    diff = np.abs(np.where(both_nan, 0.0, img.astype(np.float64) - ref.astype(np.float64)))
    diff = np.where(np.isnan(diff), np.inf, diff)
    max_err = float(diff.max()) if diff.size else 0.0
    n_bad = int((diff > TOL).sum())
    assert max_err <= TOL, f"{what}: max |diff| {max_err:.3e} > {TOL} on {n_bad} values"
    if exact:
        same = (img.view(np.uint32) == ref.view(np.uint32)) | both_nan | ((img == 0) & (ref == 0))
        assert same.all(), f"{what}: {int((~same).sum())} values differ in the last bits (max |diff| {max_err:.3e})"
    return max_err


def oracle_frame(cam21, vox, dims, bpv, W, H, nthreads=8, **kw):
    p = orc.make_params(W, H, dims, bpv, cam21, **kw)
    img, cnt, _ = orc.render(p, vox, nthreads=nthreads)
    return img, cnt
