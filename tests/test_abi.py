"""The C-ABI library loads and exports every symbol include/volren_b200.h declares (no compute
calls: this file runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

import volren_b200 as vb


def header_symbols():
    hdr = open(os.path.join(vb.REPO_ROOT, "include", "volren_b200.h")).read()
    return sorted(set(re.findall(r"VR_API\s+[\w\s\*]+?\b(vr_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(vb.ABI_SYMBOLS) == syms


def test_library_exports_every_declared_symbol():
    L = C.CDLL(vb.LIB_PATH)
    for name in header_symbols():
        assert hasattr(L, name), f"libvolren_b200.so does not export {name}"


def test_library_is_self_contained():
    # static cudart: loading must not need libcuda / libcudart to be resolvable
    L = vb.lib()
    assert b"sm_100a" in L.vr_version()


def test_params_default_matches_reference_constructor():
    p = vb.default_params()
    # RendererCore.cpp:13-26: alpha_scale 1, min/max 0, MIP and view flags off
    assert p.alpha_scale == 1.0 and p.min_val == 0 and p.max_val == 0
    assert p.is_mip == 0 and p.view_top == 0 and p.view_bottom == 0
    assert p.filter == vb.FILTER_NEAREST and p.step_scale == 1.0 and p.use_tf == 0
    assert p.opacity_correction == 0 and p.kernel == vb.KERNEL_AUTO and p.empty_skip == vb.SKIP_AUTO


def test_struct_layouts_match_header():
    assert C.sizeof(vb.Params) == 4 * 10 + 256 * 4 + 4 + 4
    assert C.sizeof(vb.RenderStats) == 20
    assert C.sizeof(vb.VolumeStats) == 8 + 1024
    assert C.sizeof(vb.MemoryInfo) == 40


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vb.VolrenError) as e:
        vb.Context(64, 64)
    assert e.value.code == -2          # VR_ERR_CUDA: fails loudly, no CPU path
    assert "cuda" in str(e.value).lower()


def test_null_arguments_are_rejected_not_crashed():
    L = vb.lib()
    assert L.vr_create(0, 0, 0, None) == -1
    assert L.vr_set_camera(None, None) == -1
    assert L.vr_render(None, None, None) == -1
    assert L.vr_upload_volume(None, None, None, 1, None) == -1
    assert b"null" in L.vr_last_error()
    L.vr_destroy(None)   # no-op
