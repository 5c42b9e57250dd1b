"""-m gpu: the RendererCore mirror (host/RendererCore.cpp) driven the way RendererGUI drives the
reference's RendererCore: setup -> loadShader -> menu choice of UINT8/UINT16 -> readVolumeData
(.raw + .raw.inf, .pvm) -> sliders -> render -> saveImage; every frame is compared with the oracle."""
import os

import numpy as np
import pytest

import volren_b200 as vb
from oracle import orc
from volren_b200 import host, workloads

from util import compare, oracle_frame

pytestmark = pytest.mark.gpu


def test_gui_session_raw_uint8(tmp_path):
    dims = (48, 40, 36)
    vox = workloads.mix_volume(dims, 255, 77)
    fn = str(tmp_path / "bonsai_like.raw")
    vox.tofile(fn)
    W, H = 320, 180                              # window_size / framebuffer_size (RendererGUI.cpp:38-39)
    core = host.RendererCore(W, H)
    core.setup()
    assert core.loadShader("VolumeRenderer.cs")
    s = core.state()
    assert (s.workgroups_x, s.workgroups_y) == (W // 16, H // 16)          # RendererCore.cpp:121-122
    # menu: "Load RAW -> UINT8"; no sidecar yet -> the modal asks for dims + spacing and a sidecar is written
    core.set_datasize_bytes(1)
    assert not core.checkRawInfFile(fn)
    core.set_raw_info(dims, (1.0, 1.0, 1.2))
    core.readVolumeData(fn)
    st = core.strings()
    assert (st["title"], st["msg"]) == ("File Loaded!", "File Loaded Successfully!")     # RendererCore.cpp:436-437
    assert st["loaded_dataset"] == "bonsai_like.raw"
    assert open(fn + ".inf").read() == "#dimensions\n48 40 36\n\n#voxel-spacing\n1 1 1.2\n"
    core.clear_popup()
    s = core.state()
    assert (s.min_val, s.max_val, s.min_dataset_val, s.max_dataset_val) == (0, 255, 0, 255)   # :380-384
    # first frame with the constructor defaults: alpha 1, window 0..255, nearest
    core.render()
    frame = core.readFrame()
    cam = core.camera_ubo()
    assert cam == pytest.approx(workloads.camera_block("K0"))
    ref, _ = oracle_frame(cam, vox, dims, 1, W, H, voxel_size=(1.0, 1.0, 1.2), alpha_scale=1.0, min_val=0, max_val=255, filter=0)
    compare(frame, ref, "first frame")
    assert core.state().kerneltime_sum > 0                                  # RendererCore.cpp:153
    core.reset_kerneltime()
    # sliders + mouse drag, like RendererGUI.cpp:336-358,382-386 and GlfwManager.cpp:179
    core.gui_alpha(0.05); core.gui_min(30); core.gui_max(180)
    for _ in range(6):
        core.camera_setOrientation(0, 0.06, -0.06)
    core.camera_setOrientation(1, 0, 0)
    core.ext_filter(vb.FILTER_TRILINEAR)
    core.render()
    frame = core.readFrame()
    ref, _ = oracle_frame(core.camera_ubo(), vox, dims, 1, W, H, voxel_size=(1.0, 1.0, 1.2), alpha_scale=0.05, min_val=30, max_val=180, filter=1)
    compare(frame, ref, "after sliders")
    # MIP toggle and the top view (which also resets the camera, RendererCore.cpp:94)
    core.gui_mip(True); core.gui_view(True, False)
    core.render()
    ref, _ = oracle_frame(workloads.camera_block("K0"), vox, dims, 1, W, H, voxel_size=(1.0, 1.0, 1.2), alpha_scale=0.05,
                          min_val=30, max_val=180, filter=1, is_mip=1, view_top=1)
    compare(core.readFrame(), ref, "mip + top view")
    # save image: RGB8, vertically flipped (RendererCore.cpp:165-182)
    png = str(tmp_path / "shot.png")
    assert core.saveImage(png, ".png")
    from test_host import _png_pixels
    img = core.readFrame()
    expect = np.rint(np.clip(img[..., :3], 0, 1) * np.float32(255)).astype(np.uint8)[::-1]
    assert np.array_equal(_png_pixels(png), expect)
    # reload: the sidecar now exists and wins over the modal values
    core.set_raw_info((1, 1, 1), (9, 9, 9))
    assert core.checkRawInfFile(fn)
    core.readVolumeData(fn)
    s = core.state()
    assert tuple(s.tex3D_dim) == dims and tuple(s.voxel_size) == pytest.approx((1.0, 1.0, 1.2))


def test_gui_session_pvm_uint16_plus1000_rule(tmp_path, golden_dir):
    if orc.ref_lib() is None:
        pytest.skip("oracle/_ref not available to write a 16-bit .pvm")
    dims = (40, 40, 24)
    vox = workloads.mix_volume(dims, 3000, 5)
    fn = str(tmp_path / "ct_like.pvm")
    orc.ref_write_pvm(fn, vox, dims, 2, (1.0, 1.0, 1.5))
    W, H = 256, 144
    core = host.RendererCore(W, H)
    core.setup()
    core.loadShader("VolumeRenderer.cs")
    core.set_datasize_bytes(2)                     # menu: "Load PVM -> UINT16"
    core.readVolumeData(fn)
    assert core.strings()["title"] == "File Loaded!"
    s = core.state()
    assert tuple(s.tex3D_dim) == dims and tuple(s.voxel_size) == pytest.approx((1.0, 1.0, 1.5))
    assert (s.min_dataset_val, s.max_dataset_val) == (int(vox.min()), int(vox.max()))   # RendererCore.cpp:362-379
    p = core.params()
    assert (p.min_val, p.max_val) == (int(vox.min()) + 1000, int(vox.max()) + 1000)      # :66-69,77-80
    core.gui_min(0); core.gui_max(1500); core.gui_alpha(0.1)
    core.render()
    ref, _ = oracle_frame(core.camera_ubo(), vox, dims, 2, W, H, voxel_size=(1.0, 1.0, 1.5), alpha_scale=0.1,
                          min_val=1000, max_val=2500, filter=0)
    compare(core.readFrame(), ref, "16-bit pvm")
    # histogram: RendererCore.cpp:386-405 -- bin = round(v*255/max), bin 0 skipped, normalised by
    # max(dataset max value, largest bin count) (the reference re-uses `max_value`)
    h = core.histogram()
    x = vox.astype(np.float32) * np.float32(255.0) / np.float32(vox.max())
    b = np.trunc(x); b = (b + ((x - b) >= np.float32(0.5))).astype(np.int64)
    counts = np.bincount(b, minlength=256).astype(np.float32); counts[0] = 0
    expect = counts * np.float32(100.0) / np.float32(max(int(vox.max()), int(counts.max())))
    assert np.array_equal(h, expect) and h[0] == 0


def test_load_errors_surface_as_popups(tmp_path):
    core = host.RendererCore(64, 64)
    core.setup()
    core.loadShader("VolumeRenderer.cs")
    core.set_datasize_bytes(1)
    core.readVolumeData(str(tmp_path / "nope.pvm"))
    st = core.strings()
    assert (st["title"], st["msg"]) == ("Error!", "Error reading PVM file")     # RendererCore.cpp:349-353
    core.clear_popup()
    fn = str(tmp_path / "v.raw")
    open(fn + ".inf", "w").write("#dimensions\n0 0 0\n#voxel-spacing\n1 1 1\n")
    core.readVolumeData(fn)
    assert core.strings()["title"] == "Invalid .raw.inf file!"
