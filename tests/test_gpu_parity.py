"""-m gpu: the CUDA path (through the C-ABI) against the CPU oracle -- and, where the reference's own shader
compiled for the CPU is present (oracle/_ref/libshader_ref.so travels with the snapshot), directly against
the reference -- on the same seeded inputs.

Bar: north_star tolerance 1e-4 per channel (util.TOL) -- and, because the device code issues
the same correctly-rounded operation sequence as the oracle, exact bit equality."""
import os

import numpy as np
import pytest

import volren_b200 as vb
from oracle import orc

import scenarios
from util import TOL, compare, oracle_frame

pytestmark = pytest.mark.gpu

OPTIMISED = (vb.KERNEL_TEXPAIR_PIPE, vb.KERNEL_NEAREST_TEX)


def run_product(vox, dims, voxel_size, cam, W, H, vkw, kernel=vb.KERNEL_AUTO, partition=None, skip=vb.SKIP_AUTO):
    with vb.Context(W, H) as ctx:
        ctx.upload_volume(vox, dims, voxel_size)
        ctx.set_camera(cam)
        ctx.set_params(vb.default_params(kernel=kernel, empty_skip=skip, **vkw))
        if partition:
            ctx.set_partition(*partition)
        img, st = ctx.render()
        return img, st


def expected_kernel(kw):
    return vb.KERNEL_TEXPAIR_PIPE if kw.get("filter", 0) == 1 else vb.KERNEL_NEAREST_TEX


# scenarios the optimised kernels do not cover by design: they must take the generic loop
DEGENERATE = {"window_min_eq_max_nan", "window_min_gt_max"}
# volumes deeper than a layered CUDA array (2048 layers)
TOO_DEEP = {"iteration_cap_reference"}


@pytest.mark.parametrize("cid", [c[0] for c in scenarios.CASES])
@pytest.mark.parametrize("mode", ["direct", "auto", "auto_skip"])
def test_case_matches_oracle(cid, mode):
    _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    okw, vkw = scenarios.split_kwargs(kw)
    ref, cnt = oracle_frame(cam, vox, dims, bpv, W, H, voxel_size=vs, **okw)
    kernel = vb.KERNEL_DIRECT if mode == "direct" else vb.KERNEL_AUTO
    skip = vb.SKIP_ON if mode == "auto_skip" else vb.SKIP_OFF
    img, st = run_product(vox, dims, vs, cam, W, H, vkw, kernel=kernel, skip=skip)
    assert st.kernel_launches >= 1 and st.kernel_ms > 0
    if mode == "direct" or cid in DEGENERATE or cid in TOO_DEEP:
        assert st.kernel_used == vb.KERNEL_DIRECT
    else:
        assert st.kernel_used == expected_kernel(kw), f"{cid}: fell back to kernel {st.kernel_used}"
    # tolerance-only: per-sample opacity correction evaluates a double-precision pow (CUDA vs glibc)
    compare(img, ref, cid, exact=cid not in scenarios.TOLERANCE_ONLY)
    if cid not in ("window_min_eq_max_nan",):
        assert cnt["rays_hit"] > 0 and np.isfinite(img).all()


@pytest.mark.skipif(orc.ref_shader_lib() is None, reason="oracle/_ref/libshader_ref.so not built")
@pytest.mark.parametrize("cid", [c[0] for c in scenarios.CASES if scenarios.is_reference_semantics(c[4])])
def test_case_matches_the_reference_shader_itself(cid):
    """No restatement in between: the CUDA frame against VolumeRenderer.cs compiled for the CPU."""
    _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    okw, vkw = scenarios.split_kwargs(kw)
    ref, _ = orc.ref_render(orc.make_params(W, H, dims, bpv, cam, voxel_size=vs, **okw), vox, nthreads=8)
    img, _ = run_product(vox, dims, vs, cam, W, H, vkw)
    compare(img, ref, "reference shader " + cid)


@pytest.mark.parametrize("cid", ["c1_trilinear_128steps", "ragged_nearest", "u16_aniso_trilinear", "mip_nearest"])
def test_golden_fixture(cid, golden_dir):
    """Committed fixtures (tests/golden/make_golden.py): inputs are regenerated from the seed,
    the expected image comes from the file, not from running the oracle now."""
    g = np.load(os.path.join(golden_dir, f"{cid}.npz"))
    _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
    vox, dims, bpv, vs = scenarios.volume(vname)
    assert int(g["voxel_crc"]) == int(np.bitwise_xor.reduce(vox.astype(np.uint32) * np.arange(1, vox.size + 1, dtype=np.uint32)))
    cam = g["cam"]
    _, vkw = scenarios.split_kwargs(kw)
    img, _ = run_product(vox, dims, vs, cam, W, H, vkw)
    compare(img, g["rgba"], "golden " + cid)


@pytest.mark.parametrize("cid", ["c1_nearest_ref_step", "ragged_trilinear", "u16_aniso_trilinear", "mip_nearest", "view_top",
                                 "window_min_gt_max", "iteration_cap_reference"])
def test_golden_fixture_written_by_the_reference_shader(cid, golden_dir):
    """tests/golden/ref_*.npz (tests/golden/make_ref_golden.py): images the reference's own shader produced."""
    g = np.load(os.path.join(golden_dir, f"ref_{cid}.npz"))
    _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
    vox, dims, bpv, vs = scenarios.volume(vname)
    _, vkw = scenarios.split_kwargs(kw)
    for skip in (vb.SKIP_OFF, vb.SKIP_ON):
        img, _ = run_product(vox, dims, vs, g["cam"], W, H, vkw, skip=skip)
        compare(img, g["rgba"], "reference golden " + cid)


# ---------------------------------------------------------------- every mode combination runs in the optimised kernels

COMBOS = [
    # nearest filter (the reference's de-facto filter) with the GUI's toggles
    ("mix64_u8", "K1", dict(alpha_scale=0.9, min_val=0, max_val=255, filter=0, is_mip=1)),
    ("mix64_u8", "K1", dict(alpha_scale=0.3, min_val=0, max_val=255, filter=0, tf=True)),
    ("rand_40x56x33_u16", "K0", dict(alpha_scale=0.05, min_val=0, max_val=4095, filter=0, view_bottom=1)),
    ("rand_48x40x36_u8", "K1", dict(alpha_scale=0.08, min_val=0, max_val=255, filter=0, view_top=1)),
    ("mix64_u8", "K0", dict(alpha_scale=0.9, min_val=0, max_val=255, filter=0, tf=True, is_mip=1)),
    ("mix_64x64x32_u16", "K2", dict(alpha_scale=0.7, min_val=1000, max_val=3000, filter=0, is_mip=1, view_top=1)),
    # trilinear: TF / MIP / views alone and combined
    ("mix64_u8", "K1", dict(alpha_scale=0.3, min_val=0, max_val=255, filter=1, tf=True)),
    ("mix64_u8", "K2", dict(alpha_scale=0.08, min_val=30, max_val=200, filter=1, tf=True, step_scale=0.5)),
    ("mix_64x64x32_u16", "K1", dict(alpha_scale=0.2, min_val=100, max_val=3900, filter=1, tf=True)),
    ("mix_64x64x32_u16", "K2", dict(alpha_scale=1.0, min_val=1000, max_val=3000, filter=1, is_mip=1)),
    ("rand_40x56x33_u16", "K0", dict(alpha_scale=0.9, min_val=0, max_val=4095, filter=1, is_mip=1, step_scale=0.5)),
    ("rand_48x40x36_u8", "K1", dict(alpha_scale=0.08, min_val=0, max_val=255, filter=1, view_top=1)),
    ("rand_40x56x33_u16", "K0", dict(alpha_scale=0.05, min_val=0, max_val=4095, filter=1, view_bottom=1)),
    ("mix64_u8", "K2", dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1, view_top=1, view_bottom=1)),   # top wins (:186)
    ("mix64_u8", "K1", dict(alpha_scale=0.8, min_val=0, max_val=255, filter=1, tf=True, is_mip=1)),
    ("mix64_u8", "K1", dict(alpha_scale=0.3, min_val=10, max_val=240, filter=1, tf=True, view_top=1)),
    ("mix_64x64x32_u16", "K0", dict(alpha_scale=0.9, min_val=1000, max_val=3000, filter=1, is_mip=1, view_bottom=1)),
    # opacity correction with a transfer function: final opacity LUT baked on the host -> exact
    ("smooth64_u8", "K1", dict(alpha_scale=0.4, min_val=0, max_val=255, filter=1, tf=True, step_scale=0.5, opacity_correction=1)),
    ("mix64_u8", "K2", dict(alpha_scale=0.4, min_val=0, max_val=255, filter=0, tf=True, step_scale=2.0, opacity_correction=1)),
]


@pytest.mark.parametrize("vname,cname,kw", COMBOS)
def test_mode_combinations_run_in_the_optimised_kernels(vname, cname, kw):
    """VolumeRenderer.cs:141-173 (MIP), :183-190 (view swizzles), the TF extension and their combinations run in
    the optimised kernels -- no silent fall-back to the generic loop -- and equal the oracle and the generic
    DIRECT loop bit for bit."""
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    W, H = 320, 200
    okw, vkw = scenarios.split_kwargs(kw)
    ref, _ = oracle_frame(cam, vox, dims, bpv, W, H, voxel_size=vs, **okw)
    for skip in (vb.SKIP_OFF, vb.SKIP_ON):
        img, st = run_product(vox, dims, vs, cam, W, H, vkw, skip=skip)
        assert st.kernel_used == expected_kernel(kw)
        compare(img, ref, f"{vname}/{cname}/{kw}")
    direct, st2 = run_product(vox, dims, vs, cam, W, H, vkw, kernel=vb.KERNEL_DIRECT)
    assert st2.kernel_used == vb.KERNEL_DIRECT
    assert np.array_equal(direct.view(np.uint32), img.view(np.uint32))


@pytest.mark.parametrize("filt", [0, 1])
def test_per_sample_opacity_correction_stays_in_the_optimised_kernel(filt):
    """a' = 1-(1-a)^step_scale without a transfer function: evaluated per sample in double precision inside the
    optimised kernel (CUDA pow vs glibc pow: held to the north-star tolerance, not to bit equality)."""
    vox, dims, bpv, vs = scenarios.volume("smooth64_u8")
    cam = scenarios.camera("K1")
    kw = dict(alpha_scale=0.1, min_val=0, max_val=255, filter=filt, step_scale=0.5, opacity_correction=1)
    ref, _ = oracle_frame(cam, vox, dims, bpv, 200, 150, voxel_size=vs, **kw)
    img, st = run_product(vox, dims, vs, cam, 200, 150, kw)
    assert st.kernel_used == expected_kernel(kw)
    compare(img, ref, "opacity correction", exact=False)
    # alpha_scale > 1 can push (1 - a) below zero: NaN semantics belong to the generic loop
    kw["alpha_scale"] = 1.5
    ref, _ = oracle_frame(cam, vox, dims, bpv, 200, 150, voxel_size=vs, **kw)
    img, st = run_product(vox, dims, vs, cam, 200, 150, kw)
    assert st.kernel_used == vb.KERNEL_DIRECT
    both_nan = np.isnan(img) & np.isnan(ref)
    assert np.allclose(np.where(both_nan, 0, img), np.where(both_nan, 0, ref), atol=TOL, equal_nan=True)


@pytest.mark.parametrize("kw", [
    dict(alpha_scale=-0.5, min_val=0, max_val=255, filter=1, is_mip=1),                 # MIP must not render as DVR
    dict(alpha_scale=-0.5, min_val=0, max_val=255, filter=1, view_top=1),
    dict(alpha_scale=-0.25, min_val=0, max_val=255, filter=0, tf=True),
    dict(alpha_scale=0.5, min_val=200, max_val=100, filter=1, is_mip=1),                # unordered window + MIP
    dict(alpha_scale=0.5, min_val=100, max_val=100, filter=0, view_bottom=1),           # 0/0 window + view
])
def test_frames_outside_the_fast_path_keep_their_mode(kw):
    """Negative alpha_scale / degenerate windows are not covered by the optimised kernels; the generic loop that
    takes over must still honour MIP, the transfer function and the view swizzles."""
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K1")
    okw, vkw = scenarios.split_kwargs(kw)
    ref, _ = oracle_frame(cam, vox, dims, bpv, 160, 120, voxel_size=vs, **okw)
    img, st = run_product(vox, dims, vs, cam, 160, 120, vkw)
    assert st.kernel_used == vb.KERNEL_DIRECT
    compare(img, ref, str(kw))


# ---------------------------------------------------------------- empty-space skipping + the cell table

def shells_volume(n=96, bpv=2):
    """CT-like: two bright shells in 'air' (values below the window), ~75 % of the voxels empty under [1000, 3000]."""
    from volren_b200 import workloads
    vox = workloads.mix_volume((n, n, n), 4095 if bpv == 2 else 255, 0x5EED0077)
    return vox, (n, n, n)


@pytest.mark.parametrize("filt", [0, 1])
@pytest.mark.parametrize("cname", ["K0", "K1", "K2", "inside"])
def test_empty_space_skipping_is_result_identical(filt, cname):
    vox, dims = shells_volume()
    cam = scenarios.camera(cname)
    W, H = 384, 216
    kw = dict(alpha_scale=0.08, min_val=1000, max_val=3000, filter=filt)
    ref, cnt = oracle_frame(cam, vox, dims, 2, W, H, **kw)
    on, st_on = run_product(vox, dims, (1, 1, 1), cam, W, H, kw, skip=vb.SKIP_ON)
    off, st_off = run_product(vox, dims, (1, 1, 1), cam, W, H, kw, skip=vb.SKIP_OFF)
    auto, st_auto = run_product(vox, dims, (1, 1, 1), cam, W, H, kw)
    assert st_on.skip_used == 1 and st_off.skip_used == 0
    assert st_auto.skip_used == 1, "AUTO should skip: most cells are empty under this window"
    for img in (on, off, auto):
        compare(img, ref, f"skip {filt}/{cname}")


def test_skipping_is_refused_when_an_empty_sample_still_contributes():
    """A transfer function whose first entry is not zero gives v = 0 samples a non-zero opacity: no skipping; a
    negative min_val cannot be compared with unsigned cell maxima: no skipping."""
    vox, dims = shells_volume(64)
    cam = scenarios.camera("K1")
    lut = scenarios.tf_lut().copy()
    lut[0] = 0.25
    for filt in (0, 1):
        kw = dict(alpha_scale=0.3, min_val=1000, max_val=3000, filter=filt, tf_lut=lut)
        ref, _ = oracle_frame(cam, vox, dims, 2, 200, 150, **kw)
        img, st = run_product(vox, dims, (1, 1, 1), cam, 200, 150, kw, skip=vb.SKIP_ON)
        assert st.skip_used == 0 and st.kernel_used in OPTIMISED
        compare(img, ref, "tf lut[0] != 0")
        kw = dict(alpha_scale=0.3, min_val=-5, max_val=3000, filter=filt)
        ref, _ = oracle_frame(cam, vox, dims, 2, 200, 150, **kw)
        img, st = run_product(vox, dims, (1, 1, 1), cam, 200, 150, kw, skip=vb.SKIP_ON)
        assert st.skip_used == 0
        compare(img, ref, "negative min_val")


@pytest.mark.parametrize("dims,bpv", [((40, 33, 20), 1), ((70, 17, 130), 2), ((8, 8, 8), 2), ((1, 1, 1), 1), ((300, 200, 9), 2),
                                      # 16-byte aligned rows: the fused ingest kernel (padded copy + cell table in one read)
                                      ((64, 40, 33), 2), ((128, 24, 20), 1), ((16, 9, 9), 1), ((8, 1, 1), 2), ((264, 31, 17), 2),
                                      ((512, 512, 256), 1), ((1024, 130, 70), 2)])
def test_cell_table_matches_numpy(dims, bpv):
    """Per-cell min/max (SURVEY 8f-2's brick table): cell c covers voxel indices [c*2^s - 1, (c+1)*2^s - 1] per axis."""
    rng = np.random.default_rng(5)
    nx, ny, nz = dims
    vox = rng.integers(0, 256 if bpv == 1 else 65536, nx * ny * nz).astype(np.uint8 if bpv == 1 else np.uint16)
    with vb.Context(16, 16) as ctx:
        ctx.upload_volume(vox, dims)
        ctx.set_params(vb.default_params(min_val=100, max_val=200))
        t = ctx.cell_table()
    s = t["shift"]
    assert t["cells"] == ((nx >> s) + 1, (ny >> s) + 1, (nz >> s) + 1) and s >= 3
    v = vox.reshape(nz, ny, nx)
    side = 1 << s
    for cz in range(t["cells"][2]):
        for cy in range(t["cells"][1]):
            for cx in range(t["cells"][0]):
                sl = tuple(slice(min(max(c * side - 1, 0), n - 1), min((c + 1) * side - 1, n - 1) + 1)
                           for c, n in ((cz, nz), (cy, ny), (cx, nx)))
                blk = v[sl]
                assert t["mins"][cz, cy, cx] == blk.min() and t["maxs"][cz, cy, cx] == blk.max(), (cx, cy, cz)
    assert t["empty_cells"] == int((t["maxs"] <= 100).sum())


def test_volume_copies_are_built_on_demand():
    """Footprint: the linear copy always; each layered array only once a frame needs it."""
    vox, dims, bpv, vs = scenarios.volume("mix_64x64x32_u16")
    nvox = int(np.prod(dims))
    with vb.Context(64, 64) as ctx:
        ctx.upload_volume(vox, dims, vs)
        ctx.set_camera(scenarios.camera("K0"))
        m = ctx.memory_info()
        assert m["linear_bytes"] > 0 and m["array_bytes"] == 0 and m["zpair_array_bytes"] == 0
        ctx.set_params(vb.default_params(alpha_scale=0.1, min_val=0, max_val=4095, filter=1))
        ctx.render()
        m = ctx.memory_info()
        assert m["zpair_array_bytes"] == dims[0] * dims[1] * (dims[2] + 1) * 4 and m["array_bytes"] == 0
        ctx.set_params(vb.default_params(alpha_scale=0.1, min_val=0, max_val=4095, filter=0))
        ctx.render()
        assert ctx.memory_info()["array_bytes"] == nvox * 2
        ctx.upload_volume(vox, dims, vs)               # a new upload drops the derived copies
        m = ctx.memory_info()
        assert m["array_bytes"] == 0 and m["zpair_array_bytes"] == 0


# ---------------------------------------------------------------- instrumentation, partitions

def test_counters_and_distinct_voxels_match_oracle():
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K1")
    for filt in (0, 1):
        p = orc.make_params(160, 120, dims, bpv, cam, alpha_scale=0.05, min_val=0, max_val=255, filter=filt)
        _, cnt, tb = orc.render(p, vox, nthreads=4, touch=True)
        with vb.Context(160, 120) as ctx:
            ctx.upload_volume(vox, dims, vs)
            ctx.set_camera(cam)
            ctx.set_params(vb.default_params(alpha_scale=0.05, min_val=0, max_val=255, filter=filt))
            got = ctx.count_frame()
        assert got["samples"] == cnt["samples"]
        assert got["rays_hit"] == cnt["rays_hit"]
        assert got["distinct_voxels"] == orc.popcount(tb, vox.size)


@pytest.mark.parametrize("world,tile_rows", [(2, 8), (3, 4), (4, 16), (8, 8), (2, 1)])
def test_row_tile_partition_equals_single_gpu(world, tile_rows):
    """N-GPU result == 1-GPU result: every rank's compact tiles, assembled, are bit-identical
    to the unpartitioned frame (all ranks emulated on one device)."""
    import torch
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K1")
    W, H = 200, 150        # H not a multiple of tile_rows*world
    kw = dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1)
    full, _ = run_product(vox, dims, vs, cam, W, H, kw)
    with vb.Context(W, H) as ctx:
        ctx.upload_volume(vox, dims, vs)
        ctx.set_camera(cam)
        ctx.set_params(vb.default_params(**kw))
        ctx.set_partition(0, world, tile_rows)
        rows = ctx.owned_rows()
        gathered = torch.full((world, rows, W, 4), float("nan"), device="cuda", dtype=torch.float32)
        acc = np.zeros((H, W, 4), dtype=np.float32)
        for rank in range(world):
            ctx.set_partition(rank, world, tile_rows)
            assert ctx.owned_rows() == rows
            ctx.render_device(gathered[rank].data_ptr(), compact=True)
            part, _ = ctx.render()              # full-frame form: rows not owned are zero
            owned = ((np.arange(H) // tile_rows) % world) == rank
            assert (part[~owned] == 0).all()
            acc[owned] = part[owned]
        frame = torch.empty((H, W, 4), device="cuda", dtype=torch.float32)
        ctx.assemble_tiles(gathered.data_ptr(), frame.data_ptr(), world, tile_rows)
        torch.cuda.synchronize()
    assert np.array_equal(frame.cpu().numpy().view(np.uint32), full.view(np.uint32))
    assert np.array_equal(acc.view(np.uint32), full.view(np.uint32))


def test_rank_without_any_row_tile_renders_nothing_and_succeeds():
    """tiles < world (H = 16, one tile of 16 rows, 2 ranks): rank 1 owns nothing; every render entry point
    must return VR_OK instead of failing on an empty grid."""
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    W, H = 64, 16
    with vb.Context(W, H) as ctx:
        ctx.upload_volume(vox, dims, vs)
        ctx.set_camera(scenarios.camera("K0"))
        ctx.set_params(vb.default_params(alpha_scale=0.1, min_val=0, max_val=255, filter=1))
        ctx.set_partition(1, 2, 16)
        img, st = ctx.render()
        assert st.kernel_launches == 0 and not img.any()
        st = ctx.render_device(0)
        assert st.kernel_launches == 0
        host = np.full((H, W, 4), 7.0, np.float32)
        ctx.render_owned_to_host_ptr(host.ctypes.data)
        assert (host == 7.0).all()


@pytest.mark.parametrize("world,tile_rows,W,H", [(2, 16, 200, 150), (3, 4, 200, 150), (8, 16, 200, 150), (2, 8, 1280, 720), (4, 8, 1000, 500)])
def test_owned_tiles_to_host_frame_assemble_the_frame(world, tile_rows, W, H):
    """Multi-GPU end to end (vr_render_owned_to_host): every rank copies only its own row tiles into a
    full host frame; all ranks together (emulated on one device) reproduce the unpartitioned frame.  The
    larger frames take the banded path (copies overlap the march of the following bands)."""
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K1")
    kw = dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1)
    full, _ = run_product(vox, dims, vs, cam, W, H, kw)
    host = np.full((H, W, 4), np.nan, dtype=np.float32)
    with vb.Context(W, H) as ctx:
        ctx.upload_volume(vox, dims, vs)
        ctx.set_camera(cam)
        ctx.set_params(vb.default_params(**kw))
        for rank in range(world):
            ctx.set_partition(rank, world, tile_rows)
            before = host.copy()
            st = ctx.render_owned_to_host_ptr(host.ctypes.data)
            assert st.kernel_launches >= 1
            owned = ((np.arange(H) // tile_rows) % world) == rank
            # rows of other ranks are untouched
            assert np.array_equal(host[~owned].view(np.uint32), before[~owned].view(np.uint32))
    assert np.array_equal(host.view(np.uint32), full.view(np.uint32))


@pytest.mark.parametrize("W,H", [(131, 67), (1, 5), (65, 8), (1280, 724)])
def test_odd_image_sizes_and_banded_frames(W, H):
    """Widths below one CTA tile, heights that are not a multiple of the tile, and a frame large enough for the
    banded host copy."""
    vox, dims, bpv, vs = scenarios.volume("mix_64x64x32_u16")
    cam = scenarios.camera("K1")
    for filt in (0, 1):
        kw = dict(alpha_scale=0.05, min_val=0, max_val=4095, filter=filt)
        ref, _ = oracle_frame(cam, vox, dims, bpv, W, H, voxel_size=vs, **kw)
        img, st = run_product(vox, dims, vs, cam, W, H, kw)
        assert st.kernel_used == expected_kernel(kw)
        compare(img, ref, f"{W}x{H} filter {filt}")


def test_launch_order_table_follows_camera_partition_and_bands():
    """The march kernels start their CTA tiles longest rays first, from a table the GPU rebuilds whenever the camera,
    the box, the partition or the band changes (cta_order_kernel).  One context, a moving camera, changing partitions,
    device frames and banded host frames: every frame must stay bit-identical to the generic loop's (which never uses
    the table) -- a stale or incomplete table would leave tiles of an earlier frame behind."""
    vox, dims, bpv, vs = scenarios.volume("mix_64x64x32_u16")
    W, H = 1280, 724                                   # large enough for the banded host path
    cams = [scenarios.camera(k) for k in ("K0", "K1", "K2", "K1", "inside", "K0")]
    with vb.Context(W, H) as ctx:
        ctx.upload_volume(vox, dims, vs)
        for i, cam in enumerate(cams):
            filt = i & 1
            kw = dict(alpha_scale=0.05 + 0.01 * i, min_val=0, max_val=4095, filter=filt)
            world = (1, 2, 3, 1, 8, 1)[i]
            for rank in sorted({0, world - 1}):
                ctx.set_partition(rank, world, 8)
                ctx.set_camera(cam)
                ctx.set_params(vb.default_params(kernel=vb.KERNEL_DIRECT, **kw))
                ref, st = ctx.render()
                assert st.kernel_used == vb.KERNEL_DIRECT
                ctx.set_params(vb.default_params(**kw))
                img, st = ctx.render()                 # banded host frame: one table per band
                assert st.kernel_used == expected_kernel(kw)
                assert np.array_equal(img.view(np.uint32), ref.view(np.uint32)), f"host frame {i} rank {rank}/{world}"
                ctx.render_device(0)                   # whole partition in one launch: another table
                dev = ctx.read_frame()
                owned = ((np.arange(H) // 8) % world) == rank      # rows of other ranks keep whatever earlier frames left
                assert np.array_equal(dev[owned].view(np.uint32), ref[owned].view(np.uint32)), f"device frame {i} rank {rank}/{world}"


@pytest.mark.parametrize("W,H", [(1280, 724), (200, 150)])      # banded frames / one launch + tile copies
def test_pipelined_host_frames_equal_synchronous_ones(W, H):
    """vr_render_submit / vr_render_wait: two frames in flight into two host buffers, cameras changing every frame,
    a window change and a partition change in mid-flight (both wait for the frames in flight by themselves): every
    frame equals the one the synchronous call produces.  A third submit and a synchronous render are refused while
    tickets are outstanding."""
    vox, dims, bpv, vs = scenarios.volume("mix_64x64x32_u16")
    cams = [scenarios.camera(k) for k in ("K0", "K1", "K2", "K1", "K0", "K2", "K1")]
    kws = [dict(alpha_scale=0.05, min_val=0, max_val=4095, filter=1)] * 3 + [dict(alpha_scale=0.07, min_val=900, max_val=3000, filter=1)] * 2 + \
          [dict(alpha_scale=0.07, min_val=900, max_val=3000, filter=0)] * 2
    parts = [(0, 1)] * 5 + [(1, 2)] * 2
    with vb.Context(W, H) as ctx:
        ctx.upload_volume(vox, dims, vs)
        refs = []
        for cam, kw, (rank, world) in zip(cams, kws, parts):
            ctx.set_partition(rank, world, 8)
            ctx.set_camera(cam)
            ctx.set_params(vb.default_params(**kw))
            host = np.full((H, W, 4), np.float32(-1.0))
            ctx.render_owned_to_host_ptr(host.ctypes.data)
            refs.append(host)
        bufs = [np.empty((H, W, 4), np.float32), np.empty((H, W, 4), np.float32)]
        tickets = []
        for i, (cam, kw, (rank, world)) in enumerate(zip(cams, kws, parts)):
            ctx.set_partition(rank, world, 8)
            ctx.set_camera(cam)
            ctx.set_params(vb.default_params(**kw))
            if i >= 2:                                    # the buffer is about to be reused: its frame must be complete
                t, j = tickets.pop(0)
                st = ctx.render_wait(t)
                assert st.kernel_launches >= 1 and st.kernel_used == expected_kernel(kws[j])
                assert np.array_equal(bufs[j % 2].view(np.uint32), refs[j].view(np.uint32)), f"frame {j}"
            bufs[i % 2][:] = -1.0
            tickets.append((ctx.render_submit(bufs[i % 2].ctypes.data), i))
            if i == 1:
                with pytest.raises(vb.VolrenError):       # two frames in flight already
                    ctx.render_submit(bufs[0].ctypes.data)
                with pytest.raises(vb.VolrenError):       # the synchronous calls refuse to run meanwhile
                    ctx.render()
        for t, j in tickets:
            ctx.render_wait(t)
            assert np.array_equal(bufs[j % 2].view(np.uint32), refs[j].view(np.uint32)), f"frame {j}"
        with pytest.raises(vb.VolrenError):
            ctx.render_wait(tickets[-1][0])               # already collected
        img, _ = ctx.render()                             # and the synchronous path works again
        assert img.shape == (H, W, 4)


# ---------------------------------------------------------------- ingest

def test_volume_stats_match_reference_loops():
    """fused pad + min/max kernel and the histogram kernel vs a numpy restatement of RendererCore.cpp:360-405."""
    rng = np.random.default_rng(3)
    # (51, 41, 31): odd voxel count -> the scalar tail behind the 16-byte vector body; 65535: full u16 range
    for bpv, hi, dims in ((1, 256, (50, 40, 30)), (2, 3000, (50, 40, 30)), (1, 256, (51, 41, 31)), (2, 65536, (51, 41, 31)),
                          (1, 256, (48, 40, 30)), (2, 3000, (48, 40, 30)), (2, 65536, (256, 9, 11))):     # aligned rows: fused ingest kernel
        vox = rng.integers(0 if bpv == 1 else 17, hi, dims[0] * dims[1] * dims[2]).astype(np.uint8 if bpv == 1 else np.uint16)
        with vb.Context(32, 32) as ctx:
            ctx.upload_volume(vox, dims)
            mn, mx, hist = ctx.volume_stats()
        if bpv == 1:
            assert (mn, mx) == (0, 255)
            bins = vox.astype(np.int64)
            max_value = -1
        else:
            assert (mn, mx) == (int(vox.min()), int(vox.max()))
            x = vox.astype(np.float32) * np.float32(255.0) / np.float32(mx)
            scaled = np.trunc(x)
            scaled = scaled + ((x - scaled) >= np.float32(0.5))       # std::round: half away from zero
            bins = scaled.astype(np.int64)
            max_value = mx
        counts = np.bincount(bins, minlength=256).astype(np.float32)
        counts[0] = 0
        max_value = max(max_value, int(counts.max()))
        expect = counts * np.float32(100.0) / np.float32(max_value)
        assert np.array_equal(hist, expect.astype(np.float32))


def test_synthetic_generator_matches_numpy():
    from volren_b200 import workloads
    for dims, bpv, vmax, seed in (((40, 32, 24), 1, 255, workloads.SEEDS["C2"]), ((33, 20, 17), 2, 4095, workloads.SEEDS["C4"])):
        for with_hash in (True, False):
            dev = vb.synthetic_to_host(dims, bpv, vmax, seed, with_hash)
            ref = workloads.mix_volume(dims, vmax, seed, with_hash)
            d = np.abs(dev.astype(np.int64) - ref.astype(np.int64))
            assert d.max() <= 1 and (d != 0).mean() < 1e-4      # libm vs CUDA exp/sin in the last ulp


def test_upload_synthetic_renders_like_uploaded_copy():
    import torch
    dims, W, H = (48, 48, 48), 128, 96
    cam = scenarios.camera("K2")
    kw = dict(alpha_scale=0.05, min_val=0, max_val=4095, filter=1)
    with vb.Context(W, H) as ctx:
        copy = torch.empty(48 * 48 * 48, dtype=torch.int16, device="cuda")
        ctx.upload_synthetic(dims, 2, 4095, 99, True, copy_out_dptr=copy.data_ptr())
        ctx.set_camera(cam)
        ctx.set_params(vb.default_params(**kw))
        a, _ = ctx.render()
        host = copy.cpu().numpy().view(np.uint16)
    ref, _ = oracle_frame(cam, host, dims, 2, W, H, **kw)
    compare(a, ref, "synthetic upload")


def test_rgb8_readback_and_flip():
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K0")
    with vb.Context(96, 64) as ctx:
        ctx.upload_volume(vox, dims, vs)
        ctx.set_camera(cam)
        ctx.set_params(vb.default_params(alpha_scale=0.2, min_val=0, max_val=255, filter=1))
        img, _ = ctx.render()
        rgb = ctx.read_rgb8(flip_vertical=True)
        rgb_nf = ctx.read_rgb8(flip_vertical=False)
    expect = np.rint(np.clip(img[..., :3], 0, 1) * np.float32(255.0)).astype(np.uint8)
    assert np.array_equal(rgb_nf, expect)
    assert np.array_equal(rgb, expect[::-1])


def test_error_paths():
    with vb.Context(32, 32) as ctx:
        with pytest.raises(vb.VolrenError) as e:
            ctx.render()
        assert e.value.code == -3                      # VR_ERR_NO_VOLUME
        ctx.upload_volume(np.zeros(8, np.uint8), (2, 2, 2))
        with pytest.raises(vb.VolrenError):
            ctx.render()                               # no camera yet
        with pytest.raises(vb.VolrenError):
            ctx.set_camera([float("nan")] * 21)
        with pytest.raises(vb.VolrenError):
            ctx.set_params(vb.default_params(step_scale=0.0))
        with pytest.raises(vb.VolrenError):
            ctx.set_params(vb.default_params(kernel=3))     # a round-1 development kernel
        with pytest.raises(vb.VolrenError):
            ctx.set_params(vb.default_params(empty_skip=7))
        with pytest.raises(vb.VolrenError):
            ctx.set_partition(2, 2, 8)
