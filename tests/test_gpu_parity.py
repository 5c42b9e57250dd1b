"""-m gpu: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar: north_star tolerance 1e-4 per channel (util.TOL) -- and, because the device code issues
the same correctly-rounded operation sequence as the oracle, exact bit equality."""
import numpy as np
import pytest

import volren_b200 as vb
from oracle import orc

import scenarios
from util import TOL, compare, oracle_frame

pytestmark = pytest.mark.gpu


def run_product(vox, dims, voxel_size, cam, W, H, vkw, kernel=vb.KERNEL_AUTO, partition=None):
    with vb.Context(W, H) as ctx:
        ctx.upload_volume(vox, dims, voxel_size)
        ctx.set_camera(cam)
        ctx.set_params(vb.default_params(kernel=kernel, **vkw))
        if partition:
            ctx.set_partition(*partition)
        img, st = ctx.render()
        return img, st


@pytest.mark.parametrize("cid", [c[0] for c in scenarios.CASES])
@pytest.mark.parametrize("kernel", [vb.KERNEL_DIRECT, vb.KERNEL_AUTO, vb.KERNEL_FAST, vb.KERNEL_WINDOWED, vb.KERNEL_TEXGATHER,
                                    vb.KERNEL_TEXPAIR, vb.KERNEL_TEXPAIR2, vb.KERNEL_TEXPAIR_PIPE, vb.KERNEL_HYBRID, vb.KERNEL_ZLSU,
                                    vb.KERNEL_NEAREST_TEX],
                         ids=["direct", "auto", "fast", "windowed", "texgather", "texpair", "texpair2", "texpair_pipe", "hybrid", "zlsu",
                              "nearest_tex"])
def test_case_matches_oracle(cid, kernel):
    _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    okw, vkw = scenarios.split_kwargs(kw)
    ref, cnt = oracle_frame(cam, vox, dims, bpv, W, H, voxel_size=vs, **okw)
    img, st = run_product(vox, dims, vs, cam, W, H, vkw, kernel=kernel)
    assert st.kernel_launches >= 1 and st.kernel_ms > 0
    compare(img, ref, cid, exact=cid not in scenarios.TOLERANCE_ONLY)
    if cid not in ("window_min_eq_max_nan",):
        assert cnt["rays_hit"] > 0 and np.isfinite(img).all()


@pytest.mark.parametrize("cid", ["c1_trilinear_128steps", "ragged_nearest", "u16_aniso_trilinear", "mip_nearest"])
def test_golden_fixture(cid, golden_dir):
    """Committed fixtures (tests/golden/make_golden.py): inputs are regenerated from the seed,
    the expected image comes from the file, not from running the oracle now."""
    import os
    g = np.load(os.path.join(golden_dir, f"{cid}.npz"))
    _, vname, cname, (W, H), kw = scenarios.case_by_id(cid)
    vox, dims, bpv, vs = scenarios.volume(vname)
    assert int(g["voxel_crc"]) == int(np.bitwise_xor.reduce(vox.astype(np.uint32) * np.arange(1, vox.size + 1, dtype=np.uint32)))
    cam = g["cam"]
    _, vkw = scenarios.split_kwargs(kw)
    img, _ = run_product(vox, dims, vs, cam, W, H, vkw)
    compare(img, g["rgba"], "golden " + cid)


@pytest.mark.parametrize("vname,cname,kw", [
    ("mix64_u8", "K0", dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1)),
    ("mix64_u8", "K2", dict(alpha_scale=0.3, min_val=20, max_val=220, filter=1)),
    ("mix_64x64x32_u16", "K1", dict(alpha_scale=0.05, min_val=0, max_val=4095, filter=1)),
    ("rand_40x56x33_u16", "K2", dict(alpha_scale=0.03, min_val=100, max_val=4000, filter=1)),
])
def test_optimised_kernels_really_run_and_match(vname, cname, kw):
    """The FAST and WINDOWED kernels are the ones selected (no silent fallback to DIRECT) for
    frames they cover, and both reproduce the oracle bit for bit."""
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    W, H = 320, 200
    ref, _ = oracle_frame(cam, vox, dims, bpv, W, H, voxel_size=vs, **kw)
    for kernel in (vb.KERNEL_FAST, vb.KERNEL_WINDOWED, vb.KERNEL_TEXGATHER, vb.KERNEL_TEXPAIR, vb.KERNEL_TEXPAIR2, vb.KERNEL_TEXPAIR_PIPE, vb.KERNEL_HYBRID, vb.KERNEL_ZLSU, vb.KERNEL_AUTO):
        img, st = run_product(vox, dims, vs, cam, W, H, kw, kernel=kernel)
        if kernel != vb.KERNEL_AUTO:
            assert st.kernel_used == kernel, f"requested kernel {kernel}, ran {st.kernel_used}"
        else:
            assert st.kernel_used in (vb.KERNEL_TEXPAIR, vb.KERNEL_TEXPAIR2, vb.KERNEL_TEXPAIR_PIPE, vb.KERNEL_HYBRID)
        compare(img, ref, f"{vname}/{cname}/kernel{kernel}")


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


def test_volume_stats_match_reference_loops():
    """min/max + histogram kernels vs a numpy restatement of RendererCore.cpp:360-405."""
    rng = np.random.default_rng(3)
    # (51, 41, 31): odd voxel count -> the scalar tail behind the 16-byte vector body; 65535: full u16 range
    for bpv, hi, dims in ((1, 256, (50, 40, 30)), (2, 3000, (50, 40, 30)), (1, 256, (51, 41, 31)), (2, 65536, (51, 41, 31))):
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
            ctx.set_partition(2, 2, 8)


def test_full_size_1024_cube_sampled_rows():
    """BASELINE's headline size (1024^3 uint16, 1920x1080, 1024 steps): the oracle renders
    every 60th row of the same volume; those rows must match bit for bit.  Also checks the
    size-independent property N-GPU == 1-GPU on the full frame."""
    import torch
    from volren_b200 import workloads
    dims, W, H = (1024, 1024, 1024), 1920, 1080
    cam = scenarios.camera("K2")
    kw = dict(alpha_scale=0.02, min_val=0, max_val=4095, filter=1)
    with vb.Context(W, H) as ctx:
        copy = torch.empty(1024 ** 3, dtype=torch.int16, device="cuda")
        ctx.upload_synthetic(dims, 2, 4095, workloads.SEEDS["C4"], True, copy_out_dptr=copy.data_ptr())
        host = copy.cpu().numpy().view(np.uint16)
        del copy
        ctx.set_camera(cam)
        ctx.set_params(vb.default_params(**kw))
        img, st = ctx.render()
        for kernel in (vb.KERNEL_DIRECT, vb.KERNEL_WINDOWED, vb.KERNEL_FAST, vb.KERNEL_TEXGATHER, vb.KERNEL_TEXPAIR, vb.KERNEL_TEXPAIR2, vb.KERNEL_TEXPAIR_PIPE, vb.KERNEL_HYBRID, vb.KERNEL_ZLSU):      # every kernel, same bits, at full size
            ctx.set_params(vb.default_params(kernel=kernel, **kw))
            other, st2 = ctx.render()
            assert st2.kernel_used == kernel
            assert np.array_equal(other.view(np.uint32), img.view(np.uint32))
        ctx.set_params(vb.default_params(**kw))
        # 2-way partition of the same frame
        acc = np.zeros_like(img)
        for rank in range(2):
            ctx.set_partition(rank, 2, 8)
            part, _ = ctx.render()
            owned = ((np.arange(H) // 8) % 2) == rank
            acc[owned] = part[owned]
    assert np.array_equal(acc.view(np.uint32), img.view(np.uint32))
    p = orc.make_params(W, H, dims, 2, cam, row_begin=7, row_stride=60, **kw)
    ref, cnt, _ = orc.render(p, host, nthreads=16)
    rows = np.arange(7, H, 60)
    compare(img[rows], ref[rows], "1024^3 sampled rows")
    assert cnt["rays_hit"] > 0.8 * rows.size * W   # K2: the volume covers ~85 % of the frame


@pytest.mark.parametrize("W,H", [(131, 67), (1, 5), (65, 8)])
def test_texpair_odd_image_sizes(W, H):
    """Two rays per thread: odd widths (the last thread of a row owns one pixel), widths below one
    CTA tile, and pairs whose rays leave the box at different steps (oblique camera)."""
    vox, dims, bpv, vs = scenarios.volume("mix_64x64x32_u16")
    cam = scenarios.camera("K1")
    kw = dict(alpha_scale=0.05, min_val=0, max_val=4095, filter=1)
    ref, _ = oracle_frame(cam, vox, dims, bpv, W, H, voxel_size=vs, **kw)
    for kernel in (vb.KERNEL_TEXPAIR, vb.KERNEL_TEXPAIR2, vb.KERNEL_TEXPAIR_PIPE, vb.KERNEL_HYBRID, vb.KERNEL_ZLSU):
        img, st = run_product(vox, dims, vs, cam, W, H, kw, kernel=kernel)
        assert st.kernel_used == kernel
        compare(img, ref, f"texpair {W}x{H} kernel{kernel}")


@pytest.mark.parametrize("vname,cname,kw", [
    ("mix64_u8", "K1", dict(alpha_scale=0.3, min_val=0, max_val=255, filter=1, tf=True)),
    ("mix64_u8", "K2", dict(alpha_scale=0.08, min_val=30, max_val=200, filter=1, tf=True, step_scale=0.5)),
    ("mix_64x64x32_u16", "K1", dict(alpha_scale=0.2, min_val=100, max_val=3900, filter=1, tf=True)),
    ("rand_40x56x33_u16", "K2", dict(alpha_scale=0.05, min_val=0, max_val=4095, filter=1, tf=True)),
])
def test_transfer_function_runs_on_the_pipelined_gather_kernel(vname, cname, kw):
    """CubicSpline transfer function (BASELINE config 3): the opacity LUT is read inside the optimised
    kernel -- no fallback to the generic DIRECT loop -- and the frame equals the oracle's bit for bit."""
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    W, H = 320, 200
    okw, vkw = scenarios.split_kwargs(kw)
    ref, _ = oracle_frame(cam, vox, dims, bpv, W, H, voxel_size=vs, **okw)
    img, st = run_product(vox, dims, vs, cam, W, H, vkw)
    assert st.kernel_used == vb.KERNEL_TEXPAIR_PIPE
    compare(img, ref, f"tf {vname}/{cname}")
    direct, st2 = run_product(vox, dims, vs, cam, W, H, vkw, kernel=vb.KERNEL_DIRECT)
    assert st2.kernel_used == vb.KERNEL_DIRECT
    assert np.array_equal(direct.view(np.uint32), img.view(np.uint32))


@pytest.mark.parametrize("vname,cname,kw", [
    ("mix_64x64x32_u16", "K2", dict(alpha_scale=1.0, min_val=1000, max_val=3000, filter=1, is_mip=1)),
    ("mix64_u8", "K1", dict(alpha_scale=0.6, min_val=0, max_val=255, filter=1, is_mip=1)),
    ("rand_40x56x33_u16", "K0", dict(alpha_scale=0.9, min_val=0, max_val=4095, filter=1, is_mip=1, step_scale=0.5)),
])
def test_mip_runs_on_the_pipelined_gather_kernel(vname, cname, kw):
    """MIP (VolumeRenderer.cs:141-173, the GUI's use_mip toggle) with the trilinear filter runs in the
    optimised kernel and equals the oracle and the generic DIRECT loop bit for bit."""
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    W, H = 320, 200
    okw, vkw = scenarios.split_kwargs(kw)
    ref, _ = oracle_frame(cam, vox, dims, bpv, W, H, voxel_size=vs, **okw)
    img, st = run_product(vox, dims, vs, cam, W, H, vkw)
    assert st.kernel_used == vb.KERNEL_TEXPAIR_PIPE
    compare(img, ref, f"mip {vname}/{cname}")
    direct, st2 = run_product(vox, dims, vs, cam, W, H, vkw, kernel=vb.KERNEL_DIRECT)
    assert st2.kernel_used == vb.KERNEL_DIRECT
    assert np.array_equal(direct.view(np.uint32), img.view(np.uint32))


@pytest.mark.parametrize("world,tile_rows", [(2, 16), (3, 4), (8, 16)])
def test_owned_tiles_to_host_frame_assemble_the_frame(world, tile_rows):
    """Multi-GPU end to end (vr_render_owned_to_host): every rank copies only its own row tiles into a
    full host frame; all ranks together (emulated on one device) reproduce the unpartitioned frame."""
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K1")
    W, H = 200, 150
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


@pytest.mark.parametrize("vname,cname,kw", [
    ("rand_48x40x36_u8", "K1", dict(alpha_scale=0.08, min_val=0, max_val=255, filter=1, view_top=1)),
    ("rand_40x56x33_u16", "K0", dict(alpha_scale=0.05, min_val=0, max_val=4095, filter=1, view_bottom=1)),
    ("mix_64x64x32_u16", "K2", dict(alpha_scale=0.05, min_val=1000, max_val=3000, filter=1, view_top=1)),
    ("mix64_u8", "K1", dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1, view_bottom=1, step_scale=0.5)),
    ("mix64_u8", "K2", dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1, view_top=1, view_bottom=1)),   # top wins (:183)
])
def test_view_swizzles_run_on_the_pipelined_gather_kernel(vname, cname, kw):
    """rotate_to_top / rotate_to_bottom (VolumeRenderer.cs:68-78,183-190) with the trilinear filter run
    in the optimised kernel and equal the oracle and the generic DIRECT loop bit for bit."""
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    W, H = 320, 200
    okw, vkw = scenarios.split_kwargs(kw)
    ref, _ = oracle_frame(cam, vox, dims, bpv, W, H, voxel_size=vs, **okw)
    img, st = run_product(vox, dims, vs, cam, W, H, vkw)
    assert st.kernel_used == vb.KERNEL_TEXPAIR_PIPE
    compare(img, ref, f"view {vname}/{cname}")
    direct, st2 = run_product(vox, dims, vs, cam, W, H, vkw, kernel=vb.KERNEL_DIRECT)
    assert st2.kernel_used == vb.KERNEL_DIRECT
    assert np.array_equal(direct.view(np.uint32), img.view(np.uint32))


@pytest.mark.parametrize("vname,cname,kw", [
    ("mix64_u8", "K0", dict(alpha_scale=0.05, min_val=0, max_val=255, filter=0)),
    ("mix64_u8", "K2", dict(alpha_scale=1.0, min_val=40, max_val=200, filter=0)),
    ("rand_48x40x36_u8", "K1", dict(alpha_scale=0.08, min_val=10, max_val=250, filter=0)),
    ("mix_64x64x32_u16", "K0", dict(alpha_scale=0.5, min_val=1000, max_val=3000, filter=0)),
    ("rand_40x56x33_u16", "orbit", dict(alpha_scale=0.03, min_val=0, max_val=4095, filter=0, step_scale=0.5)),
    ("full_u16", "K1", dict(alpha_scale=0.04, min_val=0, max_val=65535, filter=0)),
    ("one_voxel_u8", "K0", dict(alpha_scale=0.7, min_val=0, max_val=255, filter=0)),
])
def test_nearest_filter_runs_on_the_texel_load_kernel(vname, cname, kw):
    """The de-facto reference filter (integer texture => nearest texel): AUTO runs the pipelined
    texel-load kernel, bit-identical to the oracle and to the FAST and DIRECT kernels."""
    vox, dims, bpv, vs = scenarios.volume(vname)
    cam = scenarios.camera(cname)
    W, H = 320, 200
    ref, _ = oracle_frame(cam, vox, dims, bpv, W, H, voxel_size=vs, **kw)
    img, st = run_product(vox, dims, vs, cam, W, H, kw)
    assert st.kernel_used == vb.KERNEL_NEAREST_TEX
    compare(img, ref, f"nearest {vname}/{cname}")
    for kernel in (vb.KERNEL_FAST, vb.KERNEL_DIRECT):
        other, st2 = run_product(vox, dims, vs, cam, W, H, kw, kernel=kernel)
        assert st2.kernel_used == kernel
        assert np.array_equal(other.view(np.uint32), img.view(np.uint32))
