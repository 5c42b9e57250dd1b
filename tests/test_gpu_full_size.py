"""-m gpu: BASELINE.json's configurations at their FULL sizes.  The oracle renders every k-th row of the same
volume (seconds of CPU work); those rows must match bit for bit.  Size-independent properties ride along: every
kernel produces the same bits, N-GPU == 1-GPU, skipping on == skipping off."""
import numpy as np
import pytest

import volren_b200 as vb
from oracle import orc
from volren_b200 import host, workloads

import scenarios
from util import compare

pytestmark = pytest.mark.gpu


def device_volume(ctx, dims, bpv, vmax, seed, voxel_size=(1.0, 1.0, 1.0)):
    """Generate `mix` in HBM, upload it, and hand back the host copy the oracle needs."""
    import torch
    n = int(np.prod(dims))
    copy = torch.empty(n, dtype=torch.int16 if bpv == 2 else torch.uint8, device="cuda")
    ctx.upload_synthetic(dims, bpv, vmax, seed, True, voxel_size=voxel_size, copy_out_dptr=copy.data_ptr())
    host_vol = copy.cpu().numpy()
    del copy
    torch.cuda.empty_cache()
    return host_vol.view(np.uint16) if bpv == 2 else host_vol


def check_rows(img, host_vol, dims, bpv, cam, W, H, row0, stride, what, voxel_size=(1.0, 1.0, 1.0), nthreads=16, **okw):
    p = orc.make_params(W, H, dims, bpv, cam, voxel_size=voxel_size, row_begin=row0, row_stride=stride, **okw)
    ref, cnt, _ = orc.render(p, host_vol, nthreads=nthreads)
    rows = np.arange(row0, H, stride)
    compare(img[rows], ref[rows], what)
    return cnt


def test_c4_1024_cube_1080p_sampled_rows():
    """BASELINE's headline (C4: 1024^3 uint16, 1920x1080, 1024 steps, trilinear + nearest): every 60th row vs the
    oracle; every kernel the same bits; skipping on == off; 2-way partition == full frame."""
    dims, W, H = (1024, 1024, 1024), 1920, 1080
    cam = scenarios.camera("K2")
    with vb.Context(W, H) as ctx:
        host_vol = device_volume(ctx, dims, 2, 4095, workloads.SEEDS["C4"])
        ctx.set_camera(cam)
        for filt in (1, 0):
            kw = dict(alpha_scale=0.02, min_val=0, max_val=4095, filter=filt)
            ctx.set_params(vb.default_params(**kw))
            img, st = ctx.render()
            assert st.kernel_used == (vb.KERNEL_TEXPAIR_PIPE if filt else vb.KERNEL_NEAREST_TEX)
            for kernel, skip in ((vb.KERNEL_DIRECT, vb.SKIP_OFF), (vb.KERNEL_AUTO, vb.SKIP_ON)):
                ctx.set_params(vb.default_params(kernel=kernel, empty_skip=skip, **kw))
                other, st2 = ctx.render()
                assert np.array_equal(other.view(np.uint32), img.view(np.uint32)), (filt, kernel, skip)
            cnt = check_rows(img, host_vol, dims, 2, cam, W, H, 7, 60, f"C4 filter {filt}", **kw)
            assert cnt["rays_hit"] > 0.8 * len(range(7, H, 60)) * W      # K2: the volume covers ~85 % of the frame
        # 2-way partition of the trilinear frame
        kw = dict(alpha_scale=0.02, min_val=0, max_val=4095, filter=1)
        ctx.set_params(vb.default_params(**kw))
        full, _ = ctx.render()
        acc = np.zeros_like(full)
        for rank in range(2):
            ctx.set_partition(rank, 2, 8)
            part, _ = ctx.render()
            owned = ((np.arange(H) // 8) % 2) == rank
            acc[owned] = part[owned]
    assert np.array_equal(acc.view(np.uint32), full.view(np.uint32))


def test_c2_256_cube_u8_1024sq_512_steps():
    """BASELINE config 2: 256^3 uint8, 1024x1024, 512 steps (2 samples per voxel), windows [0,255] and [30,180]."""
    dims, W, H = (256, 256, 256), 1024, 1024
    with vb.Context(W, H) as ctx:
        host_vol = device_volume(ctx, dims, 1, 255, workloads.SEEDS["C2"])
        for cname, lo, hi, alpha, filt in (("K2", 0, 255, 0.05, 1), ("K0", 30, 180, 1.0, 1), ("K1", 30, 180, 0.05, 0)):
            cam = scenarios.camera(cname)
            kw = dict(alpha_scale=alpha, min_val=lo, max_val=hi, filter=filt, step_scale=0.5)
            ctx.set_camera(cam)
            ctx.set_params(vb.default_params(**kw))
            img, st = ctx.render()
            assert st.kernel_used != vb.KERNEL_DIRECT
            check_rows(img, host_vol, dims, 1, cam, W, H, 3, 32, f"C2 {cname} [{lo},{hi}]", **kw)
            ctx.set_params(vb.default_params(empty_skip=vb.SKIP_ON, **kw))
            other, _ = ctx.render()
            assert np.array_equal(other.view(np.uint32), img.view(np.uint32))


@pytest.mark.skipif(orc.ref_lib() is None, reason="oracle/_ref/libddsbase_ref.so not built")
def test_c3_512_cube_u16_from_a_pvm_written_by_the_reference_encoder(tmp_path):
    """BASELINE config 3: 512^3 uint16 CT-like volume stored as a DDS-compressed .pvm by the REFERENCE's encoder
    (ddsbase.cpp writePVMvolume), read back by the product loader, window [1000,3000], 1024 steps, CubicSpline
    transfer function on and off."""
    dims, W, H = (512, 512, 512), 1920, 1080
    vol = vb.synthetic_to_host(dims, 2, 4095, workloads.SEEDS["C3"])
    path = str(tmp_path / "c3.pvm")
    orc.ref_write_pvm(path, vol, dims, 2, (1.0, 1.0, 1.0))
    dec = host.pvm_decode(path=path)
    assert dec["ok"] and dec["dims"] == dims and dec["components"] == 2
    # the reference hands the 16-bit payload to GL untouched (RendererCore.cpp:347,419): host byte order as stored
    loaded = np.frombuffer(dec["payload"], dtype=np.uint16, count=int(np.prod(dims)))
    assert np.array_equal(loaded, vol)
    cam = scenarios.camera("K2")
    lut = scenarios.tf_lut()
    with vb.Context(W, H) as ctx:
        ctx.upload_volume(loaded, dims, dec["scale"])
        ctx.set_camera(cam)
        for tf in (None, lut):
            for alpha in (0.05, 1.0):
                kw = dict(alpha_scale=alpha, min_val=1000, max_val=3000, filter=1, step_scale=0.5)
                ctx.set_params(vb.default_params(tf_lut=tf, **kw))
                img, st = ctx.render()
                assert st.kernel_used == vb.KERNEL_TEXPAIR_PIPE and st.skip_used == 1      # most of C3 is below the window
                check_rows(img, loaded, dims, 2, cam, W, H, 11, 90, f"C3 tf={'on' if tf is not None else 'off'} alpha {alpha}", tf_lut=tf, **kw)
                ctx.set_params(vb.default_params(tf_lut=tf, empty_skip=vb.SKIP_OFF, **kw))
                other, st2 = ctx.render()
                assert st2.skip_used == 0 and np.array_equal(other.view(np.uint32), img.view(np.uint32))


def test_non_power_of_two_anisotropic_ct_shape():
    """512x512x300 uint16, spacing 0.7/0.7/1.5 mm: no tex-coord divisor is a power of two (Markstein division,
    verified on the device), the window needs the clamp; the variants real CT data selects."""
    dims, W, H = (512, 512, 300), 1920, 1080
    vs = (0.7, 0.7, 1.5)
    with vb.Context(W, H) as ctx:
        host_vol = device_volume(ctx, dims, 2, 4095, 0x5EED0011, voxel_size=vs)
        for cname, filt in (("K2", 1), ("K1", 0)):
            cam = scenarios.camera(cname)
            kw = dict(alpha_scale=0.03, min_val=200, max_val=3500, filter=filt)
            ctx.set_camera(cam)
            ctx.set_params(vb.default_params(**kw))
            img, st = ctx.render()
            assert st.kernel_used != vb.KERNEL_DIRECT
            check_rows(img, host_vol, dims, 2, cam, W, H, 5, 72, f"CT shape {cname}", voxel_size=vs, **kw)


def test_c5_8gib_volume_4k_frame_one_gpu():
    """BASELINE config 5 on ONE GPU: 2048x2048x1024 uint16 (8 GiB), 3840x2160, 2048 steps; every 240th row."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * (1 << 30):
        pytest.skip("needs ~40 GiB of HBM")
    dims, W, H = (2048, 2048, 1024), 3840, 2160
    cam = scenarios.camera("K2")
    kw = dict(alpha_scale=0.02, min_val=0, max_val=4095, filter=1)
    with vb.Context(W, H) as ctx:
        host_vol = device_volume(ctx, dims, 2, 4095, workloads.SEEDS["C5"])
        ctx.set_camera(cam)
        ctx.set_params(vb.default_params(**kw))
        img, st = ctx.render()
        assert st.kernel_used == vb.KERNEL_TEXPAIR_PIPE
        m = ctx.memory_info()
        assert m["linear_bytes"] + m["array_bytes"] + m["zpair_array_bytes"] < 26 * (1 << 30)      # 8 GiB linear + 16 GiB z-pair
        check_rows(img, host_vol, dims, 2, cam, W, H, 100, 240, "C5", **kw)
