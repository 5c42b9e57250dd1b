"""-m gpu: the fused multi-GPU hand-off -- rank 1 maps rank 0's frame through CUDA IPC and its
march kernel stores its row tiles straight into it.  Two real processes; both use cuda:0 when
the box has a single GPU (the IPC mapping path is the same), cuda:rank otherwise."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import volren_b200 as vb
    import scenarios
    dev = rank if torch.cuda.device_count() >= world else 0
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K1")
    W, H, tile_rows = 200, 150, 16
    kw = dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1)
    with vb.Context(W, H, device=dev) as ctx:
        ctx.upload_volume(vox, dims, vs)
        ctx.set_camera(cam)
        ctx.set_params(vb.default_params(**kw))
        full = None
        if rank == 0:
            full, _ = ctx.render()                                   # unpartitioned frame, for comparison
            # overwrite rank 0's frame with a different image so rows nobody stores would be noticed
            ctx.set_params(vb.default_params(alpha_scale=0.9, min_val=50, max_val=200, filter=0))
            poison, _ = ctx.render()
            assert not np.array_equal(poison, full)
            ctx.set_params(vb.default_params(**kw))
        handle = [ctx.frame_export_ipc() if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        ptr = ctx.frame_device_ptr() if rank == 0 else ctx.frame_open_ipc(handle[0])
        ctx.set_partition(rank, world, tile_rows)
        dist.barrier()
        st = ctx.render_device(ptr, compact=False)
        assert st.kernel_launches >= 1
        dist.barrier()                                               # every rank's stores have landed
        if rank == 0:
            got = ctx.read_frame()
            np.save(out_path, np.array([1 if np.array_equal(got.view(np.uint32), full.view(np.uint32)) else 0]))
        dist.barrier()
        if rank != 0:
            ctx.frame_close_ipc(ptr)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_stores_assemble_the_frame_in_rank0(tmp_path, world):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert np.load(out)[0] == 1


def _worker_device_barrier(rank, world, port, out_path, fused=False):
    """Several frames back to back with NO host-side collective between them: completion and reuse of
    rank 0's frame are ordered only by the barrier words in peer memory (vr_peer_frame_arrive/release)."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python"), os.path.join(ROOT, "tests")]
    # the rank processes of this test may time-share one GPU (a spinning wait kernel then holds its whole
    # time slice while the kernel it waits for sits in another process): give the spins a generous bound
    os.environ["VR_PEER_TIMEOUT_MS"] = "60000"
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import volren_b200 as vb
    import scenarios
    dev = rank if torch.cuda.device_count() >= world else 0
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K1")
    W, H, tile_rows, frames = 200, 150, 16, 6
    alphas = [0.05, 0.4, 0.11, 0.9, 0.02, 0.3]
    ok = 1
    with vb.Context(W, H, device=dev) as ctx:
        ctx.upload_volume(vox, dims, vs)
        ctx.set_camera(cam)
        refs = []
        if rank == 0:
            for a in alphas:
                ctx.set_params(vb.default_params(alpha_scale=a, min_val=0, max_val=255, filter=1))
                refs.append(ctx.render()[0])
        handle = [ctx.frame_export_ipc() if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        ptr = ctx.frame_device_ptr() if rank == 0 else ctx.frame_open_ipc(handle[0])
        ctx.set_partition(rank, world, tile_rows)
        dist.barrier()                                               # setup only; none inside the frame loop
        for f in range(1, frames + 1):
            ctx.set_params(vb.default_params(alpha_scale=alphas[f - 1], min_val=0, max_val=255, filter=1))
            if rank != 0:
                ctx.peer_frame_release(ptr, f - 1, is_owner=False)   # rank 0 is done with the previous frame
            if fused:
                # vr_render_peer: the march kernel's last CTA publishes the arrival, no signal kernel
                st = ctx.render_peer(ptr, f, world, is_owner=(rank == 0))
                if st.kernel_launches != (2 if rank == 0 else 1):    # march (+ the owner's wait); never a signal kernel
                    ok = 0
                    print(f"rank {rank}: {st.kernel_launches} launches in the fused hand-off", flush=True)
            else:
                ctx.render_device(ptr, compact=False)
                ctx.peer_frame_arrive(ptr, f, world, is_owner=(rank == 0))
            if rank == 0:
                got = ctx.read_frame()                               # same stream: after the arrival wait
                if not np.array_equal(got.view(np.uint32), refs[f - 1].view(np.uint32)):
                    ok = 0
                    print(f"frame {f}: {int((got.view(np.uint32) != refs[f - 1].view(np.uint32)).any(axis=(1, 2)).sum())} rows differ, "
                          f"status {ctx.peer_frame_status(ptr)}", flush=True)
                ctx.peer_frame_release(ptr, f, is_owner=True)
        if rank == 0:
            torch.cuda.synchronize(dev)      # the last release is a kernel in the context's (non-blocking) stream; the status read is not ordered after it
            st = ctx.peer_frame_status(ptr)
            if st["timed_out"] or st["arrivals"] != frames * world or st["released"] != frames:
                ok = 0
                print(f"barrier status {st}, expected {frames * world} arrivals, {frames} released", flush=True)
            np.save(out_path, np.array([ok]))
        dist.barrier()
        if rank != 0:
            ctx.frame_close_ipc(ptr)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("fused", [False, True], ids=["signal_kernel", "fused_epilogue"])
def test_peer_memory_frame_barrier(tmp_path, world, fused):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker_device_barrier, args=(world, _free_port(), out, fused), nprocs=world, join=True)
    assert np.load(out)[0] == 1


def _worker_two_target_frames(rank, world, port, out_path):
    """bench.py's peer hand-off: TWO target frames on rank 0, asynchronous marches (vr_render_peer with stats == NULL),
    the owner's arrival wait + release on a consumer stream.  No host synchronisation and no collective inside the
    frame loop on the producing ranks; rank 0 checks every frame bit for bit."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python"), os.path.join(ROOT, "tests")]
    os.environ["VR_PEER_TIMEOUT_MS"] = "60000"
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import volren_b200 as vb
    import scenarios
    dev = rank if torch.cuda.device_count() >= world else 0
    torch.cuda.set_device(dev)
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K1")
    W, H, tile_rows = 200, 150, 8
    alphas = [0.05, 0.4, 0.11, 0.9, 0.02, 0.3, 0.07]
    ok = 1
    with vb.Context(W, H, device=dev) as ctx:
        ctx.upload_volume(vox, dims, vs)
        ctx.set_camera(cam)
        refs = []
        if rank == 0:
            for a in alphas:
                ctx.set_params(vb.default_params(alpha_scale=a, min_val=0, max_val=255, filter=1))
                refs.append(ctx.render()[0])
        ctx_b = vb.Context(W, H, device=dev) if rank == 0 else None      # owns the second target frame, nothing else
        handles = [[ctx.frame_export_ipc(), ctx_b.frame_export_ipc()] if rank == 0 else None]
        dist.broadcast_object_list(handles, src=0)
        ptrs = [ctx.frame_device_ptr(), ctx_b.frame_device_ptr()] if rank == 0 else [ctx.frame_open_ipc(h) for h in handles[0]]
        cstream = torch.cuda.Stream() if rank == 0 else None
        ctx.set_partition(rank, world, tile_rows)
        dist.barrier()                                               # setup only
        uses = [0, 0]
        for f in range(1, len(alphas) + 1):
            b = f % 2
            uses[b] += 1
            u = uses[b]
            ctx.set_params(vb.default_params(alpha_scale=alphas[f - 1], min_val=0, max_val=255, filter=1))
            ctx.peer_frame_release(ptrs[b], u - 1, is_owner=False)   # device-side wait: the frame's previous occupant was consumed
            assert ctx.render_peer(ptrs[b], f, world, is_owner=False, wait=False) is None
            if rank == 0:
                ctx.peer_frame_wait_arrivals(ptrs[b], u, world, stream=cstream.cuda_stream)
                cstream.synchronize()                                # consumer: the frame is complete
                got = (ctx if b == 0 else ctx_b).read_frame()
                if not np.array_equal(got.view(np.uint32), refs[f - 1].view(np.uint32)):
                    ok = 0
                    print(f"frame {f}: {int((got.view(np.uint32) != refs[f - 1].view(np.uint32)).any(axis=(1, 2)).sum())} rows differ", flush=True)
                ctx.peer_frame_release(ptrs[b], u, is_owner=True, stream=cstream.cuda_stream)
        ms = [ctx.peer_kernel_ms(f) for f in range(1, len(alphas) + 1)]
        if not all(m > 0.0 for m in ms):
            ok = 0
            print(f"rank {rank}: kernel times {ms}", flush=True)
        torch.cuda.synchronize(dev)
        if rank == 0:
            for b in (0, 1):
                st = ctx.peer_frame_status(ptrs[b])
                if st["timed_out"] or st["arrivals"] != uses[b] * world or st["released"] != uses[b]:
                    ok = 0
                    print(f"target frame {b}: status {st}, expected {uses[b] * world} arrivals, {uses[b]} released", flush=True)
        okt = torch.tensor([ok])
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if rank == 0:
            np.save(out_path, np.array([int(okt.item())]))
        dist.barrier()
        if rank != 0:
            for p in ptrs:
                ctx.frame_close_ipc(p)
        if ctx_b is not None:
            ctx_b.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_target_frames_asynchronous_hand_off(tmp_path, world):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker_two_target_frames, args=(world, _free_port(), out), nprocs=world, join=True)
    assert np.load(out)[0] == 1


def test_a_barrier_timeout_is_reported_not_swallowed(monkeypatch):
    """a wait that gives up must surface.  One process plays the owner of a 2-rank frame whose peer never
    arrives: the bounded wait times out, the NEXT hand-off call fails with VR_ERR_TIMEOUT until the caller resets."""
    sys.path[:0] = [os.path.join(ROOT, "volume-renderer_b200", "python"), os.path.join(ROOT, "tests")]
    monkeypatch.setenv("VR_PEER_TIMEOUT_MS", "150")
    import volren_b200 as vb
    import scenarios
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    with vb.Context(96, 64) as ctx:
        ctx.upload_volume(vox, dims, vs)
        ctx.set_camera(scenarios.camera("K0"))
        ctx.set_params(vb.default_params(alpha_scale=0.1, min_val=0, max_val=255, filter=1))
        ctx.set_partition(0, 2, 8)
        ptr = ctx.frame_device_ptr()
        ctx.render_peer(ptr, 1, 2, is_owner=True)            # arrivals stay at 1 of 2: the wait gives up after 150 ms
        torch.cuda.synchronize()
        assert ctx.peer_frame_status(ptr)["timed_out"] == 1
        for call in (lambda: ctx.render_peer(ptr, 2, 2, is_owner=True), lambda: ctx.peer_frame_arrive(ptr, 2, 2, True),
                     lambda: ctx.peer_frame_release(ptr, 1, True)):
            with pytest.raises(vb.VolrenError) as e:
                call()
            assert e.value.code == -7 and "timed out" in str(e.value)
        ctx.peer_frame_reset(ptr)
        assert ctx.peer_frame_status(ptr)["timed_out"] == 0
        ctx.peer_frame_release(ptr, 1, True)                 # works again
