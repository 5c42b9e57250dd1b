"""CPU, world_size 2 and 3 over gloo: the multi-GPU host logic -- interleaved
row-tile ownership, the compact per-rank tile buffers, the single gather to rank 0 and the
de-interleave.  The per-rank "render" here is the CPU oracle restricted to the rank's rows
(the product march has no CPU path); what is under test is the partition/gather/assemble
plumbing of volren_b200.dist, which the GPU path shares."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tile_rows, H, W, out_path):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from volren_b200 import dist as vdist
    from oracle import orc
    import scenarios
    r, w, _ = vdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    vox, dims, bpv, vs = scenarios.volume("mix64_u8")
    cam = scenarios.camera("K1")
    kw = dict(alpha_scale=0.05, min_val=0, max_val=255, filter=1)
    rows = vdist.compact_rows(H, world, tile_rows)
    local = torch.full((rows, W, 4), float("nan"), dtype=torch.float32)
    # this rank's rows, one oracle call per owned tile
    tiles = (H + tile_rows - 1) // tile_rows
    for t in range(rank, tiles, world):
        y0, y1 = t * tile_rows, min((t + 1) * tile_rows, H)
        p = orc.make_params(W, H, dims, bpv, cam, row_begin=y0, row_end=y1, row_stride=1, **kw)
        img, _, _ = orc.render(p, vox, nthreads=1)
        for y in range(y0, y1):
            assert vdist.owner_of_row(y, world, tile_rows) == rank
            local[vdist.local_row_of(y, world, tile_rows)] = torch.from_numpy(img[y])
    gathered = torch.empty((world, rows, W, 4), dtype=torch.float32) if rank == 0 else None
    vdist.gather_tiles(local, gathered, dst=0)
    if rank == 0:
        frame = vdist.assemble_reference(gathered, H, tile_rows)
        full, _, _ = orc.render(orc.make_params(W, H, dims, bpv, cam, **kw), vox, nthreads=2)
        ok = np.array_equal(frame.numpy().view(np.uint32), full.view(np.uint32))
        np.save(out_path, np.array([1 if ok else 0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,tile_rows,H", [(2, 8, 70), (2, 1, 33), (3, 4, 50)])
def test_gather_and_assemble_equals_single_process(tmp_path, world, tile_rows, H):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(world, _free_port(), tile_rows, H, 96, out), nprocs=world, join=True)
    assert np.load(out)[0] == 1


def test_partition_maps_are_a_bijection():
    from volren_b200 import dist as vdist
    for world in (1, 2, 3, 4, 8):
        for tile_rows in (1, 4, 8, 16):
            for H in (1, 7, 64, 1080, 2160):
                rows = vdist.compact_rows(H, world, tile_rows)
                seen = set()
                for y in range(H):
                    r, l = vdist.owner_of_row(y, world, tile_rows), vdist.local_row_of(y, world, tile_rows)
                    assert 0 <= r < world and 0 <= l < rows
                    seen.add((r, l))
                assert len(seen) == H
                # balance: no rank owns more than one tile more than another
                counts = [sum(1 for y in range(0, H, 1) if vdist.owner_of_row(y, world, tile_rows) == r) for r in range(world)]
                assert max(counts) - min(counts) <= tile_rows


def _shared_frame_worker(rank, world, name, tile_rows, H, W, frames, out_path, buffers=1):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python")]
    from volren_b200 import dist as vdist
    import time
    if rank != 0:
        for _ in range(200):                       # wait for rank 0 to create the segment
            try:
                sf = vdist.SharedHostFrame(name, W, H, rank, world, create=False, register_cuda=False, buffers=buffers)
                break
            except FileNotFoundError:
                time.sleep(0.02)
    else:
        sf = vdist.SharedHostFrame(name, W, H, rank, world, create=True, register_cuda=False, buffers=buffers)
    ok = 1
    tiles = (H + tile_rows - 1) // tile_rows
    for f in range(1, frames + 1):
        sf.wait_writable(f)                        # never overwrite a frame the consumer still reads
        if rank == 1 and f == 2:
            time.sleep(0.05)                       # a straggler: with two buffers the others run one frame ahead
        for t in range(rank, tiles, world):
            y0, y1 = t * tile_rows, min((t + 1) * tile_rows, H)
            ys = np.arange(y0, y1, dtype=np.float32)[:, None, None]
            sf.buffer_of(f)[y0:y1] = ys * 1000.0 + f      # value encodes (row, frame)
        sf.mark_done(f)
        if rank == 0:
            sf.wait_all_done(f)
            expect = np.arange(H, dtype=np.float32)[:, None, None] * 1000.0 + f
            if not np.array_equal(sf.buffer_of(f), np.broadcast_to(expect, (H, W, 4))):
                ok = 0
            sf.release(f)
    if rank == 0:
        np.save(out_path, np.array([ok]))
        time.sleep(0.2)                            # let the others detach before the segment is unlinked
    sf.close()


@pytest.mark.parametrize("buffers", [1, 2])
@pytest.mark.parametrize("world,tile_rows,H", [(2, 16, 70), (3, 4, 50)])
def test_shared_host_frame_protocol(tmp_path, world, tile_rows, H, buffers):
    """The multi-GPU end-to-end hand-off on the host: every rank writes its row tiles into one shared
    frame; done/released flags order completion and reuse without any collective (CPU only: the
    device->host copies of vr_render_owned_to_host are replaced by numpy stores)."""
    out = str(tmp_path / "ok.npy")
    name = f"volren_test_{os.getpid()}_{world}_{buffers}"
    mp.spawn(_shared_frame_worker, args=(world, name, tile_rows, H, 24, 6, out, buffers), nprocs=world, join=True)
    assert np.load(out)[0] == 1


def _pipelined_frame_worker(rank, world, name, tile_rows, H, W, frames, out_path):
    """bench.py's pipelined end-to-end loop on the host side: frame f is SUBMITTED (here: remembered) before frame f - 1
    is collected (here: written), hand-shaken and released; two shared buffers."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python")]
    from volren_b200 import dist as vdist
    import time
    if rank != 0:
        for _ in range(200):
            try:
                sf = vdist.SharedHostFrame(name, W, H, rank, world, create=False, register_cuda=False, buffers=2)
                break
            except FileNotFoundError:
                time.sleep(0.02)
    else:
        sf = vdist.SharedHostFrame(name, W, H, rank, world, create=True, register_cuda=False, buffers=2)
    ok = [1]
    tiles = (H + tile_rows - 1) // tile_rows
    tickets = {}

    def complete(g):
        if rank == world - 1 and g == 3:
            time.sleep(0.05)                       # a straggler
        buf = tickets.pop(g)
        for t in range(rank, tiles, world):        # "the copies of frame g have landed"
            y0, y1 = t * tile_rows, min((t + 1) * tile_rows, H)
            ys = np.arange(y0, y1, dtype=np.float32)[:, None, None]
            buf[y0:y1] = ys * 1000.0 + g
        sf.mark_done(g)
        if rank == 0:
            sf.wait_all_done(g)
            expect = np.arange(H, dtype=np.float32)[:, None, None] * 1000.0 + g
            if not np.array_equal(sf.buffer_of(g), np.broadcast_to(expect, (H, W, 4))):
                ok[0] = 0
            sf.release(g)

    for f in range(1, frames + 1):
        sf.wait_writable(f)                        # the buffer's previous frame (f - 2) has been consumed
        tickets[f] = sf.buffer_of(f)               # vr_render_submit(buffer of frame f)
        if f - 1 in tickets:
            complete(f - 1)
    complete(frames)
    if rank == 0:
        np.save(out_path, np.array([ok[0]]))
        time.sleep(0.2)
    sf.close()


@pytest.mark.parametrize("world,tile_rows,H", [(2, 8, 70), (4, 8, 50)])
def test_pipelined_shared_host_frames(tmp_path, world, tile_rows, H):
    """Order of operations of bench.py's pipelined multi-GPU end-to-end loop (vr_render_submit / vr_render_wait, two
    frames in flight, two shared host buffers): no deadlock, no frame overwritten before the consumer has released it."""
    out = str(tmp_path / "ok.npy")
    name = f"volren_test_pipe_{os.getpid()}_{world}"
    mp.spawn(_pipelined_frame_worker, args=(world, name, tile_rows, H, 24, 9, out), nprocs=world, join=True)
    assert np.load(out)[0] == 1


def test_numa_interleave_is_best_effort(tmp_path):
    """The shared host frame asks for its pages to be interleaved over the NUMA nodes before first touch; on a
    single-node machine or where mbind is filtered it reports False and the frame works all the same."""
    sys.path[:0] = [os.path.join(ROOT, "volume-renderer_b200", "python")]
    from volren_b200 import dist as vdist
    sf = vdist.SharedHostFrame(f"volren_test_numa_{os.getpid()}", 32, 16, 0, 1, create=True, register_cuda=False, buffers=2)
    try:
        assert sf.numa_interleaved in (True, False)
        sf.frames[:] = 3.0
        assert float(sf.buffer_of(1).sum()) == 3.0 * 32 * 16 * 4
    finally:
        sf.close()

