"""Multi-GPU plumbing: one process per GPU, screen-row tiles interleaved across
ranks (tile t belongs to rank t % world), replicated volume, ONE gather of the finished RGBA
tiles to rank 0 per frame over torch.distributed (NCCL on GPUs, gloo in the CPU tests),
followed by the tile de-interleave.  No other collective touches the data path."""
from __future__ import annotations

import os
import time
from multiprocessing import shared_memory

import numpy as np
import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend: str = None):
    rank, world, local_rank = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def compact_rows(height: int, world: int, tile_rows: int) -> int:
    """Rows of every rank's compact tile buffer (rank 0's tile count, padded to whole tiles)."""
    tiles = (height + tile_rows - 1) // tile_rows
    return ((tiles + world - 1) // world) * tile_rows


def owner_of_row(y: int, world: int, tile_rows: int) -> int:
    return (y // tile_rows) % world


def local_row_of(y: int, world: int, tile_rows: int) -> int:
    tile = y // tile_rows
    return (tile // world) * tile_rows + (y % tile_rows)


def gather_tiles(local_compact: torch.Tensor, gathered: torch.Tensor | None, dst: int = 0):
    """local_compact [rows, W, 4] on every rank -> gathered [world, rows, W, 4] on rank dst."""
    world = dist.get_world_size()
    if dist.get_rank() == dst:
        assert gathered is not None and gathered.shape[0] == world
        dist.gather(local_compact, list(gathered.unbind(0)), dst=dst)
    else:
        dist.gather(local_compact, None, dst=dst)


def open_peer_frame(ctx, dst: int = 0) -> int:
    """Fused hand-off: returns the device pointer every rank renders its rows into -- rank dst's own
    frame on rank dst, the CUDA-IPC mapping of that frame (NVLink peer memory) on the others.
    The 64-byte handle travels as a CUDA tensor over the process group."""
    rank = dist.get_rank()
    h = torch.zeros(64, dtype=torch.uint8, device="cuda")
    if rank == dst:
        h.copy_(torch.frombuffer(bytearray(ctx.frame_export_ipc()), dtype=torch.uint8))
    dist.broadcast(h, src=dst)
    if rank == dst:
        return ctx.frame_device_ptr()
    return ctx.frame_open_ipc(bytes(h.cpu().numpy().tobytes()))


def assemble_reference(gathered: torch.Tensor, height: int, tile_rows: int) -> torch.Tensor:
    """Pure-torch statement of vr_assemble_tiles (used on CPU tensors by the gloo tests and as
    the checker of the CUDA de-interleave kernel)."""
    world, rows, W, _ = gathered.shape
    ys = torch.arange(height)
    tile = ys // tile_rows
    rank = tile % world
    local = (tile // world) * tile_rows + (ys % tile_rows)
    return gathered[rank, local]


def _interleave_over_numa_nodes(addr: int, nbytes: int) -> bool:
    """Best effort, before the pages of a shared host frame are first touched: spread them round-robin over the NUMA
    nodes (mbind, MPOL_INTERLEAVE; the policy of a tmpfs mapping is shared by every process that maps it).  Left alone,
    the whole frame lands on the node of whichever process touches it first and half of the GPUs of a two-socket box
    write their tiles across the socket link, all in one direction.  Returns False when there is
    one node, or the call is not permitted (containers often filter it) -- nothing else changes then."""
    try:
        with open("/sys/devices/system/node/online") as f:
            spec = f.read().strip()
        nodes = []
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            nodes.extend(range(int(lo), int(hi or lo) + 1))
        if len(nodes) < 2:
            return False
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        page = os.sysconf("SC_PAGE_SIZE")
        start = addr & ~(page - 1)
        length = (addr + nbytes) - start
        mask = 0
        for n in nodes:
            mask |= 1 << n
        words = (max(nodes) // 64) + 1
        arr = (ctypes.c_ulong * words)(*[(mask >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(words)])
        SYS_mbind, MPOL_INTERLEAVE = 237, 3                      # x86-64
        if os.uname().machine != "x86_64":
            return False
        rc = libc.syscall(SYS_mbind, ctypes.c_void_p(start), ctypes.c_ulong(length), MPOL_INTERLEAVE, arr,
                          ctypes.c_ulong(words * 64 + 1), 0)
        return rc == 0
    except Exception:
        return False


class SharedHostFrame:
    """`buffers` W x H RGBA32F frames in POSIX shared memory, mapped by every rank process (one process per
    GPU) and page-locked in each with cudaHostRegister, so that every rank copies its own row tiles
    device->host over its own PCIe link (vr_render_owned_to_host).  Two tiny flag arrays in the same
    segment order the hand-off without any collective: done[r] = last frame rank r has fully written,
    released = last frame the consumer (rank 0) is finished with.  x86 total store order + the stream
    synchronisation inside vr_render_owned_to_host make plain stores sufficient.
    Frame f (1, 2, ...) lives in buffer f % buffers.  With two buffers the producers may run one frame ahead of
    the consumer: a rank starts frame f as soon as frame f - buffers has been released, instead of waiting for
    the slowest rank and the consumer after every frame."""

    HEADER = 4096       # bytes: int64 done[world] at 0, int64 released at 2048

    def __init__(self, name: str, width: int, height: int, rank: int, world: int, create: bool, register_cuda: bool = True,
                 buffers: int = 1):
        self.rank, self.world, self.W, self.H, self.buffers = rank, world, width, height, buffers
        nbytes = self.HEADER + buffers * width * height * 16
        if create:
            # a tmpfs that is too small lets shm_open/ftruncate/mmap succeed and then kills the process with
            # SIGBUS on first touch: refuse up front (the caller falls back to rank 0 reading the frame back)
            try:
                st = os.statvfs("/dev/shm")
                free = st.f_bavail * st.f_frsize
            except OSError:
                free = None
            if free is not None and free < nbytes + (8 << 20):
                raise RuntimeError(f"/dev/shm has {free >> 20} MiB free, the shared frame needs {nbytes >> 20} MiB")
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=nbytes)
        self.created = create
        if not create:
            # the creator owns the name: keep this process's resource tracker from unlinking it at exit
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        buf = np.ndarray((nbytes,), dtype=np.uint8, buffer=self.shm.buf)
        self.done = buf[:8 * world].view(np.int64)
        self.released = buf[2048:2056].view(np.int64)
        self.frames = buf[self.HEADER:].view(np.float32).reshape(buffers, height, width, 4)
        self.frame = self.frames[0]
        if create:
            self.done[:] = 0
            self.released[:] = 0
        if create:
            self.numa_interleaved = _interleave_over_numa_nodes(self.frames.ctypes.data, buffers * width * height * 16)
        self.registered = False
        if register_cuda:
            rc = torch.cuda.cudart().cudaHostRegister(self.frames.ctypes.data, buffers * width * height * 16, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed: {rc}")
            self.registered = True

    @property
    def frame_ptr(self) -> int:
        return self.frame.ctypes.data

    def buffer_of(self, frame_no: int) -> np.ndarray:
        return self.frames[frame_no % self.buffers]

    def buffer_ptr(self, frame_no: int) -> int:
        return self.buffer_of(frame_no).ctypes.data

    def wait_writable(self, frame_no: int, timeout_s: float = 10.0):
        """Producer side, before writing frame `frame_no`: its buffer's previous occupant (frame_no - buffers) has
        been released by the consumer."""
        self.wait_released(frame_no - self.buffers, timeout_s)

    def wait_released(self, frame_no: int, timeout_s: float = 10.0):
        """Producer side: the consumer is finished with frame `frame_no` (the buffer may be overwritten)."""
        t0 = time.perf_counter()
        while int(self.released[0]) < frame_no:
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError(f"rank {self.rank}: frame {frame_no} was never released")

    def mark_done(self, frame_no: int):
        self.done[self.rank] = frame_no

    def wait_all_done(self, frame_no: int, timeout_s: float = 10.0):
        """Consumer side: every rank has written its rows of frame `frame_no`."""
        t0 = time.perf_counter()
        while int(self.done.min()) < frame_no:
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError(f"frame {frame_no}: done flags {self.done.tolist()}")

    def release(self, frame_no: int):
        self.released[0] = frame_no

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.frames.ctypes.data)
            self.registered = False
        self.done = self.released = self.frame = self.frames = None
        try:
            self.shm.close()
            if self.created:
                self.shm.unlink()
        except Exception:
            pass
