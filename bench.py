#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native volume ray march.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one frame of the hot path (ray generation, AABB test, front-to-back march with
windowing + compositing + early-ray termination) over BASELINE.json's headline workload:
1024^3 uint16 synthetic `mix` volume, 1920x1080, reference step
(1024 steps across the cube), trilinear filter, camera K2 (the volume fills the frame),
alpha scale 0.02.  Prints ONE JSON line (rank 0).

value      Mrays/s, whole job, device-timed (CUDA events), inputs resident in HBM.
e2e        same metric through the C-ABI render-to-host call: per frame the camera/uniform
           block goes host->device and the finished RGBA32F frame comes back to pinned host
           memory inside the timed region.
roofline   HBM bound: algorithmic bytes (2 B x distinct voxels referenced + 16 B x pixels)
           / average march-kernel duration, against MEASURED_PEAKS.json's copy bandwidth;
           `traffic` = DRAM bytes per launch from the ncu capture of THIS workload/kernel
           (profiles/r02/traffic.json, written by tools/gpu/r2_profile.sh), else null.
dense      the same frame with alpha so small that no ray reaches the 0.95 opacity threshold
           (every ray marches through the whole box): the honest samples/s figure.
cpu_baseline  the REFERENCE'S OWN SHADER compiled for the CPU (oracle/_ref/libshader_ref.so,
           kind "reference"; the restated oracle, kind "port", when that is absent or the frame
           uses an extension) on this box's host cores, on a bounded sample of rows of the SAME
           frame; doubles as a full-size parity check.

--impl reference times the reference's own algorithm on the host cores (its OpenGL path cannot
run: no GL stack on this image): VolumeRenderer.cs compiled for the CPU, all host threads.  That
arm loads nothing but oracle/ libraries (input generated on the host).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python")]

import numpy as np  # noqa: E402

METRIC = "Mrays/sec at 1920x1080, 1024^3 uint16, 1024 steps"
UNIT = "Mrays/s"
TILE_ROWS = 8            # = the CTA's 8 rows: 135 tiles over 8 ranks = 17 vs 16 (was 9 vs 8 with 16-row tiles)
DENSE_ALPHA = 0.001      # 1 - (1 - 0.001)^1774 = 0.83 < 0.95: early-ray termination cannot fire


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C4", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--camera", default="K2", choices=["K0", "K1", "K2"])
    ap.add_argument("--alpha", type=float, default=0.02)
    ap.add_argument("--filter", default="trilinear", choices=["nearest", "trilinear"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "direct", "texpair_pipe", "nearest_tex"])
    ap.add_argument("--cpu-row-stride", type=int, default=1, help="cpu_baseline renders every n-th row")
    ap.add_argument("--mip", action="store_true", help="maximum-intensity projection (the reference's use_mip toggle)")
    ap.add_argument("--tf", action="store_true", help="CubicSpline transfer function (default alpha knots of the reference's TF editor)")
    ap.add_argument("--window", type=int, nargs=2, default=None, metavar=("MIN", "MAX"), help="window uniforms (default: the config's, SURVEY.md 8d)")
    ap.add_argument("--skip", default="auto", choices=["auto", "on", "off"], help="result-identical empty-space skipping")
    ap.add_argument("--no-dense", action="store_true", help="skip the extra dense (no early-ray-termination) measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-count", action="store_true", help="skip the distinct-voxel instrumentation pass")
    ap.add_argument("--handoff", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU: peer = kernels store straight into rank 0's frame over NVLink (CUDA IPC), frame barrier in peer memory; "
                         "nccl = compact tiles, one NCCL gather, de-interleave kernel")
    return ap.parse_args()


def workload(args):
    from volren_b200 import workloads
    cfg = dict(workloads.CONFIGS[args.config])
    cfg["seed"] = workloads.SEEDS[args.config]
    if args.window:
        cfg["window"] = tuple(args.window)
    cfg["name"] = (f"{args.config}: {cfg['dims'][0]}x{cfg['dims'][1]}x{cfg['dims'][2]} uint{8 * cfg['bpv']} synthetic mix, "
                   f"{cfg['image'][0]}x{cfg['image'][1]}, step_scale {cfg['step_scale']}, camera {args.camera}, window [{cfg['window'][0]},{cfg['window'][1]}], "
                   f"alpha {args.alpha}, {args.filter}" + (", CubicSpline TF" if args.tf else "") + (", MIP" if args.mip else ""))
    cfg["key"] = f"{args.config}/{args.camera}/{args.filter}/{args.alpha}/{cfg['window'][0]}-{cfg['window'][1]}" + ("/tf" if args.tf else "") + ("/mip" if args.mip else "")
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "10", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def wait_for_samples(self, n=1, timeout_s=5.0):
        """nvidia-smi needs a moment to start: block until it has delivered n samples."""
        t0 = time.perf_counter()
        while self.proc and len(self.rows) < n and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.005)
        return len(self.rows)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def default_tf_lut():
    """256-entry opacity LUT of the reference's default alpha knots (AlphaControlSplineWidget.cpp:56-59)."""
    from volren_b200.host import CubicSpline
    return CubicSpline([(0, 0.0), (141, 0.759), (149, 0.45), (255, 1.0)]).bakeAlphaLUT()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic(key, kernel):
    """DRAM bytes per launch of the march kernel from the committed ncu capture of THIS workload (keyed by
    config/camera/filter/alpha/window and kernel name), or None: never a number measured on another frame."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02", "traffic.json")) as f:
            ent = json.load(f).get(key)
        if ent and ent.get("kernel") == kernel:
            return ent.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


def tf_lut_cpu():
    """The same LUT as default_tf_lut() from the oracle's spline (pinned to CubicSpline.cpp): no product library."""
    from oracle import orc
    return orc.OracleSpline([(0, 0.0), (141, 0.759), (149, 0.45), (255, 1.0)]).alpha_lut()


def cpu_march(cfg, args, host_vol, cam, row_stride, nthreads, lut=None):
    """The reference's own shader compiled for the CPU (or, for frames that use an extension, the restated oracle) on
    every row_stride-th row of the frame.  -> (Mrays/s, seconds, rows, image, counters, kind)"""
    from oracle import orc
    W, H = cfg["image"]
    p = orc.make_params(W, H, cfg["dims"], cfg["bpv"], cam, alpha_scale=args.alpha,
                        min_val=cfg["window"][0], max_val=cfg["window"][1],
                        filter=1 if args.filter == "trilinear" else 0, step_scale=cfg["step_scale"],
                        tf_lut=lut if args.tf else None, is_mip=1 if args.mip else 0,
                        row_begin=row_stride // 2, row_stride=row_stride)
    use_ref = orc.ref_shader_lib() is not None and cfg["step_scale"] == 1.0 and not args.tf
    t0 = time.perf_counter()
    if use_ref:
        img, cnt = orc.ref_render(p, host_vol, nthreads=nthreads)
    else:
        img, cnt, _ = orc.render(p, host_vol, nthreads=nthreads)
    dt = time.perf_counter() - t0
    rows = np.arange(row_stride // 2, H, row_stride)
    return rows.size * W / dt / 1e6, dt, rows, img, cnt, ("reference" if use_ref else "port")


CPU_KIND_NOTE = {"reference": "the reference's own VolumeRenderer.cs compiled for the CPU (oracle/_ref/libshader_ref.so)",
                 "port": "CPU restatement of VolumeRenderer.cs (oracle/march_oracle.c, pinned bit for bit to the reference shader); "
                         "used because this frame needs an extension the shader lacks (step override / transfer function) or oracle/_ref is absent"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = workload(args)
    from oracle import orc
    from volren_b200 import workloads          # pure numpy: loads no native library
    cam = workloads.camera_block_const(args.camera)
    W, H = cfg["image"]
    nthreads = os.cpu_count() or 1
    # input creation is not part of the timed path; generated on the host cores so that this arm touches oracle/ only
    host_vol = orc.synth_mix(cfg["dims"], cfg["bpv"], cfg["vmax"], cfg["seed"], True, nthreads)
    lut = tf_lut_cpu() if args.tf else None
    stride = max(args.cpu_row_stride * 4, 1)
    for _ in range(args.warmup):
        cpu_march(cfg, args, host_vol, cam, stride * 8, nthreads, lut)        # short warm-up frames
    times, rays, kind = [], 0, "port"
    for _ in range(args.steps):
        v, dt, rows, _, _, kind = cpu_march(cfg, args, host_vol, cam, stride, nthreads, lut)
        times.append(dt); rays = rows.size * W
    total = float(sum(times))
    value = rays * args.steps / total / 1e6
    sample = f"every {stride}th row of the {W}x{H} frame ({rays} rays per step), full march"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "note": "reference OpenGL path not runnable (no GL stack); this is " + CPU_KIND_NOTE[kind]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import volren_b200 as vb
    from volren_b200 import dist as vdist
    from volren_b200 import workloads

    rank, world, local_rank = vdist.init_from_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = workload(args)
    W, H = cfg["image"]
    dims, bpv = cfg["dims"], cfg["bpv"]
    cam = workloads.camera_block(args.camera)
    kernel = {"auto": vb.KERNEL_AUTO, "direct": vb.KERNEL_DIRECT, "texpair_pipe": vb.KERNEL_TEXPAIR_PIPE, "nearest_tex": vb.KERNEL_NEAREST_TEX}[args.kernel]
    params = vb.default_params(alpha_scale=args.alpha, min_val=cfg["window"][0], max_val=cfg["window"][1],
                               filter=vb.FILTER_TRILINEAR if args.filter == "trilinear" else vb.FILTER_NEAREST,
                               step_scale=cfg["step_scale"], kernel=kernel, tf_lut=default_tf_lut() if args.tf else None,
                               is_mip=1 if args.mip else 0,
                               empty_skip={"auto": vb.SKIP_AUTO, "on": vb.SKIP_ON, "off": vb.SKIP_OFF}[args.skip])

    ctx = vb.Context(W, H, device=local_rank)
    want_cpu = (rank == 0 and world == 1 and not args.no_cpu_baseline)
    nvox = dims[0] * dims[1] * dims[2]
    copy = torch.empty(nvox * bpv, dtype=torch.uint8, device=dev) if want_cpu else None
    ctx.upload_synthetic(dims, bpv, cfg["vmax"], cfg["seed"], True, copy_out_dptr=copy.data_ptr() if want_cpu else 0)
    host_vol = None
    if want_cpu:
        host_vol = copy.cpu().numpy().view(np.uint8 if bpv == 1 else np.uint16)
        del copy
        torch.cuda.empty_cache()
    ctx.set_camera(cam)
    ctx.set_params(params)
    ctx.set_partition(rank, world, TILE_ROWS)
    rows = ctx.owned_rows()

    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    local = torch.empty((rows, W, 4), dtype=torch.float32, device=dev)
    gathered = torch.empty((world, rows, W, 4), dtype=torch.float32, device=dev) if (world > 1 and rank == 0) else None
    frame = torch.empty((H, W, 4), dtype=torch.float32, device=dev) if rank == 0 else None
    pinned = torch.empty((H, W, 4), dtype=torch.float32).pin_memory() if rank == 0 else None

    kernel_ms = []
    launches = [0]
    used = [0, 0]

    # fused hand-off: every rank maps rank 0's TWO target frames (NVLink peer memory) and renders into them alternately
    peer_ptrs = []
    ctx_b = None                              # rank 0: a second context that only owns the second target frame
    cstream = None                            # rank 0: consumer stream (arrival wait + release), so its marches are not held back
    mstreams = None                           # one march stream per target frame: the tail of frame f overlaps the head of frame f + 1
    handoff = args.handoff if world > 1 else "none"
    if handoff == "peer":
        ok = 1
        try:
            if rank == 0:
                ctx_b = vb.Context(W, H, device=local_rank)
                cstream = torch.cuda.Stream()
            peer_ptrs.append(vdist.open_peer_frame(ctx, dst=0))
            peer_ptrs.append(vdist.open_peer_frame(ctx_b if rank == 0 else ctx, dst=0))
            mstreams = [torch.cuda.Stream(), torch.cuda.Stream()]
        except Exception as e:              # e.g. CUDA IPC unavailable in this container
            ok = 0
            if rank == 0:
                print(f"bench.py: peer hand-off unavailable ({e}); using the NCCL gather", file=sys.stderr)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        if int(flag.item()) == 0:
            handoff = "nccl"

    frame_no = [0]
    peer_uses = [0, 0]                        # frames rendered into each target frame so far

    def peer_target(f):
        """target frame of frame f: (buffer index, device pointer, its use count including f)"""
        b = f % 2
        return b, peer_ptrs[b], (f + 1) // 2 if b == 1 else f // 2

    if handoff == "peer":                     # which kernel / form runs (the asynchronous calls below do not report it)
        st0 = ctx.render_device(local.data_ptr(), compact=True, stream=sptr)
        used[0], used[1] = st0.kernel_used, st0.skip_used

    def step(release=True):
        if world == 1:
            st = ctx.render_device(frame.data_ptr(), compact=False, stream=sptr)
        elif handoff == "peer":
            # no collective call and no host synchronisation at all: completion and reuse of rank 0's frames are ordered
            # by the barrier words behind their pixels (NVLink peer memory).  Every rank: wait (on the device) until the
            # target frame's previous occupant has been consumed, march -- the kernel's last CTA publishes the arrival.
            # Rank 0, on its consumer stream: wait for all arrivals, consume, release.  Two target frames let every rank
            # run one frame ahead of the consumer, so the marches of consecutive frames run back to back.
            frame_no[0] += 1
            b, ptr, u = peer_target(frame_no[0])
            peer_uses[b] = u
            mptr = mstreams[b].cuda_stream
            ctx.peer_frame_release(ptr, u - 1, is_owner=False, stream=mptr)
            ctx.render_peer(ptr, frame_no[0], world, is_owner=False, stream=mptr, wait=False)
            launches[0] += 2
            if rank == 0:
                ctx.peer_frame_wait_arrivals(ptr, u, world, stream=cstream.cuda_stream)
                launches[0] += 1
                if release:
                    ctx.peer_frame_release(ptr, u, is_owner=True, stream=cstream.cuda_stream)
                    launches[0] += 1
            return
        else:
            st = ctx.render_device(local.data_ptr(), compact=True, stream=sptr)
            vdist.gather_tiles(local, gathered, dst=0)
            if rank == 0:
                ctx.assemble_tiles(gathered.data_ptr(), frame.data_ptr(), world, TILE_ROWS, stream=sptr)
                launches[0] += 1
        kernel_ms.append(st.kernel_ms)
        launches[0] += st.kernel_launches
        used[0] = st.kernel_used
        used[1] = st.skip_used

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # algorithmic bytes (instrumentation pass, not timed)
    counted = None
    if not args.no_count:
        counted = ctx.count_frame()

    for _ in range(max(args.warmup, 0)):
        step()
    kernel_ms.clear(); launches[0] = 0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_for_samples(1)
    barrier()
    idle_samples = len(sampler.rows)        # taken before the GPU was under load: not reported
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    if mstreams is not None and handoff == "peer":
        for m in mstreams:
            m.wait_event(ev0)                 # the march streams start after the clock
    first_timed = frame_no[0] + 1
    for _ in range(args.steps):
        step()
    if mstreams is not None and handoff == "peer":
        for m in mstreams:
            stream.wait_stream(m)             # every march has finished ...
    if cstream is not None:
        stream.wait_stream(cstream)           # ... and rank 0 has seen all arrivals of the last frame before the clock stops
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    timed_launches = launches[0]
    if handoff == "peer":                     # the marches' CUDA-event brackets of the timed frames (ring of 64 in the library)
        kernel_ms.extend(ctx.peer_kernel_ms(f) for f in range(max(first_timed, frame_no[0] - 63), frame_no[0] + 1))
    timed_kernel_ms = list(kernel_ms)

    # e2e: through the C-ABI with host buffers, copies inside the timed region.
    # N > 1: one host frame in shared memory, page-locked in every rank process; each rank copies its own
    # row tiles device->host over its own PCIe link (vr_render_owned_to_host) -- N links instead of rank 0's one.
    shared = None
    if world > 1:
        ok = 1
        name = f"volren_b200_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
        for creating in (True, False):                            # rank 0 creates, barrier, the others attach
            try:
                if creating == (rank == 0):
                    shared = vdist.SharedHostFrame(name, W, H, rank, world, create=creating, buffers=2)
            except Exception as e:
                ok = 0
                print(f"bench.py: rank {rank}: shared host frame unavailable ({e}); rank 0 reads the frame back alone", file=sys.stderr)
            if creating:
                torch.distributed.barrier()
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if shared is not None:
                shared.close()
            shared = None
    e2e_mode = ("single GPU: vr_render_submit / vr_render_wait, two frames in flight into two pinned host frames (row bands overlap each band's device->host copy); "
                "e2e.sync_value = the synchronous vr_render, one frame at a time") if world == 1 else \
               ("every rank copies its own row tiles (banded, overlapping its march) into a shared page-locked host frame; vr_render_submit / vr_render_wait, two frames in flight" if shared is not None
                else "hand-off to rank 0 on the device, rank 0 copies the frame to the host")
    pipelined = world == 1 or shared is not None
    pinned2 = [pinned, torch.empty((H, W, 4), dtype=torch.float32).pin_memory()] if world == 1 else None
    tickets = {}

    def complete(g):
        """frame g is in host memory (and, N > 1, every rank's rows of it: the consumer's hand-shake)"""
        ctx.render_wait(tickets.pop(g))
        if shared is not None:
            shared.mark_done(g)
            if rank == 0:
                shared.wait_all_done(g)                               # the whole frame is in host memory
                shared.release(g)

    # synchronous figure beside the pipelined one (single GPU): vr_render, one frame at a time
    sync_ms = None
    if world == 1:
        for i in range(args.warmup + args.steps):
            if i == args.warmup:
                ts = time.perf_counter()
            ctx.set_camera(cam)
            ctx.set_params(params)
            ctx.render_to_host_ptr(pinned.data_ptr())
        sync_ms = (time.perf_counter() - ts) * 1e3 / args.steps
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        f = i + 1
        if i == args.warmup:                  # the end-to-end path warms up like the device path (first-call
            if pipelined and (f - 1) in tickets:
                complete(f - 1)               # nothing in flight when the clock starts
            barrier()                         # stream/event creation, first copies into the pinned frame)
            t0 = time.perf_counter()
        ctx.set_camera(cam)
        ctx.set_params(params)
        if pipelined:
            # two host frames: frame f is submitted while frame f - 1 is still on the GPU / on its way to the host
            if shared is not None:
                shared.wait_writable(f)                               # the buffer's previous frame (f - 2) has been consumed
                tickets[f] = ctx.render_submit(shared.buffer_ptr(f))
            else:
                tickets[f] = ctx.render_submit(pinned2[f % 2].data_ptr())
            if (f - 1) in tickets:
                complete(f - 1)
        else:
            step(release=False)
            if rank == 0:
                if handoff == "peer":
                    b, ptr, u = peer_target(frame_no[0])
                    cstream.synchronize()                           # arrival wait done: the frame is complete
                    (ctx if b == 0 else ctx_b).read_frame_into(pinned.data_ptr())
                    ctx.peer_frame_release(ptr, u, is_owner=True, stream=cstream.cuda_stream)   # consumed: may be overwritten
                else:
                    pinned.copy_(frame, non_blocking=True)
                    torch.cuda.synchronize()
    if pipelined:
        complete(args.warmup + args.steps)    # the last frame is in host memory before the clock stops
    barrier()
    e2e_s = time.perf_counter() - t0
    # a short timed region (multi-GPU frames take < 1 ms) can end before nvidia-smi has sampled it
    # a few times: keep the same load running, untimed, until there are at least 5 samples under load
    need_more = torch.tensor([1 if (rank == 0 and len(sampler.rows) - idle_samples < 5) else 0], dtype=torch.int32, device=dev)
    for _ in range(200):
        if world > 1:
            torch.distributed.broadcast(need_more, src=0)
        if int(need_more.item()) == 0:
            break
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        need_more[0] = 1 if (rank == 0 and len(sampler.rows) - idle_samples < 5) else 0
    if rank == 0:
        sampler.rows = sampler.rows[idle_samples:]
    clocks = sampler.stop() if rank == 0 else None

    # N-GPU frame == 1-GPU frame (every kernel is bit-exact and pixels are independent)
    same_as_single = None
    host_same = None
    if world == 1:
        # both host frames of the pipelined end-to-end loop hold the frame the device-timed loop left on the GPU
        dev_frame = frame.cpu().numpy().view(np.uint32)
        host_same = bool(all(np.array_equal(pb.numpy().view(np.uint32), dev_frame) for pb in pinned2))
    if world > 1:
        step(release=False)
        if rank == 0:
            torch.cuda.synchronize()
            if handoff == "peer":
                b, ptr, u = peer_target(frame_no[0])
                multi = torch.from_numpy((ctx if b == 0 else ctx_b).read_frame()).to(dev)
                ctx.peer_frame_release(ptr, u, is_owner=True, stream=cstream.cuda_stream)
            else:
                multi = frame.clone()
        barrier()
        if rank == 0:
            ctx.set_partition(0, 1, TILE_ROWS)
            single = torch.empty((H, W, 4), dtype=torch.float32, device=dev)
            ctx.render_device(single.data_ptr(), compact=False, stream=sptr)
            torch.cuda.synchronize()
            same_as_single = bool(torch.equal(multi.view(torch.int32), single.view(torch.int32)))
            if shared is not None:
                last = args.warmup + args.steps                      # the last frame the e2e loop produced
                host_same = bool(np.array_equal(shared.buffer_of(last).view(np.uint32), single.cpu().numpy().view(np.uint32))
                                 and np.array_equal(shared.buffer_of(last - 1).view(np.uint32), single.cpu().numpy().view(np.uint32)))
            ctx.set_partition(rank, world, TILE_ROWS)
        barrier()

    if handoff == "peer" and rank == 0:
        torch.cuda.synchronize()
        for b in (0, 1):
            pst = ctx.peer_frame_status(peer_ptrs[b])
            if pst["timed_out"] or pst["arrivals"] != peer_uses[b] * world:
                raise SystemExit(f"bench.py: peer-memory frame barrier failed on target frame {b}: {pst}, expected {peer_uses[b] * world} arrivals")

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if shared is not None:
        if world > 1:
            torch.distributed.barrier()
        shared.close()
    if peer_ptrs and rank != 0:
        for ptr in peer_ptrs:
            ctx.frame_close_ipc(ptr)
    if rank != 0:
        ctx.close()
        if world > 1:
            torch.distributed.destroy_process_group()
        return 0

    rays_per_frame = W * H
    value = rays_per_frame * args.steps / (dev_ms * 1e-3) / 1e6
    e2e_value = rays_per_frame * args.steps / (e2e_ms * 1e-3) / 1e6
    avg_kernel_ms = float(np.mean(timed_kernel_ms))
    peak, peak_src = measured_peak()
    roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
            "peak_source": peak_src, "kernel": vb.KERNEL_NAMES.get(used[0], "?"), "empty_space_skipping": bool(used[1]),
            "kernel_ms_avg": avg_kernel_ms}
    if counted is not None:
        owned_px = sum(min(TILE_ROWS, H - t0_ * TILE_ROWS) for t0_ in range(rank, (H + TILE_ROWS - 1) // TILE_ROWS, world)) * W
        bytes_alg = bpv * counted["distinct_voxels"] + 16 * owned_px
        roof["achieved"] = bytes_alg / (avg_kernel_ms * 1e-3) / 1e9
        roof["frac"] = roof["achieved"] / peak
        roof["algorithmic_bytes_per_launch"] = bytes_alg
        roof["distinct_voxels"] = counted["distinct_voxels"]
        roof["samples_per_launch"] = counted["samples"]
        roof["rays_hit"] = counted["rays_hit"]
        roof["gsamples_per_s"] = counted["samples"] / (avg_kernel_ms * 1e-3) / 1e9
    wkey = cfg["key"] + ("/skip" if used[1] else "")
    roof["traffic"] = ncu_traffic(wkey, roof["kernel"])
    roof["workload_key"] = wkey

    # the same frame without early-ray termination: every ray marches through the whole box
    dense = None
    if world == 1 and not args.no_dense and not args.mip:
        dp = vb.default_params(alpha_scale=DENSE_ALPHA, min_val=cfg["window"][0], max_val=cfg["window"][1],
                               filter=params.filter, step_scale=cfg["step_scale"], kernel=kernel, tf_lut=default_tf_lut() if args.tf else None,
                               empty_skip=params.empty_skip)
        ctx.set_params(dp)
        dc = ctx.count_frame() if not args.no_count else None
        for _ in range(3):
            ctx.render_device(frame.data_ptr(), compact=False, stream=sptr)
        dms = [ctx.render_device(frame.data_ptr(), compact=False, stream=sptr).kernel_ms for _ in range(10)]
        dense = {"alpha": DENSE_ALPHA, "kernel_ms_avg": float(np.mean(dms)), "mrays_per_s": W * H / (float(np.mean(dms)) * 1e-3) / 1e6,
                 "max_alpha_in_frame": float(frame[..., 3].max().item())}
        if dc is not None:
            dense["samples_per_launch"] = dc["samples"]
            dense["gsamples_per_s"] = dc["samples"] / (dense["kernel_ms_avg"] * 1e-3) / 1e9
            dense["roofline_frac"] = (bpv * dc["distinct_voxels"] + 16 * W * H) / (dense["kernel_ms_avg"] * 1e-3) / 1e9 / peak
        ctx.set_params(params)

    cpu = None
    if want_cpu:
        nthreads = os.cpu_count() or 1
        v, dt, rows_idx, img_cpu, cnt, kind = cpu_march(cfg, args, host_vol, cam, args.cpu_row_stride, nthreads, default_tf_lut() if args.tf else None)
        ctx.render_device(frame.data_ptr(), compact=False, stream=sptr)
        torch.cuda.synchronize()
        img_gpu = frame.cpu().numpy()
        err = float(np.abs(img_gpu[rows_idx].astype(np.float64) - img_cpu[rows_idx].astype(np.float64)).max())
        cpu = {"value": v, "unit": UNIT, "cores": nthreads, "kind": kind, "what": CPU_KIND_NOTE[kind],
               "sample": f"every {args.cpu_row_stride}th row of the same {W}x{H} frame ({rows_idx.size * W} rays, "
                         f"{cnt['samples']} samples, {dt:.1f} s of CPU work on {nthreads} threads)",
               "parity_max_abs_err_on_sample": err,
               "parity_bit_exact_on_sample": bool(np.array_equal(img_gpu[rows_idx].view(np.uint32), img_cpu[rows_idx].view(np.uint32)))}

    d2h = W * H * 16
    h2d = 84 + 4 * 10 + 1024 + 4
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "parallelism": f"screen-row tiles of {TILE_ROWS} rows interleaved over {world} GPU(s), replicated volume, hand-off: "
                                  + {"none": "n/a", "peer": "march kernels (one stream per target frame, so consecutive frames overlap) store into one of rank 0's two target frames over NVLink peer memory and their last CTA publishes the arrival in the frame barrier words of the same peer memory; rank 0 waits / releases on a consumer stream (no collective call, no signal kernel, no host synchronisation between frames)", "nccl": "one NCCL gather + de-interleave"}[handoff],
                   "multi_gpu_frame_equals_single_gpu_frame": same_as_single,
                   "e2e_path": e2e_mode, "multi_gpu_host_frame_equals_single_gpu_frame": host_same if world > 1 else None, "host_frames_equal_device_frame": host_same,
                   "l2": f"inputs larger than L2 ({nvox * bpv / 2**30:.2f} GiB volume vs 126 MB L2); no flush needed" if nvox * bpv > 2**28
                         else "volume fits in L2 (correctness/plumbing config)",
                   "kernel": roof["kernel"]},
        "roofline": roof, "dense": dense, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps,
                "sync_value": (rays_per_frame / (sync_ms * 1e-3) / 1e6) if sync_ms else None, "sync_ms_per_step": sync_ms},
        "gpu_launches": timed_launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if ctx_b is not None:
        ctx_b.close()
    ctx.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
