#!/usr/bin/env python3
"""bench.py - headline benchmark of the render hot path (BASELINE.json metric).

A "step" is one Renderer::Update() pass (1 spp) of the workload over the whole frame.
  value     Mrays/s (primary + shadow + secondary rays of all ranks / max-over-ranks device time),
            accumulator resident in HBM, frames enqueued with VP_FLAG_DEFER (four in flight per GPU)
            and timed with one CUDA-event pair around exactly K steps.
  e2e       the same metric through the reference-facing C-ABI calls with HOST buffers (page-locked):
            every step uploads the float4 accumulator and downloads accumulator + RGB8 image.
            vp_render_async / vp_render_wait keep six frames in flight (the reference's Tick only
            needs frame k's pixels when it blits them); the synchronous vp_render figure is reported
            next to it.  With N > 1 GPUs rank 0 drives all N GPUs through ONE multi-device context
            (vp_create_multi): one C call per step, the frame assembled in ONE host buffer.
  roofline  dominant kernel (k_wf_extend, the Renderer::FindNearest walk): algorithmic bytes
            (SURVEY.md 8d formula) / its average device time measured by CUDA events inside
            the timed region, against the measured HBM copy bandwidth.
  cpu_baseline  the reference's own CPU Trace (oracle/_ref, the build without instrumentation inside
            the DDA loops) timed on this box's host cores.
`--impl reference` times that CPU path alone with the same config / metric; it reads its inputs from
data/bench/ (plain arrays) and maps no library of this repository.
Multi-GPU (torchrun, one rank per GPU): the frame is sharded as interleaved 8-row tiles (scene replicated,
no per-sample communication).  The exchange is fused into the resolve stage: every rank's k_wf_resolve stores
its RGB8 tiles straight into one image in rank 0's HBM over NVLink (vp_share_*: CUDA IPC mapping + system-scope
arrival flags) - no pack kernel, no all_gather, no de-interleave; torch.distributed only carries the 128-byte
mapping handle and the barriers.  VP_BENCH_EXCHANGE=nccl selects the round-1 exchange (pack -> NCCL all_gather ->
de-interleave) for A/B.  Total work is fixed => "scaling": "strong".
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# Passes in flight use two streams each (4-6 passes); the default of 8 hardware queues makes streams
# share queues and serialises kernels that do not depend on each other (a rank's share of an 8-GPU frame: 0.486 -> 0.456 ms).
# Read by the driver when the CUDA context is created, so it is set before anything touches CUDA; the library's
# load-time constructor does the same for hosts that load it first.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
vp = importlib.import_module("raytracer-voxpopuli_b200")
scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
abi = vp.abi

METRIC = "Mrays/s (primary+shadow+secondary)"
ROW_TILE = 8
DEFOCUS_JITTER = float(os.environ.get("VP_BENCH_DEFOCUS", "2.0"))  # defocusJitter = 2.0 for throughput runs (0 for the bit-exact-hit runs of tests/)
DATA = "reference assets (room.vox grid fixture), synthetic camera/lights"


def frame_params(fpd, **over):
    d = dict(fpd)
    d.update(over)
    return vp.frame_params(d["width"], d["height"], max_bounces=d["max_bounces"], spp=d.get("spp", 1),
                           frame_index=d.get("frame_index", 0), seed=scenes.SEED, aa_strength=d["aa_strength"],
                           row_begin=d.get("row_begin", 0), row_end=d.get("row_end", 0), row_tile=d.get("row_tile", 0), row_tile_stride=d.get("row_tile_stride", 1),
                           row_tile_phase=d.get("row_tile_phase", 0), flags=d.get("flags", 0))


def base_config(cfg_idx, fpd):
    """The `config` object of the JSON line: identical in both arms."""
    return {"workload": scenes.WORKLOAD_NAMES[cfg_idx], "resolution": [fpd["width"], fpd["height"]], "spp_per_step": 1,
            "max_bounces": fpd["max_bounces"], "defocus_jitter": DEFOCUS_JITTER}


def shard_params(world, rank):
    return dict(row_tile=ROW_TILE, row_tile_stride=world, row_tile_phase=rank) if world > 1 else {}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled through NVML during the timed region
    (same fields as the nvidia-smi clocks line of B200_PROFILING.md)."""

    def __init__(self, gpu_index, period_s=0.005):
        super().__init__(daemon=True)
        self.gpu, self.period, self.stop_flag = gpu_index, period_s, False
        self.sm, self.reasons, self.max_sm, self.err = [], set(), None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                    "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while not self.stop_flag:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
                time.sleep(self.period)
        except Exception as e:  # NVML unavailable: report it rather than invent numbers
            self.err = repr(e)

    def mark(self):
        """start of the timed region: drop what was sampled before it"""
        self.sm, self.reasons = [], set()

    def finish(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


def load_checker():
    """The reference itself when its build travelled with the repo - the variant without instrumentation inside the DDA
    loops (oracle/ref/build_ref.py) - else the plain-C port.  The CPU path gets every core this process may run on,
    whatever OMP_NUM_THREADS says (torchrun sets it to 1 for its workers)."""
    for name, kind in (("libvpref_fast.so", "reference"), ("libvpref.so", "reference (instrumented build)")):
        ref = os.path.join(ROOT, "oracle", "_ref", name)
        if os.path.exists(ref):
            lib = abi.Lib(ref, "vpref_")
            break
    else:
        so = os.path.join(ROOT, "oracle", "libvporacle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libvporacle.so"])
        lib, kind = abi.Lib(so, "vpo_"), "port"
    threads = getattr(lib.dll, lib.prefix + "threads")
    threads.restype = C.c_int
    setter = getattr(lib.dll, lib.prefix + "set_threads", None)
    try:
        want = len(os.sched_getaffinity(0))
    except AttributeError:
        want = os.cpu_count() or 1
    if setter is not None and want > 0:
        setter(want)
    return lib, kind, threads()


def cpu_frame_times(lib, sc, cam, fpd, reps, first_frame=0):
    """Times `reps` full-frame Update() passes of the CPU path; returns (seconds list, rays list)."""
    W, H = fpd["width"], fpd["height"]
    times, rays = [], []
    with vp.Session(lib) as s:
        s.set_scene(sc)
        acc = np.zeros((H, W, 4), np.float32)
        for i in range(reps):
            t0 = time.perf_counter()
            acc, _, st = s.render(cam, frame_params(fpd, frame_index=first_frame + i), accum=acc)
            times.append(time.perf_counter() - t0)
            rays.append(st.rays)
    return times, rays


def bench_inputs(cfg_idx):
    """The workload as plain arrays (data/bench/, numpy only); throughput runs use the survey's defocus jitter."""
    sc, cam, fpd = scenes.load_inputs(os.path.join(scenes.BENCH_DIR, f"cfg{cfg_idx}.npz"))
    cam.defocus_jitter = DEFOCUS_JITTER
    return sc, cam, fpd


def run_reference(args, cfg_idx):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc, cam, fpd = bench_inputs(cfg_idx)
    lib, kind, cores = load_checker()
    times, rays = cpu_frame_times(lib, sc, cam, fpd, args.warmup + args.steps)
    t, r = times[args.warmup:], rays[args.warmup:]
    total = sum(t)
    value = sum(r) / total / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": DATA,
            "config": base_config(cfg_idx, fpd), "rays_per_step": float(np.mean(r)),
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind,
                             "sample": f"{args.steps} full {fpd['width']}x{fpd['height']} frames (1 spp each) after {args.warmup} warm-up"},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class NcclGather:
    """Round-1 exchange, kept for A/B (VP_BENCH_EXCHANGE=nccl): pack own tiles -> one all_gather -> de-interleave."""

    def __init__(self, frame, world, rank):
        sharding = importlib.import_module("raytracer-voxpopuli_b200.sharding")
        self.g = sharding.FrameGather(frame, world, rank)

    def __call__(self, frame):
        return self.g(frame)


def async_e2e(lib, sess, cc, fpd, steps, W, H, torch, in_flight=5):
    """Update() through vp_render_async with page-locked host buffers: every step uploads the accumulator and downloads
    accumulator + image; two accumulators / images alternate so that the host could read frame k while k+1 is in flight."""
    h_acc = [torch.zeros((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    h_rgb = [torch.zeros((H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
    tickets = []

    def issue(i):
        fp = frame_params(fpd, frame_index=i)
        t = C.c_uint64()
        lib.check(lib.c_render_async(sess.ctx, C.byref(cc), C.byref(fp), h_acc[i & 1].data_ptr(), h_acc[(i + 1) & 1].data_ptr(),
                                     h_rgb[i & 1].data_ptr(), C.byref(t)), "render_async")
        tickets.append(t.value)
        if len(tickets) > in_flight:
            lib.check(lib.c_render_wait(sess.ctx, tickets[-1 - in_flight]), "render_wait")

    WARM = 12  # every lane of every device has allocated its path state and captured its graph (six lanes, eight for small shards)
    for i in range(WARM):
        issue(i)
    lib.check(lib.c_render_wait(sess.ctx, tickets[-1]), "render_wait")
    lib.check(lib.c_stats_begin(sess.ctx), "stats_begin")
    t0 = time.perf_counter()
    for i in range(WARM, WARM + steps):
        issue(i)
    lib.check(lib.c_render_wait(sess.ctx, tickets[-1]), "render_wait")
    dt = time.perf_counter() - t0
    st = abi.VpStats()
    lib.check(lib.c_stats_end(sess.ctx, C.byref(st)), "stats_end")
    return st.rays / dt / 1e6, 1e3 * dt / steps, st.kernel_launches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="workload (SURVEY.md 8d); 2 = room.vox 1920x1080, the metric's config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the short runs of the other configs (config.also)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args, args.config)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # barriers around phases in which only rank 0 works wait on the HOST (gloo): an NCCL barrier is a kernel that spins
        # on every waiting rank's GPU, and a GPU shared by two processes is time-sliced between them
        host_group = dist.new_group(backend="gloo")
    dev = torch.device("cuda", local)

    lib = vp.load_product()
    sc, cam, fpd = bench_inputs(args.config)
    W, H = fpd["width"], fpd["height"]
    shard = shard_params(world, rank)
    sess = vp.Session(lib, local)
    sess.set_scene(sc)
    cc = cam.to_c()
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    # Frames in flight: a step is enqueued with VP_FLAG_DEFER, so its latency-bound tail overlaps the next step's wide first
    # bounces.  With several ranks every frame in flight needs its own image (its tiles land in rank 0's HBM while later
    # frames are traced), plus two so that a frame never waits for the consumer of the frame that last used its image.
    NBUF = int(os.environ.get("VP_BENCH_IMAGES", "8")) if world > 1 else 1
    stream = torch.cuda.current_stream(dev)
    comm = torch.cuda.Stream(dev) if world > 1 else None
    img_bytes = W * H * 4

    # ---- exchange set-up (N > 1): one image ring in rank 0's HBM, mapped by every rank ----------------------------
    exchange = "none"
    share, img_ptrs, rgb8s, gather, gathered = None, None, None, None, [None] * NBUF
    if world > 1 and os.environ.get("VP_BENCH_EXCHANGE", "share") == "share":
        info = abi.VpShareInfo()
        ok = 1
        if rank == 0:
            share = lib.c_share_create(sess.ctx, NBUF * img_bytes, C.byref(info))
            ok = 1 if share else 0
        blob = torch.frombuffer(bytearray(bytes(info)), dtype=torch.uint8).to(dev)
        dist.broadcast(blob, 0)
        if rank != 0:
            info = abi.VpShareInfo.from_buffer_copy(bytes(blob.cpu().numpy().tobytes()))
            share = lib.c_share_open(sess.ctx, C.byref(info)) if info.bytes else None
            ok = 1 if share else 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            exchange = "peer stores from k_wf_resolve into rank 0's image (vp_share: CUDA IPC + NVLink), arrival flags"
            base = lib.c_share_ptr(share)
            img_ptrs = [base + k * img_bytes for k in range(NBUF)]
        else:
            if rank == 0:
                print(f"bench: shared image unavailable ({lib.c_last_error().decode()}), falling back to NCCL all_gather", file=sys.stderr)
            if share:
                lib.c_share_close(share)
            share = None
    if world > 1 and share is None:
        exchange = "pack -> NCCL all_gather -> de-interleave (torch.distributed)"
        rgb8s = [torch.zeros((H, W), dtype=torch.int32, device=dev) for _ in range(NBUF)]
        gather = NcclGather(rgb8s[0], world, rank)
    if world == 1:
        rgb8s = [torch.zeros((H, W), dtype=torch.int32, device=dev)]

    def fence(on, back=0):
        lib.check(lib.c_render_fence(sess.ctx, on.cuda_stream, back), "render_fence")

    lone_ms = [0.0]

    def step(i, flags=abi.VP_FLAG_DEFER):
        k = i % NBUF
        if share is not None:
            if i >= NBUF:  # the consumer (rank 0) has released the image this frame is about to overwrite
                lib.check(lib.c_share_wait(share, stream.cuda_stream, 32, 1, i - NBUF + 1, 20.0), "share_wait(release)")
            img = img_ptrs[k]
        else:
            if gathered[k] is not None:
                stream.wait_event(gathered[k])  # the gather that last read this image
            img = rgb8s[k].data_ptr()
        fp = frame_params(fpd, frame_index=i, flags=flags, **shard)
        st1 = abi.VpStats() if (flags & abi.VP_FLAG_TIME_STAGES) else None  # a step that runs alone also reports its own device time
        lib.check(lib.c_render_device(sess.ctx, C.byref(cc), C.byref(fp), accum.data_ptr(), img, stream.cuda_stream,
                                      C.byref(st1) if st1 is not None else None), "render_device")
        if st1 is not None:
            lone_ms[0] += st1.device_ms
        if world > 1:
            with torch.cuda.stream(comm):
                fence(comm)   # this frame resolved: its tiles are stored
                if share is not None:
                    lib.check(lib.c_share_signal(share, comm.cuda_stream, rank, i + 1), "share_signal")
                    if rank == 0:
                        lib.check(lib.c_share_wait(share, comm.cuda_stream, 0, world, i + 1, 20.0), "share_wait(arrival)")
                        # image k now holds every rank's tiles of frame i (a display would read it here); release it
                        lib.check(lib.c_share_signal(share, comm.cuda_stream, 32, i + 1), "share_signal(release)")
                else:
                    gather(rgb8s[k])
                    gathered[k] = torch.cuda.Event()
                    gathered[k].record(comm)

    def drain():
        if world > 1:
            stream.wait_stream(comm)
        fence(stream)

    # The first timed step runs alone (pipeline empty after the warm-up, not deferred) with CUDA events around its
    # extend-stage launches: the roofline's kernel duration is measured inside the timed region, without other frames'
    # kernels sharing the SMs.  (Round 1 ran every 20th step alone; draining the pipeline that often costs a 1/8-frame
    # shard 12 % of its step time.)
    sampler = ClockSampler(local, period_s=float(os.environ.get("VP_BENCH_CLOCK_PERIOD", "0.02")))
    if os.environ.get("VP_BENCH_NO_SAMPLER") != "1":
        sampler.start()
    # set-up, not warm-up: the first frame on each of the library's lanes (six, eight for small shards) allocates that lane's path state and
    # captures its CUDA graph (into a scratch accumulator; frame indices of the measured frames are untouched)
    scratch = torch.zeros_like(accum)
    for k in range(12):
        fp0 = frame_params(fpd, frame_index=0, flags=abi.VP_FLAG_DEFER, **shard)
        lib.check(lib.c_render_device(sess.ctx, C.byref(cc), C.byref(fp0), scratch.data_ptr(), None, stream.cuda_stream, None), "render_device(set-up)")
    fence(stream)
    torch.cuda.synchronize(dev)
    del scratch
    for i in range(args.warmup):
        step(i)
    drain()
    torch.cuda.synchronize(dev)

    lib.check(lib.c_stats_begin(sess.ctx), "stats_begin")
    sampler.mark()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t_begin.record(stream)
    # N = 1: the first timed step runs alone with stage events (the roofline's kernel time is measured inside the timed
    # region).  N > 1: a step that runs alone costs every rank a pipeline fill on top of a 0.5 ms step, so all K steps
    # are pipelined and the stage-timed step is one EXTRA step after the timed region (roofline.stage_events says which).
    staged_inside = world == 1
    for i in range(args.steps):
        step(args.warmup + i, flags=abi.VP_FLAG_TIME_STAGES if (i == 0 and staged_inside) else abi.VP_FLAG_DEFER)
    drain()
    t_end.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    total_ms = float(t_begin.elapsed_time(t_end))
    clocks = sampler.finish()
    st = abi.VpStats()
    lib.check(lib.c_stats_end(sess.ctx, C.byref(st)), "stats_end")
    if not staged_inside:
        acc_x = torch.zeros_like(accum)
        img_x = torch.zeros((H, W), dtype=torch.int32, device=dev)
        fpx = frame_params(fpd, frame_index=args.warmup + args.steps, flags=abi.VP_FLAG_TIME_STAGES, **shard)
        stx = abi.VpStats()
        lib.check(lib.c_render_device(sess.ctx, C.byref(cc), C.byref(fpx), acc_x.data_ptr(), img_x.data_ptr(), stream.cuda_stream, C.byref(stx)),
                  "render_device(stage events)")
        lone_ms[0] += stx.device_ms
        del acc_x, img_x
    ext_ms, ext_n = C.c_double(), C.c_uint64()
    lib.check(lib.c_stage_times(sess.ctx, C.byref(ext_ms), C.byref(ext_n)), "stage_times")
    share_err = 0
    if share is not None:
        e = C.c_uint32()
        lib.check(lib.c_share_error(share, C.byref(e)), "share_error")
        share_err = int(e.value)
    t = torch.tensor([total_ms, float(st.rays), float(st.kernel_launches), float(share_err)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, rays_total, launches, share_err = float(tmax[0]), float(tsum[1]), float(tsum[2]), int(tmax[3])
    else:
        rays_total, launches = float(st.rays), float(st.kernel_launches)
    if share_err:
        raise SystemExit(f"bench: a rank never arrived at the shared image (flag slot {share_err - 1}); no number reported")
    value = rays_total / (total_ms * 1e-3) / 1e6

    # the assembled image on rank 0 == the image a single GPU renders (checked once, outside the timed region)
    exchange_check = None
    if share is not None:
        last = args.warmup + args.steps - 1
        if rank == 0:
            class _Dev:
                def __init__(self, ptr):
                    self.__cuda_array_interface__ = dict(shape=(H, W), typestr="<i4", data=(int(ptr), False), version=2)
            got = torch.as_tensor(_Dev(img_ptrs[last % NBUF]), device=dev).clone()
            # frame `last` of a running average only depends on the accumulator history: re-render all frames unsharded
            acc1 = torch.zeros_like(accum)
            one = torch.zeros((H, W), dtype=torch.int32, device=dev)
            fp = frame_params(fpd, frame_index=0, spp=last + 1)
            lib.check(lib.c_render_device(sess.ctx, C.byref(cc), C.byref(fp), acc1.data_ptr(), one.data_ptr(), stream.cuda_stream, None), "render_device(check)")
            torch.cuda.synchronize(dev)
            exchange_check = {"pixels_differing_from_single_gpu_image": int((got != one).sum().item()), "frames": last + 1}
        dist.barrier(group=host_group)

    # ---- e2e: the reference-facing C-ABI calls with host buffers, copies inside the timed region ----
    e2e_steps = max(3, min(args.steps, 20))
    e2e = None
    if world == 1:
        av, ams, _ = async_e2e(lib, sess, cc, fpd, e2e_steps, W, H, torch)
        h_acc = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
        h_rgb = torch.zeros((H, W), dtype=torch.int32).pin_memory()
        s_rays, s_t = 0, 0.0
        for i in range(2 + min(e2e_steps, 10)):
            fp = frame_params(fpd, frame_index=i)
            s1 = abi.VpStats()
            t0 = time.perf_counter()
            lib.check(lib.c_render(sess.ctx, C.byref(cc), C.byref(fp), h_acc.data_ptr(), h_rgb.data_ptr(), C.byref(s1)), "render")
            dt = time.perf_counter() - t0
            if i >= 2:
                s_rays += s1.rays
                s_t += dt
        e2e = {"value": av, "unit": "Mrays/s", "h2d_bytes_per_step": W * H * 16, "d2h_bytes_per_step": W * H * 16 + W * H * 4, "ms_per_step": ams,
               "api": "vp_render_async + vp_render_wait (page-locked host float4 accumulator in / out + RGB8 out every step, 6 frames in flight)",
               "steps": e2e_steps, "timer": "host wall clock around the steps, last ticket waited for",
               "synchronous": {"value": s_rays / s_t / 1e6, "ms_per_step": 1e3 * s_t / min(e2e_steps, 10), "api": "vp_render (one blocking call per frame)"}}
    elif rank == 0:
        # one process, one multi-device context over all N GPUs, ONE host accumulator / image
        g = vp.Session(lib, list(range(world)))
        g.set_scene(sc)
        av, ams, _ = async_e2e(lib, g, cc, fpd, e2e_steps, W, H, torch)
        ginfo = g.info()
        g.close()
        e2e = {"value": av, "unit": "Mrays/s", "h2d_bytes_per_step": W * H * 16, "d2h_bytes_per_step": W * H * 16 + W * H * 4, "ms_per_step": ams,
               "api": f"vp_create_multi({world} devices) + vp_render_async / vp_render_wait from ONE process: one C call per step, every GPU moves its own "
                      "row tiles over its own PCIe link into ONE page-locked host accumulator / image, 6 frames in flight",
               "steps": e2e_steps, "timer": "host wall clock around the steps, last ticket waited for", "peer_access": ginfo.get("peer_access")}
    if world > 1:
        dist.barrier(group=host_group)

    if rank != 0:
        if share is not None:
            lib.c_share_close(share)
        dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: k_wf_extend (Renderer::FindNearest stage) -----------------------
    # duration: CUDA events recorded by the library around every extend launch of the timed region, on the
    # stream they were launched on; algorithmic bytes (SURVEY.md 8d terms of this stage): one byte per DDA cell
    # visit + 128 B of volume transform per ray and volume, counted by an instrumented (untimed) pass.
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    inst = abi.VpStats()
    acc2 = torch.zeros_like(accum)
    one = torch.zeros((H, W), dtype=torch.int32, device=dev)
    lib.check(lib.c_stats_begin(sess.ctx), "stats_begin")
    fp = frame_params(fpd, frame_index=args.warmup, flags=abi.VP_FLAG_COUNT_STEPS, **shard)
    lib.check(lib.c_render_device(sess.ctx, C.byref(cc), C.byref(fp), acc2.data_ptr(), one.data_ptr(), stream.cuda_stream, C.byref(inst)),
              "render_device(count steps)")
    meta = sess.info()
    nvol = len(sc.volumes)
    own_rows = sum(min(ROW_TILE, H - k * ROW_TILE) for k in range(rank, (H + ROW_TILE - 1) // ROW_TILE, world))
    px_rank = own_rows * W if world > 1 else W * H
    ext_rays = inst.primary_rays + inst.secondary_rays
    timed_steps = 1
    ext_launches_per_frame = max(1, ext_n.value // max(1, timed_steps))
    ext_bytes_frame = meta["extend_steps"] * 1 + ext_rays * nvol * 128
    frame_bytes = inst.dda_steps * 1 + (inst.rays + inst.exit_rays) * nvol * 128 + px_rank * 32 + px_rank * 4 + inst.sky_samples * 12
    ext_ms_frame = ext_ms.value / max(1, timed_steps)
    achieved = ext_bytes_frame / (ext_ms_frame * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k_wf_extend_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # sector-granular view: every DDA cell visit of the reference is a dependent byte load = one 32 B
    # sector; against the L2 read bandwidth measured here by csrc/vp_probe.cu (48 MiB buffer, ld.global.cg)
    l2 = None
    probe_so = os.path.join(ROOT, "raytracer-voxpopuli_b200", "libvoxpop_probe.so")
    if os.path.exists(probe_so):
        pl = C.CDLL(probe_so)
        pl.vp_probe_l2_read.argtypes = [C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        gbs = C.c_double()
        if pl.vp_probe_l2_read(local, 48 << 20, 40, C.byref(gbs)) == 0:
            sector_bytes = meta["extend_steps"] * 32
            l2 = {"peak": gbs.value, "unit": "GB/s", "peak_source": "measured here: 48 MiB L2-resident buffer, 128-bit ld.global.cg, 1184 blocks",
                  "sector_bytes_per_step": int(sector_bytes), "achieved": sector_bytes / (ext_ms_frame * 1e-3) / 1e9,
                  "note": "what the reference's one-byte-per-cell loads would move at sector granularity; the kernel itself "
                          "replaces them by register bit tests (DESIGN.md section 3), its measured L2 traffic is in profiles/"}
            l2["frac"] = l2["achieved"] / l2["peak"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": meta.get("extend_kernel", "k_wf_extend"), "peak_source": peak_src,
                "launches_per_step": int(ext_launches_per_frame), "kernel_ms_per_step": ext_ms_frame, "steps_with_stage_events": timed_steps,
                "stage_events": "the first step of the timed region" if staged_inside else "one extra step after the timed region (its rays and time are not in `value`)",
                "kernel_share_of_step": ext_ms.value / lone_ms[0] if lone_ms[0] > 0 else None,
                "kernel_share_note": "extend-stage time / device time of the same steps (the steps that run alone with stage events); "
                                     "against the average pipelined step it is kernel_ms_per_step / ms_per_step",
                "algorithmic_bytes_per_launch": int(ext_bytes_frame / ext_launches_per_frame),
                "algorithmic_bytes_per_step_all_stages": int(frame_bytes),
                "dda_steps_per_step": int(inst.dda_steps), "l2_sector_view": l2,
                "note": "traversal of an L2-resident grid, bound by the ALU pipe / divergence / latency, not by HBM: the HBM "
                        "fraction is a few percent by construction (DESIGN.md section 4)"}

    cpu_baseline, also = None, None
    if world == 1 and not args.no_cpu_baseline:
        clib, kind, cores = load_checker()
        reps = 4
        times, rays = cpu_frame_times(clib, sc, cam, fpd, reps)
        best = int(np.argmin(times[1:])) + 1
        cpu_baseline = {"value": rays[best] / times[best] / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
                        "ms_per_frame": 1e3 * times[best],
                        "sample": f"best of {reps - 1} full {W}x{H} frames (1 spp) after 1 warm-up, same scene/camera/seeds"}
        # the 1-thread figure, on a BAND of rows around the image centre (not a whole frame)
        setter = getattr(clib.dll, clib.prefix + "set_threads", None)
        if setter is not None:
            band = (H // 2 - 32, H // 2 + 32)
            setter(1)
            with vp.Session(clib) as s1:
                s1.set_scene(sc)
                acc1 = np.zeros((H, W, 4), np.float32)
                t0 = time.perf_counter()
                _, _, st1 = s1.render(cam, frame_params(fpd, frame_index=0, row_begin=band[0], row_end=band[1]), accum=acc1)
                dt1 = time.perf_counter() - t0
            setter(cores)
            cpu_baseline["one_thread"] = {"value": st1.rays / dt1 / 1e6, "unit": "Mrays/s",
                                          "sample": f"64-row band (rows {band[0]}..{band[1] - 1}) of one {W}x{H} frame, 1 thread"}
    if world == 1 and not args.no_also:
        # the other configurations of BASELINE.json and room.vox with its native (glass) materials, device-resident,
        # six frames in flight, 1 spp per step at full size: short runs so that they are driver-visible
        also = {}
        for idx in (1, 3, 4, 5, 6):
            sc_i, cam_i, fpd_i = bench_inputs(idx)
            Wi, Hi = fpd_i["width"], fpd_i["height"]
            sess.set_scene(sc_i)
            cci = cam_i.to_c()
            acc_i = torch.zeros((Hi, Wi, 4), dtype=torch.float32, device=dev)
            rgb_i = torch.zeros((Hi, Wi), dtype=torch.int32, device=dev)
            n_i = 12

            def run(first, count):
                for f in range(first, first + count):
                    fp_i = frame_params(fpd_i, frame_index=f, spp=1, flags=abi.VP_FLAG_DEFER)
                    lib.check(lib.c_render_device(sess.ctx, C.byref(cci), C.byref(fp_i), acc_i.data_ptr(), rgb_i.data_ptr(), stream.cuda_stream, None), "render_device")
                fence(stream)

            run(0, 12)  # set-up: every lane (six; eight for small calls) has sized its path state for this frame and captured its graph
            torch.cuda.synchronize(dev)
            lib.check(lib.c_stats_begin(sess.ctx), "stats_begin")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run(12, n_i)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            sti = abi.VpStats()
            lib.check(lib.c_stats_end(sess.ctx, C.byref(sti)), "stats_end")
            ms = float(e0.elapsed_time(e1))
            key = "cfg2g_native_glass" if idx == 6 else f"cfg{idx}"
            also[key] = {"workload": scenes.WORKLOAD_NAMES[idx], "value": sti.rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": ms / n_i,
                         "rays_per_frame": sti.rays / n_i, "steps": n_i, "hbm_in_use": sess.info()["hbm_in_use"]}
            del acc_i, rgb_i
        sess.set_scene(sc)

    cfg = base_config(args.config, fpd)
    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": DATA, "config": cfg, "rays_per_step": rays_total / args.steps,
            "detail": {"sharding": f"interleaved {ROW_TILE}-row tiles x{world}" if world > 1 else "none", "exchange": exchange,
                       "exchange_check": exchange_check,
                       "l2": "not flushed: every step streams its own path state through HBM (~1.3 GB per 1080p frame, >> 126 MB L2); the "
                             "2.3 MB scene is L2-resident by design (DESIGN.md section 3)",
                       "frames_in_flight": int(meta.get("lanes", 6)), "traversal": "brick walk (one exact merge step per empty 4x4x4 brick), prepare stage",
                       "timed_region": "one CUDA-event pair around all steps; " + ("the first step runs alone with stage events (VP_FLAG_TIME_STAGES)" if staged_inside else "all steps pipelined (the stage-timed step follows the region)"),
                       "hbm_in_use": meta.get("hbm_in_use"), "graph_node_updates_per_replay": (meta["graph_node_updates"] / meta["graph_replays"]) if meta.get("graph_replays") else None,
                       "also": also},
            "ms_per_frame": total_ms / args.steps,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline}
    print(json.dumps(line), flush=True)
    if share is not None:
        lib.c_share_close(share)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
