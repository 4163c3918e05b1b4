#!/usr/bin/env python3
"""bench.py - headline benchmark of the render hot path (BASELINE.json metric).

A "step" is one Renderer::Update() pass (1 spp) of the workload over the whole frame.
  value     Mrays/s (primary + shadow + secondary rays of all ranks / max-over-ranks device
            time), accumulator resident in HBM, kernels enqueued on torch's current stream and
            timed with CUDA events around exactly K steps.
  e2e       the same metric through the reference-facing C-ABI call vp_render() with HOST
            buffers: float4 accumulator in (pinned) + accumulator and RGB8 out, every step
            (the library cuts such a call into row bands so that copies overlap the trace).
  roofline  dominant kernel (k_wf_extend, the Renderer::FindNearest walk): algorithmic bytes
            (SURVEY.md 8d formula) / its average device time measured by CUDA events inside
            the timed region, against the measured HBM copy bandwidth.
  cpu_baseline  the reference's own CPU Trace (oracle/_ref) timed on this box's host cores.
`--impl reference` times that CPU path alone with the same config/metric.
Multi-GPU (torchrun): the frame is sharded as interleaved 8-row tiles (scene replicated,
no per-sample communication); finished RGB8 tiles are gathered with one NCCL all_gather
per step inside the timed region.  Total work is fixed => "scaling": "strong".
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
vp = importlib.import_module("raytracer-voxpopuli_b200")
scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
sharding = importlib.import_module("raytracer-voxpopuli_b200.sharding")
abi = vp.abi

METRIC = "Mrays/s (primary+shadow+secondary)"
ROW_TILE = sharding.ROW_TILE


def frame_params(fpd, **over):
    d = dict(fpd)
    d.update(over)
    return vp.frame_params(d["width"], d["height"], max_bounces=d["max_bounces"], spp=d.get("spp", 1),
                           frame_index=d.get("frame_index", 0), seed=scenes.SEED, aa_strength=d["aa_strength"],
                           row_begin=d.get("row_begin", 0), row_end=d.get("row_end", 0), row_tile=d.get("row_tile", 0), row_tile_stride=d.get("row_tile_stride", 1),
                           row_tile_phase=d.get("row_tile_phase", 0), flags=d.get("flags", 0))


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
        self.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


def load_checker():
    """The reference itself when its build travelled with the repo, else the plain-C port.  The CPU path gets every core
    this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for its workers)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libvpref.so")
    if os.path.exists(ref):
        lib, kind = abi.Lib(ref, "vpref_"), "reference"
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


def run_reference(args, cfg_idx):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc, cam, fpd = scenes.config(cfg_idx)
    lib, kind, cores = load_checker()
    times, rays = cpu_frame_times(lib, sc, cam, fpd, args.warmup + args.steps)
    t, r = times[args.warmup:], rays[args.warmup:]
    total = sum(t)
    value = sum(r) / total / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "reference assets (room.vox grid fixture), synthetic camera/lights",
            "config": {"workload": scenes.WORKLOAD_NAMES[cfg_idx], "rays_per_step": int(np.mean(r))},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind,
                             "sample": f"{args.steps} full {fpd['width']}x{fpd['height']} frames (1 spp each) after {args.warmup} warm-up"},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="workload (SURVEY.md 8d); 2 = room.vox 1920x1080, the metric's config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    lib = vp.load_product()
    sc, cam, fpd = scenes.config(args.config)
    W, H = fpd["width"], fpd["height"]
    shard = sharding.shard_params(world, rank)
    sess = vp.Session(lib, local)
    sess.set_scene(sc)
    cc = cam.to_c()
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    # frames in flight: a step is enqueued with VP_FLAG_DEFER, so its latency-bound tail overlaps the next step's wide
    # first bounces; with several ranks the finished image of step k is gathered on a second stream meanwhile, hence
    # one RGB8 image per frame in flight, and two more so that a frame never waits for the gather that last read its image
    NBUF = int(os.environ.get("VP_BENCH_IMAGES", "6")) if world > 1 else 1
    rgb8s = [torch.zeros((H, W), dtype=torch.int32, device=dev) for _ in range(NBUF)]
    rgb8 = rgb8s[0]
    stream = torch.cuda.current_stream(dev)
    comm = torch.cuda.Stream(dev) if world > 1 else None
    gathered = [None] * NBUF
    tiles_r = [sharding.tiles_of_rank(H, world, r) for r in range(world)]
    gather = sharding.FrameGather(rgb8, world, rank) if world > 1 else None

    def fence(on, back=0):
        lib.check(lib.c_render_fence(sess.ctx, on.cuda_stream, back), "render_fence")

    lone_ms = [0.0]

    def step(i, flags=abi.VP_FLAG_DEFER):
        img = rgb8s[i % NBUF]
        if gathered[i % NBUF] is not None:
            stream.wait_event(gathered[i % NBUF])  # the gather that last read this image
        fp = frame_params(fpd, frame_index=i, flags=flags, **shard)
        st1 = abi.VpStats() if (flags & abi.VP_FLAG_TIME_STAGES) else None  # a step that runs alone also reports its own device time
        lib.check(lib.c_render_device(sess.ctx, C.byref(cc), C.byref(fp), accum.data_ptr(), img.data_ptr(), stream.cuda_stream,
                                      C.byref(st1) if st1 is not None else None), "render_device")
        if st1 is not None:
            lone_ms[0] += st1.device_ms
        if world > 1:
            with torch.cuda.stream(comm):
                fence(comm)   # this frame resolved
                gather(img)   # pack own tiles -> one NCCL all_gather -> de-interleave: every rank ends with the image
                gathered[i % NBUF] = torch.cuda.Event()
                gathered[i % NBUF].record(comm)

    def drain():
        if world > 1:
            stream.wait_stream(comm)
        fence(stream)

    STAGE_EVERY = 20
    # L2 flush between timed iterations (the frame's working set is far below the 126 MB L2)
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    sampler = ClockSampler(local)
    sampler.start()
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
    for i in range(args.steps):
        flush.fill_(float(i))
        if i % STAGE_EVERY == 0:
            # every 20th step runs alone (pipeline drained, not deferred) with CUDA events around its extend-stage
            # launches: the roofline's kernel duration is measured inside the timed region, without other frames'
            # kernels sharing the SMs
            drain()
            step(args.warmup + i, flags=abi.VP_FLAG_TIME_STAGES)
        else:
            step(args.warmup + i)
    drain()
    t_end.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    total_ms = float(t_begin.elapsed_time(t_end))
    clocks = sampler.finish()
    st = abi.VpStats()
    lib.check(lib.c_stats_end(sess.ctx, C.byref(st)), "stats_end")
    ext_ms, ext_n = C.c_double(), C.c_uint64()
    lib.check(lib.c_stage_times(sess.ctx, C.byref(ext_ms), C.byref(ext_n)), "stage_times")
    t = torch.tensor([total_ms, float(st.rays), float(st.kernel_launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, rays_total, launches = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        rays_total, launches = float(st.rays), float(st.kernel_launches)
    value = rays_total / (total_ms * 1e-3) / 1e6

    # ---- e2e: the reference-facing C-ABI call with host buffers, copies inside the timed region ----
    h_acc = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
    h_rgb = torch.zeros((H, W), dtype=torch.int32).pin_memory()
    e2e_rays, e2e_t = 0, 0.0
    e2e_steps = max(3, min(args.steps, 10))
    for i in range(2 + e2e_steps):
        fp = frame_params(fpd, frame_index=i, **shard)
        s1 = abi.VpStats()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        lib.check(lib.c_render(sess.ctx, C.byref(cc), C.byref(fp), h_acc.data_ptr(), h_rgb.data_ptr(), C.byref(s1)), "render")
        dt = time.perf_counter() - t0
        if i >= 2:
            e2e_rays += s1.rays
            e2e_t += dt
    e = torch.tensor([e2e_t, float(e2e_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        emax, esum = e.clone(), e.clone()
        dist.all_reduce(emax, op=dist.ReduceOp.MAX)
        dist.all_reduce(esum, op=dist.ReduceOp.SUM)
        e2e_t, e2e_rays = float(emax[0]), float(esum[1])
    e2e_value = e2e_rays / e2e_t / 1e6
    own_rows = sum(min(ROW_TILE, H - k * ROW_TILE) for k in tiles_r[rank])
    # vp_render moves only the rows a rank renders; summed over the ranks that is the whole frame
    h2d = W * H * 16
    d2h = W * H * 16 + W * H * 4

    if rank != 0:
        if world > 1:
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
    lib.check(lib.c_stats_begin(sess.ctx), "stats_begin")
    fp = frame_params(fpd, frame_index=args.warmup, flags=abi.VP_FLAG_COUNT_STEPS, **shard)
    lib.check(lib.c_render_device(sess.ctx, C.byref(cc), C.byref(fp), acc2.data_ptr(), rgb8.data_ptr(), stream.cuda_stream, C.byref(inst)),
              "render_device(count steps)")
    info = C.create_string_buffer(4096)
    lib.check(lib.c_info(sess.ctx, info, 4096), "info")
    meta = json.loads(info.value.decode())
    nvol = len(sc.volumes)
    own_rows = sum(min(ROW_TILE, H - k * ROW_TILE) for k in tiles_r[rank])
    px_rank = own_rows * W if world > 1 else W * H
    ext_rays = inst.primary_rays + inst.secondary_rays
    timed_steps = len(range(0, args.steps, STAGE_EVERY))
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
                "kernel": "k_wf_extend", "peak_source": peak_src,
                "launches_per_step": int(ext_launches_per_frame), "kernel_ms_per_step": ext_ms_frame, "steps_with_stage_events": timed_steps,
                "kernel_share_of_step": ext_ms.value / lone_ms[0] if lone_ms[0] > 0 else None,
                "kernel_share_note": "k_wf_extend time / device time of the same steps (the steps that run alone with stage events); "
                                     "against the average pipelined step it is kernel_ms_per_step / ms_per_step",
                "algorithmic_bytes_per_launch": int(ext_bytes_frame / ext_launches_per_frame),
                "algorithmic_bytes_per_step_all_stages": int(frame_bytes),
                "dda_steps_per_step": int(inst.dda_steps), "l2_sector_view": l2,
                "note": "traversal of an L2-resident grid, bound by issue slots / divergence / latency, not by HBM: the HBM "
                        "fraction is a few percent by construction (DESIGN.md section 4)"}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        clib, kind, cores = load_checker()
        reps = 4
        times, rays = cpu_frame_times(clib, sc, cam, fpd, reps)
        best = int(np.argmin(times[1:])) + 1
        cpu_baseline = {"value": rays[best] / times[best] / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
                        "ms_per_frame": 1e3 * times[best],
                        "sample": f"best of {reps - 1} full {W}x{H} frames (1 spp) after 1 warm-up, same scene/camera/seeds"}
        # the 1-thread figure, on a band of rows around the image centre
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
                                          "sample": f"rows {band[0]}..{band[1] - 1} of one {W}x{H} frame, 1 thread"}

    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "reference assets (room.vox grid fixture), synthetic camera/lights",
            "config": {"workload": scenes.WORKLOAD_NAMES[args.config], "resolution": [W, H], "spp_per_step": 1,
                       "rays_per_step": rays_total / args.steps, "sharding": f"interleaved {ROW_TILE}-row tiles x{world}" if world > 1 else "none",
                       "l2": "flushed between timed iterations (160 MiB fill, inside the timed region)",
                       "frames_in_flight": 4, "traversal": "brick walk (one exact merge step per empty 4x4x4 brick), prepare stage",
                       "timed_region": "one CUDA-event pair around all steps incl. the L2 flushes; every "
                       "20th step runs alone with stage events (VP_FLAG_TIME_STAGES)"},
            "ms_per_frame": total_ms / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_t / e2e_steps, "api": "vp_render (host float4 accumulator in/out + RGB8 out)"},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
