#!/usr/bin/env python
"""bench.py -- headline benchmark of the FFT-upscale hot path (BASELINE.json: frames/s,
2048x1024 -> 4096x2048 2x upscale, fp32, at 1/2/4/8 B200; HBM GB/s vs peak).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on host cores

One "step" = FRAMES_PER_STEP frames pushed through the whole pipeline (R2C rows -> fused column
FFT/shift/zero-pad/inverse -> C2R rows -> sharpen).  `value` is measured with the frames already
resident in HBM (a ring of distinct device frames so that no iteration can reuse another's lines
in L2); `e2e` goes through the C-ABI with pinned HOST buffers, H2D and D2H inside the timed region.
N > 1: launched by torchrun, one rank per GPU, whole frames sharded across ranks, no collective on
the data path -- torch.distributed is only the barrier and the max-over-ranks.
The reference arm cannot run the Vulkan binary (no Vulkan loader / lavapipe in the image, see
DESIGN.md): it times the CPU oracle port (oracle/, pocketfft + numpy, all host threads).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    # name: (W, H, upscale, precision, sharpen)
    "c1": (256, 128, 2.0, 0, 0.2),
    "c2": (2048, 1024, 2.0, 0, 0.2),
    "c3": (1920, 1080, 2.0, 2, 0.2),
    "c4": (2048, 1024, 2.0, 2, 0.2),
    "c5": (3840, 2160, 2.0, 0, 0.2),
}
METRIC = "frames/s 2048x1024->4096x2048 2x upscale"
UNIT = "frames/s"


def workload_name(cfg):
    w, h, up, prec, s = CONFIGS[cfg]
    return f"{cfg}: {w}x{h}->{int(up * w)}x{int(up * h)} {'fp16' if prec == 2 else 'fp32'} {up:g}x upscale + sharpen {s}"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_algorithmic_bytes(w, h, up_w, up_h, elem):
    """compulsory HBM bytes of each kernel per frame: what it must read once
    plus what it must write once"""
    nx = w // 2 + 1
    b_in, b_s1, b_s2 = 3 * h * w * elem, 3 * h * nx * 8, 3 * up_h * nx * 8
    b_pre = b_out = 3 * up_h * up_w * elem
    return {"r2c_rows": b_in + b_s1, "cols": b_s1 + b_s2, "c2r_rows": b_s2 + b_pre, "sharpen": b_pre + b_out,
            "frame": b_in + b_out}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """summary of the samples taken inside [t0, t1] (the timed region); if the region was too short
        to catch one, of all samples since start() (warm-up + timed region, same load) -- `window` says which"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (t, r) in self.rows if t0 is not None and t0 <= t <= t1 + 0.02]
        window = "timed region" if inside else "warm-up + timed region"
        for r in (inside if inside else [r for (_, r) in self.rows]):
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU arms
def cpu_frames_per_s(cfg, n_frames, workers):
    """the oracle port on host cores (checker timed as the CPU baseline; never the product path)"""
    from oracle import vkresample_oracle as vo
    w, h, up, prec, s = CONFIGS[cfg]
    x = vo.synthetic_frame("noise", w, h)
    if prec == 2:
        x = x.astype(np.float16)
    vo.upscale_frame(x, up, s, prec, dtype=np.float32, workers=workers)  # warm pocketfft plans
    t0 = time.perf_counter()
    for _ in range(n_frames):
        vo.upscale_frame(x, up, s, prec, dtype=np.float32, workers=workers)
    return n_frames / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    from oracle import vkresample_oracle as vo
    w, h, up, prec, s = CONFIGS[args.config]
    x = vo.synthetic_frame("noise", w, h)
    if prec == 2:
        x = x.astype(np.float16)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        time.sleep(3.0)   # torchrun: let the idle ranks finish importing and exit before the host cores are timed
    for _ in range(min(args.warmup, 2)):
        vo.upscale_frame(x, up, s, prec, dtype=np.float32, workers=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vo.upscale_frame(x, up, s, prec, dtype=np.float32, workers=cores)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = f"{args.steps} steps x 1 frame of {workload_name(args.config)} (oracle port: pocketfft complex64 + numpy sharpen)"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.config), "frames_per_step": 1,
                       "note": "reference Vulkan binary not runnable here (no Vulkan loader/lavapipe); CPU oracle port timed instead"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(gpu_index):
    """pin this rank to the CPUs NVML reports as local to its GPU before any pinned allocation (first-touch
    puts the staging buffers on that node); returns a short description for the JSON line"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f"{len(cpus)} cpus [{cpus[0]}..{cpus[-1]}] (nvmlDeviceSetCpuAffinity)"
    except Exception as e:  # noqa: BLE001
        return f"unchanged ({type(e).__name__})"


def run_b200(args):
    import torch
    import vkresample_b200 as vb
    from vkresample_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = bind_to_gpu_numa_node(local) if (world > 1 and not args.no_affinity) else "unchanged (single rank)"
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available() or vb.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback): " + vb.load_library().b2r_last_error().decode())
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    w, h, up, prec, s = CONFIGS[args.config]
    plan_flags = vb.FLAG_EXACT_SHARPEN if args.exact_sharpen else 0
    plan = vb.Plan(w, h, up, prec, s, device=local, flags=plan_flags)
    plan.set_lanes(args.lanes)
    elem = 2 if prec == 2 else 4
    np_dt = np.float16 if prec == 2 else np.float32

    # ring of distinct device-resident frames: 2x the L2 (126 MB) worth of inputs + their outputs
    # a multiple of the lane count: frames that share a buffer then run on the same lane (stream order),
    # so no two frames in flight ever touch the same output buffer
    ring = max(2, min(args.ring, 16))
    ring = -(-ring // args.lanes) * args.lanes
    rng = np.random.default_rng(1234 + rank)
    d_in, d_out = [], []
    for i in range(ring):
        x = rng.random((3, h, w), dtype=np.float32).astype(np_dt)
        host = plan.pack_input(x)
        d_in.append(torch.from_numpy(host.view(np.uint8)).to(dev))
        d_out.append(torch.empty(plan.output_bytes, dtype=torch.uint8, device=dev))
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t_ = torch.tensor([v], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    def run_frames(p, n, i0=0):
        for f in range(n):
            k = (i0 + f) % ring
            p.enqueue_device(d_in[k].data_ptr(), d_out[k].data_ptr())

    # ---- calibration (untimed): frames per step such that the timed region lasts >= --min-seconds
    run_frames(plan, 4 * ring)
    plan.synchronize()
    plan.timer_start(); run_frames(plan, 8 * ring); t_cal = plan.timer_stop() * 1e-3 / (8 * ring)
    t_cal = allmax(t_cal)
    F = args.frames_per_step
    if F <= 0:
        F = int(np.ceil(args.min_seconds / (args.steps * t_cal)))
        F = max(ring, -(-F // ring) * ring)
    # the stream of one step, sharded like the reference's file loop (VkResample.cpp:1622-1629):
    # global frame f (1-based) -> rank (f-1) mod world; every rank owns exactly F of the world*F frames
    mine = sharding.frames_for_worker(world * F, world, rank)
    assert len(mine) == F and all((f - 1) % world == rank for f in mine)

    sampler = ClockSampler(local)
    sampler.start()                  # runs through warm-up and the timed region; summarised per window below
    time.sleep(0.05)
    for i in range(args.warmup):
        run_frames(plan, F, i * F)
    plan.synchronize()

    launches0 = plan.launch_count
    barrier()
    t_wall0 = time.perf_counter()
    plan.timer_start()
    for i in range(args.steps):
        run_frames(plan, F, i * F)
    ms = plan.timer_stop()          # CUDA events on the launching stream
    t_wall1 = time.perf_counter()
    t_wall = t_wall1 - t_wall0
    clocks = sampler.stop(t_wall0, t_wall1)
    barrier()
    launches = plan.launch_count - launches0
    ms_max = allmax(ms)
    value = sharding.aggregate_frames_per_s(args.steps * F, world, ms_max * 1e-3)

    # ---- burst figure (round 1's timed region: 20 x 32 frames, ~0.08 s) beside the sustained one
    barrier()
    plan.timer_start(); run_frames(plan, 640); ms_b = allmax(plan.timer_stop())
    burst = {"value": world * 640 / (ms_b * 1e-3), "unit": UNIT, "frames": 640, "timed_region_s": ms_b * 1e-3}

    # ---- per-kernel device time for the roofline (events between kernels, same plan and workload, directly after
    # the timed region: later legs create further plans and pinned buffers)
    pk = plan.profile_kernels(20)

    # ---- the same device-resident measurement with B2R_FLAG_EXACT_SHARPEN (bit-exact sharpen kernels);
    # reported beside the headline, never instead of it
    exact = None
    if not args.exact_sharpen and not args.no_exact_leg:
        with vb.Plan(w, h, up, prec, s, device=local, flags=vb.FLAG_EXACT_SHARPEN) as px:
            px.set_lanes(args.lanes)
            n_x = max(ring, min(F * args.steps, 2048))
            run_frames(px, 3 * ring)
            px.synchronize()
            barrier()
            px.timer_start(); run_frames(px, n_x); ms_x = allmax(px.timer_stop())
            barrier()
            t_dt = torch.float16 if prec == 2 else torch.float32
            scratch = torch.empty(plan.output_bytes, dtype=torch.uint8, device=dev)
            px.enqueue_device(d_in[0].data_ptr(), scratch.data_ptr()); px.synchronize()
            plan.enqueue_device(d_in[0].data_ptr(), d_out[0].data_ptr()); plan.synchronize()
            diff = float((scratch.view(t_dt).float() - d_out[0].view(t_dt).float()).abs().max().item())
            pkx = px.profile_kernels(10)
            exact = {"value": world * n_x / (ms_x * 1e-3), "unit": UNIT, "frames": n_x,
                     "flag": "B2R_FLAG_EXACT_SHARPEN", "sharpen_us": round(pkx["sharpen"] * 1e3, 2),
                     "max_abs_default_vs_exact_output": diff}
            del scratch

    # ---- when the plan runs the fused C2R + sharpen kernel: the same workload with the two kernels kept apart
    # (B2R_FLAG_SEPARATE_SHARPEN), for the per-kernel roofline of the pieces the fused kernel replaced
    separate = None
    if int(plan.info.fused_strips_per_plane) > 0 and not args.no_exact_leg:
        with vb.Plan(w, h, up, prec, s, device=local, flags=plan_flags | vb.FLAG_SEPARATE_SHARPEN) as ps:
            ps.set_lanes(args.lanes)
            n_s = max(ring, min(F * args.steps, 2048))
            run_frames(ps, 3 * ring)
            ps.synchronize()
            barrier()
            ps.timer_start(); run_frames(ps, n_s); ms_s = allmax(ps.timer_stop())
            barrier()
            pks = ps.profile_kernels(20)
            separate = {"value": world * n_s / (ms_s * 1e-3), "unit": UNIT, "frames": n_s, "flag": "B2R_FLAG_SEPARATE_SHARPEN",
                        "kernel_us": {k: round(v * 1e3, 2) for k, v in pks.items()}}

    # ---- end to end through the C-ABI with pinned HOST buffers: per frame H2D + frame + D2H, frames
    # rotating over the plan's lanes so that the copies of one frame overlap the kernels of another.
    # No drain between steps: buffer k is always used by lane k mod lanes (n_host is a multiple of the lane
    # count), so stream order alone keeps frames that share a buffer apart; one synchronize ends the region.
    n_host = 2 * args.lanes
    h_in = [torch.from_numpy(plan.pack_input(rng.random((3, h, w), dtype=np.float32).astype(np_dt)).view(np.uint8)).pin_memory()
            for _ in range(n_host)]
    h_out = [torch.empty(plan.output_bytes, dtype=torch.uint8).pin_memory() for _ in range(n_host)]
    u_in = [torch.from_numpy(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).pin_memory() for _ in range(n_host)]
    u_out = [torch.empty((plan.up_h, plan.up_w, 3), dtype=torch.uint8).pin_memory() for _ in range(n_host)]
    e_steps = max(1, min(args.steps, 5))

    def timed_host_leg(enq, bufs_in, bufs_out, min_s):
        for i in range(n_host):
            enq(bufs_in[i].data_ptr(), bufs_out[i].data_ptr())
        plan.synchronize()
        t0 = time.perf_counter()
        for i in range(2 * n_host):
            enq(bufs_in[i % n_host].data_ptr(), bufs_out[i % n_host].data_ptr())
        plan.synchronize()
        per = allmax((time.perf_counter() - t0) / (2 * n_host))   # calibration: slowest rank's seconds per frame
        frames = max(n_host, int(np.ceil(min_s / (e_steps * per))))
        frames = -(-frames // n_host) * n_host
        barrier()
        t0 = time.perf_counter()
        for st in range(e_steps):
            for i in range(frames):
                k = (st * frames + i) % n_host
                enq(bufs_in[k].data_ptr(), bufs_out[k].data_ptr())
        plan.synchronize()
        dt = allmax(time.perf_counter() - t0)
        return world * e_steps * frames / dt, frames, dt

    def copy_ceiling(bufs_in, bufs_out, nbytes_in, nbytes_out, frames):
        """bare pinned H2D + D2H of the same bytes per frame on `lanes` streams, no kernels"""
        streams = [torch.cuda.Stream(device=dev) for _ in range(args.lanes)]
        dd_in = [torch.empty(nbytes_in, dtype=torch.uint8, device=dev) for _ in range(args.lanes)]
        dd_out = [torch.empty(nbytes_out, dtype=torch.uint8, device=dev) for _ in range(args.lanes)]
        flat_in = [b.view(-1) if b.dtype == torch.uint8 else b.view(torch.uint8).view(-1) for b in bufs_in]
        flat_out = [b.view(-1) if b.dtype == torch.uint8 else b.view(torch.uint8).view(-1) for b in bufs_out]

        def go(n):
            for i in range(n):
                k, l = i % n_host, i % args.lanes
                with torch.cuda.stream(streams[l]):
                    dd_in[l].copy_(flat_in[k], non_blocking=True)
                    flat_out[k].copy_(dd_out[l], non_blocking=True)
            for st_ in streams:
                st_.synchronize()
        go(n_host)
        barrier()
        t0 = time.perf_counter()
        go(e_steps * frames)
        dt = allmax(time.perf_counter() - t0)
        return world * e_steps * frames / dt

    e2e_value, e_frames, e_dt = timed_host_leg(plan.enqueue_host, h_in, h_out, args.min_seconds_e2e)
    e2e_ceiling = copy_ceiling(h_in, h_out, plan.input_bytes, plan.output_bytes, e_frames)
    # ---- the same through the byte-pixel extension (u8 RGB in -> u8 RGB out, conversions on the GPU)
    e2e_u8_value, u_frames, u_dt = timed_host_leg(plan.enqueue_host_u8, u_in, u_out, args.min_seconds_e2e)
    e2e_u8_ceiling = copy_ceiling(u_in, u_out, 3 * w * h, 3 * plan.up_w * plan.up_h, u_frames)

    result_checksum = float(np.frombuffer(h_out[0].numpy().tobytes()[:4096], dtype=np_dt).astype(np.float64).sum())

    # ---- roofline of the dominant kernel
    alg = kernel_algorithmic_bytes(w, h, plan.up_w, plan.up_h, elem)
    if int(plan.info.fused_strips_per_plane) > 0:   # fused C2R + sharpen (+ boundary-row fix-up): no pre-sharpen plane
        pk = {"r2c_rows": pk["r2c_rows"], "cols": pk["cols"], "c2r_sharpen": pk["c2r_rows"]}
        alg["c2r_sharpen"] = alg["c2r_rows"]   # S2 in + output once (the pre-sharpen plane never reaches HBM)
    dom = max(pk, key=pk.get)
    peak, peak_src = peaks()
    achieved = alg[dom] / (pk[dom] * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tj.get(args.config, {}).get(dom)
    except Exception:
        pass
    # the fused kernel is issue-bound, not HBM-bound: its instruction count (ncu, profiles/r2_ncu_c2.md) against
    # the SM issue rate is reported beside the byte roofline
    issue = None
    if dom == "c2r_sharpen" and args.config == "c2":
        winst = 45.76e6
        t_issue = winst / (148 * 4 * 1.965e9)
        issue = {"warp_instructions_per_launch": winst, "issue_limit_us": round(t_issue * 1e6, 1),
                 "frac_of_issue_limit": round(t_issue / (pk[dom] * 1e-3), 3), "source": "ncu smsp__inst_executed.sum, profiles/r2_ncu_c2.md"}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom],
                "kernel_us": {k: round(v * 1e3, 2) for k, v in pk.items()},
                "kernel_frac": {k: round(alg[k] / (v * 1e-3) / 1e9 / peak, 4) for k, v in pk.items()},
                "frame_algorithmic_gbs": alg["frame"] * value / world / 1e9,
                "frame_frac": alg["frame"] * value / world / 1e9 / peak}
    if issue:
        roofline["issue"] = issue

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n = 4 if args.config != "c5" else 2
            v = cpu_frames_per_s(args.config, n, cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n} frames of {workload_name(args.config)} (oracle port: pocketfft complex64 + numpy sharpen, {cores} threads)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if prec == 0 else "f16 storage / f32 FFT",
                "data": "synthetic",
                "config": {"workload": workload_name(args.config), "frames_per_step": F, "frames_per_gpu_per_step": F,
                           "timed_region_s": ms_max * 1e-3,
                           "parallelism": f"frames sharded over {world} GPU(s) (frame f -> rank (f-1) mod {world}), no collective",
                           "l2": f"ring of {ring} distinct device-resident frames ({ring * (plan.input_bytes + plan.output_bytes) >> 20} MiB in+out, "
                                 f"~{(alg['r2c_rows'] + alg['cols'] + alg['c2r_rows'] + alg['sharpen']) >> 20} MiB touched per frame) > 126 MB L2",
                           "lanes": int(plan.lanes),
                           "radix_schedule": plan.radix_schedule(), "column_tile": int(plan.info.column_tile),
                           "static_kernels": int(plan.info.static_kernels),
                           "kernels_per_frame": int(plan.info.kernels_per_frame),
                           "fused_c2r_sharpen_strips_per_plane": int(plan.info.fused_strips_per_plane),
                           "cpu_affinity": affinity,
                           "sharpen_arithmetic": "correctly rounded, bit-identical to the oracle (B2R_FLAG_EXACT_SHARPEN)" if args.exact_sharpen
                                                 else "tolerance-bound (default): within 1e-5 fp32 / 1e-2 fp16 of oracle.sharpen on the same plane"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e_frames * plan.input_bytes,
                        "d2h_bytes_per_step": e_frames * plan.output_bytes, "frames_per_step": e_frames,
                        "steps": e_steps, "timed_region_s": e_dt,
                        "api": "b2r_enqueue_host + b2r_synchronize (pinned host in -> pinned host out)",
                        "copy_ceiling": e2e_ceiling, "frac_of_copy_ceiling": e2e_value / e2e_ceiling,
                        "copy_ceiling_gb_s": e2e_ceiling * (plan.input_bytes + plan.output_bytes) / 1e9,
                        "checksum": result_checksum},
                "e2e_u8": {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": u_frames * 3 * w * h,
                           "d2h_bytes_per_step": u_frames * 3 * plan.up_w * plan.up_h, "frames_per_step": u_frames,
                           "steps": e_steps, "timed_region_s": u_dt,
                           "api": "b2r_enqueue_host_u8 (u8 RGB in -> u8 RGB out, /255 fill and truncating "
                                  "quantiser of launchResample run on the GPU)",
                           "copy_ceiling": e2e_u8_ceiling, "frac_of_copy_ceiling": e2e_u8_value / e2e_u8_ceiling,
                           "copy_ceiling_gb_s": e2e_u8_ceiling * 3 * (w * h + plan.up_w * plan.up_h) / 1e9,
                           "checksum": int(u_out[0].numpy()[:8, :8].astype(np.int64).sum())},
                "copy_ceiling_note": "bare pinned cudaMemcpyAsync H2D + D2H of the same bytes per frame on the same number of "
                                     "streams, no kernels (torch copy_ on torch streams), all ranks at once, max over ranks",
                "gpu_launches": int(launches), "wall_s_timed_region": t_wall,
                "burst": burst, "roofline": roofline, "clocks": clocks}
        if exact:
            line["exact_sharpen"] = exact
        if separate:
            separate["kernel_frac"] = {k: round(alg[k] / (v * 1e-6) / 1e9 / peak, 4) for k, v in separate["kernel_us"].items()}
            line["separate_kernels"] = separate
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    plan.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--frames-per-step", type=int, default=0,
                    help="frames per step per GPU; 0 (default): calibrated so that the timed region lasts >= --min-seconds")
    ap.add_argument("--min-seconds", type=float, default=2.8, help="length of the device-resident timed region")
    ap.add_argument("--min-seconds-e2e", type=float, default=1.2, help="length of each host-fed timed region")
    ap.add_argument("--no-affinity", action="store_true", help="N > 1: do not bind ranks to their GPU's NUMA-local CPUs")
    ap.add_argument("--ring", type=int, default=8)
    ap.add_argument("--lanes", type=int, default=3,
                    help="concurrent frames in flight per GPU (like the reference's -numthreads on one device)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exact-leg", action="store_true", help="skip the extra B2R_FLAG_EXACT_SHARPEN measurement (profiling runs)")
    ap.add_argument("--no-fast-leg", action="store_true", help="(round-1 name) same as --no-exact-leg")
    ap.add_argument("--exact-sharpen", action="store_true",
                    help="create the plan with B2R_FLAG_EXACT_SHARPEN (bit-exact sharpen; not the default; recorded in config)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    args.no_exact_leg = args.no_exact_leg or args.no_fast_leg
    # stdout carries exactly ONE JSON line: everything libraries print to fd 1 while we run (e.g. NCCL's
    # "NCCL version ..." banner under torchrun) is sent to stderr, the result line goes to the real stdout
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


_RESULT_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


if __name__ == "__main__":
    sys.exit(main())
