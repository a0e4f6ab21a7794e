#!/usr/bin/env python
"""bench.py - the headline measurement: Mrays/s and ms/frame at 1080p, 64 spp.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload demo|demo4k|mesh] [--scaling strong|weak] [--reduce push|nccl]

Workload (BASELINE.json configs[1]): the demo scene at 1920x1080, 64 samples per
pixel, rendered headless through the unchanged plugin C ABI. One STEP is one
whole frame: 64 Q_plugin_render calls (the first with camera_dirty set, so every
step restarts accumulation and renders RNG frame indices 0..63). A ray is one
closest-hit query (trace_scene, /root/reference/backends/metal/plugin.mm:199-212),
counted on the device.

  value   rays per step / device time per step (CUDA events on the frame stream,
          max over ranks); nothing crosses PCIe inside the timed region.
  e2e     the same frames through the same C ABI, wall clock, including the
          per-call parameter blocks going to the device and the four HDR layers
          (132.7 MB at 1080p) arriving in host memory: Q_plugin_readback_aov on one
          GPU, the shared host frame of quasi_b200.multi_gpu.shard_exchange on several.
  roofline    the dominant kernel (pathtrace_accumulate_kernel) against measured HBM
              bandwidth with the algorithmic 128 B/pixel per render call x the
              calls a launch processes, as the contract asks; the kernel is bound by instruction issue, so `issue` holds
              the figure that explains it: warp instructions per launch (ncu, only when
              the loaded library's SASS is the build that was profiled) over the issue
              slots the launch had (SMs x 4 schedulers x SM clock x kernel time).
  parity  outside every timed region: a 320x180 frame rendered the same way (same shard
          plan, same exchange) compared with the CPU oracle bit for bit.
  cpu_baseline / --impl reference
          the reference's own shader text compiled for the CPU (oracle/_ref),
          or the plain-C port when that is absent, on every host core.

Multi-GPU (launched by torchrun, one rank per GPU): STRONG scaling by default: the same
1080p x 64 spp frame, rank r of N rendering the 64/N RNG frame indices r, r+N, ... and one
exchange per frame (sum over ranks in rank order, 1/N scale, frame into host memory).
--scaling weak gives every rank 64 spp (64 N in total).
--workload demo4k is BASELINE configs[3]: 3840x2160, 1024 spp in total.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tools"))

WORKLOADS = {"demo": (1920, 1080, 64), "demo4k": (3840, 2160, 1024)}
ALGORITHMIC_BYTES_PER_PIXEL_PER_CALL = 128  # 4 layers x (16 B read + 16 B write) per Q_plugin_render call, SURVEY.md 8d
PARITY_SIZE = (320, 180)


def metric_name(workload):
    w, h, spp = WORKLOADS[workload]
    return ("Mrays/s @1080p 64spp (demo scene, headless via plugin C ABI)" if workload == "demo" else
            f"Mrays/s @{w}x{h} {spp}spp (demo scene, sample-sharded, headless via plugin C ABI)")


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class clock_sampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU every 5 ms via NVML."""

    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
        0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons = index, [], set()
        self.max_mhz, self.power, self._stop_evt, self.ok = None, [], threading.Event(), False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        s = sorted(self.samples)
        reasons = sorted(r for r in self.reasons if r != "gpu_idle")
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "power_w_max": round(max(self.power), 1) if self.power else None, "samples": len(s)}


# ---------------------------------------------------------------- CPU arm ----

def cpu_arm(budget_s: float, steps: int, warmup: int, workload: str = "demo"):
    """Times the reference's CPU form of the path on all host cores.

    Returns (mrays_per_s, ms_per_step, descriptor). Each step renders `spp_step` of the frame's
    samples per pixel at the full size (frames 0..spp_step-1), sized so that warmup+steps fit in
    `budget_s`.
    """
    from oracle_lib import Oracle, Reference

    width, height, spp = WORKLOADS[workload]
    oracle = Oracle()
    use_ref = Reference.available()
    impl = Reference() if use_ref else oracle
    cores = impl.cores()
    u = oracle.scene(width, height)

    def run(n_frames):
        t0 = time.perf_counter()
        impl.render_frames(u, width, height, n_frames)
        return time.perf_counter() - t0

    t1 = run(1)  # also pages the library in
    per_step_budget = budget_s / max(1, steps + warmup)
    spp_step = int(max(1, min(spp, per_step_budget // max(t1, 1e-3))))
    # ray count of exactly those frames, from the bit-identical C port, outside the timed region
    _, rays = oracle.render_frames(u, width, height, spp_step)
    for _ in range(warmup):
        run(spp_step)
    t = [run(spp_step) for _ in range(steps)]
    total = sum(t)
    mrays = rays * steps / total / 1e6
    desc = {
        "value": round(mrays, 3), "unit": "Mrays/s", "cores": cores,
        "kind": "reference" if use_ref else "port",
        "sample": f"{spp_step} of {spp} spp per step at {width}x{height} (frames 0..{spp_step - 1}), "
                  f"{steps} timed steps, {rays} rays per step; "
                  + ("reference shader text (oracle/_ref) compiled g++ -O2 -ffp-contract=off"
                     if use_ref else "plain-C port (oracle/quasi_oracle.c) gcc -O2 -ffp-contract=off"),
        "ms_per_full_frame_extrapolated": round(total / steps * 1e3 * spp / spp_step, 1),
    }
    return mrays, total / steps * 1e3, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    workload = args.workload if args.workload in WORKLOADS else "demo"
    width, height, spp = WORKLOADS[workload]
    mrays, ms, desc = cpu_arm(budget_s=args.cpu_budget or 120.0, steps=args.steps, warmup=args.warmup, workload=workload)
    line = {
        "impl": "reference", "metric": metric_name(workload), "value": round(mrays, 3), "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 and args.scaling != "weak" else "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (procedural Cornell box, counter-based RNG)",
        "config": {"workload": f"demo scene {width}x{height}, {spp} spp per frame, CPU, bounded sample per step"},
        "cpu_baseline": desc,
        "e2e": {"value": round(mrays, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- GPU arm ----

def profiled_kernel(name_prefix: str):
    """The ncu figures of a kernel from profiles/ncu_summary.json, but only if the library that is loaded now is,
    instruction for instruction, the build they were captured on (tools/sass_fingerprint.py); else None."""
    summary = ROOT / "profiles" / "ncu_summary.json"
    if not summary.exists():
        return None
    try:
        import sass_fingerprint

        entry = json.loads(summary.read_text()).get(name_prefix)
        if not entry or "sass_md5" not in entry:
            return None
        now = sass_fingerprint.fingerprints()
        if now.get(entry["kernel"], {}).get("md5") != entry["sass_md5"]:
            return None
        return entry
    except Exception:
        return None


def parity_check(args, dist, rank, world, local_rank, spp_rank):
    """Outside every timed region: a 320x180 frame with this run's shard plan, against the CPU oracle."""
    import numpy as np
    from oracle_lib import Oracle, bits
    from quasi_b200.headless import headless_host

    w, h = PARITY_SIZE
    samples = max(1, min(spp_rank, 8))
    host = headless_host(w, h, device_ordinal=local_rank, shard_index=rank, shard_count=world, single_stream=True)

    def render():
        host.render_frame(camera_dirty=True)
        for _ in range(samples - 1):
            host.render_frame(camera_dirty=False)

    out = {"size": f"{w}x{h}", "spp_per_gpu": samples}
    oracle = Oracle() if rank == 0 else None
    if world == 1:
        render()
        got = host.readback_aov()
        want, rays = oracle.render_frames(oracle.scene(w, h), w, h, samples)
        out["bit_exact"] = bool(all(np.array_equal(bits(g), bits(x)) for g, x in zip(got, want)))
        out["rays_equal"] = bool(host.stats().rays == rays)
    else:
        from quasi_b200.multi_gpu import shard_exchange

        want = None
        if rank == 0:
            u = oracle.scene(w, h)
            shards = [oracle.render_frames(u, w, h, samples, frame_first=g, frame_stride=world)[0] for g in range(world)]
            want = [oracle.combine_shards([s[l] for s in shards]) for l in range(4)]
        for mode in ("push", "fused", "nccl"):
            exchange = shard_exchange(host, dist, mode=mode)
            exchange.begin_frame(render_calls=samples)
            render()
            exchange.combine()
            got = exchange.result_numpy()
            exchange.close()
            if rank == 0:
                if mode != "nccl":
                    out[mode + "_bit_exact"] = bool(all(np.array_equal(bits(g), bits(x)) for g, x in zip(got, want)))
                else:
                    out["nccl_max_abs"] = float(max(np.abs(g - x).max() for g, x in zip(got, want)))
    host.close()
    return out


def cpp_host_leg(width, height, spp):
    """The C++ headless host (quasi_b200/host, the delivery vehicle) on the same frame: wall time of a frame set
    from the first Q_plugin_render to the four layers on the host, warm (median of sets 2..)."""
    exe, lib = ROOT / "quasi_b200" / "bin" / "quasi_headless", ROOT / "quasi_b200" / "lib" / "libquasi_cuda.so"
    if not exe.exists():
        return None
    try:
        out = subprocess.run([str(exe), "--render", str(spp), "--repeat", "6", "--readback-each", "--width", str(width),
                              "--height", str(height), "--json", "--out", "/tmp/quasi_bench_host.exr", str(lib)],
                             capture_output=True, text=True, timeout=120)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
        sets = sorted(json.loads(line)["set_ms"][1:])
        return {"ms_per_frame_e2e": round(sets[len(sets) // 2], 3), "sets": len(sets) + 1,
                "what": f"quasi_headless --render {spp} --repeat 6 --readback-each: wall time per {spp}-call frame set "
                        "including the blocking AOV readback, first set (cold) left out"}
    except Exception as e:  # the bench line must not depend on this leg
        return {"error": str(e)[:200]}


def run_ours(args):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libquasi_cuda.so has no CPU fallback")
    import quasi_b200  # noqa: F401
    from quasi_b200.headless import headless_host

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    width, height, spp = WORKLOADS[args.workload]
    strong = world > 1 and args.scaling != "weak"
    if strong and spp % world != 0:
        raise SystemExit(f"bench.py: {spp} spp do not split over {world} GPUs")
    spp_rank = spp // world if strong else spp
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    host = headless_host(width, height, device_ordinal=local_rank, shard_index=rank, shard_count=world,
                         single_stream=True)
    device = host.context.device
    stream = host.context.frame_stream()
    n_px = width * height
    exchange = None
    if world > 1:
        from quasi_b200.multi_gpu import shard_exchange

        exchange = shard_exchange(host, dist, mode=args.reduce)

    def step(to_host=False):
        if exchange is not None:
            exchange.begin_frame(to_host=to_host, render_calls=spp_rank)
        host.render_frame(camera_dirty=True)
        for _ in range(spp_rank - 1):
            host.render_frame(camera_dirty=False)
        if exchange is not None:
            # one exchange per frame: sum over ranks in rank order, 1/N scale; the reduced slices stay on the GPUs
            # in the device-timed loop and go to the shared host frame in the end-to-end loop
            exchange.combine(to_host=to_host)
        else:
            host.plugin.fn("Q_cuda_plugin_flush")(host.plugin._handle, stream.cuda_stream)  # launch what was collected

    def barrier():
        torch.cuda.synchronize(device)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-timed throughput (value) ----
    for _ in range(args.warmup):
        step()
    barrier()
    s0 = host.stats()
    exchange_launches0 = exchange.launches if exchange else 0
    sampler = clock_sampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    s1 = host.stats()
    rays_local = s1.rays - s0.rays
    render_launches_local = s1.kernel_launches - s0.kernel_launches
    launches_local = render_launches_local + (exchange.launches - exchange_launches0 if exchange else 0)

    # ---- end to end through the public API with host buffers (e2e) ----
    def e2e_step():
        step(to_host=True)
        if world == 1:
            rb = host.plugin.readback_aov()          # blocking: the four layers in host memory
            assert rb.buffers[0].data, "readback returned no data"
            host.plugin.readback_aov_free(rb)
        else:
            exchange.finish_frame()                  # blocking: rank 0 returns when every slice is in host memory
            exchange.release_frame()

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    if dist is not None:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0]), float(t[1])
        c = torch.tensor([rays_local, launches_local], dtype=torch.int64, device=device)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        rays_total, launches = int(c[0]), int(c[1])
    else:
        rays_total, launches = rays_local, launches_local

    # ---- parity of this configuration against the oracle, outside every timed region ----
    parity = None
    if not args.no_parity:
        parity = parity_check(args, dist, rank, world, local_rank, spp_rank)

    if rank == 0:
        ms_step = ms_total / args.steps
        rays_step = rays_total / args.steps
        value = rays_step / (ms_step * 1e-3) / 1e6
        e2e_value = rays_step / (e2e_s / args.steps) / 1e6
        peak, peak_src = measured_peaks()
        launches_per_step = max(1.0, render_launches_local / args.steps)
        # the step of one GPU is back-to-back launches of the one kernel (one per frame when the plugin batches)
        # (several GPUs, fused exchange: a step is that launch plus a device-side barrier of a few microseconds, and
        # the launch is the Reduce instantiation of the same kernel, whose tail per finished tile is not in the ncu
        # capture: the ncu-only fields are left out)
        kernel_ms = ms_step / launches_per_step if world == 1 or args.reduce == "fused" else None
        roofline = None
        if kernel_ms is not None:
            # SURVEY.md 8d's unit is one Q_plugin_render call (one sample per pixel: 128 B/px); a launch processes
            # samples_per_launch of them (64 when the plugin batches, 1 otherwise)
            samples_per_launch = spp_rank / launches_per_step
            algorithmic = ALGORITHMIC_BYTES_PER_PIXEL_PER_CALL * n_px * samples_per_launch
            achieved = algorithmic / (kernel_ms * 1e-3) / 1e9
            batched = launches_per_step < spp_rank
            prof = profiled_kernel("pathtrace_accumulate_kernel_batched" if batched else "pathtrace_accumulate_kernel")
            issue = None
            if prof and clocks.get("sm_mhz"):
                per_launch = prof["warp_instructions_per_sample"] * (spp_rank / launches_per_step)
                slots = 148 * 4 * clocks["sm_mhz"] * 1e6 * kernel_ms * 1e-3
                issue = {"warp_inst_per_launch": int(per_launch), "frac": round(per_launch / slots, 4),
                         "warp_inst_per_ray": round(per_launch * launches_per_step / rays_step, 2),
                         "issue_slot_utilization_ncu": prof.get("issue_slot_utilization") if world == 1 else None,
                         "threads_per_instruction_ncu": prof.get("threads_per_instruction") if world == 1 else None,
                         "source": prof.get("source")}
            roofline = {
                "bound": "hbm", "kernel": "pathtrace_accumulate_kernel",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4),
                "traffic": prof.get("dram_bytes_per_launch") if prof and world == 1 else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(algorithmic),
                "samples_per_launch": samples_per_launch,
                "kernel_ms": round(kernel_ms, 5),
                "issue": issue,
                "note": "the contract's HBM figure: SURVEY.md 8d's 128 B/pixel per render call x the calls one launch "
                        "processes. `traffic` (ncu, DRAM bytes of one launch) is far below it for a batched launch because "
                        "a pixel's history stays in L2 between its samples. The kernel is bound by instruction issue, not "
                        "HBM (a 16-quad scan per ray): `issue` is the fraction of the launch's issue slots it fills; "
                        "traffic / issue come from ncu and are null unless the loaded library is the profiled build",
            }
        param_bytes = int(launches_per_step) * (8624 + 80 + 144)  # scene, frame and batch parameter blocks per launch
        line = {
            "metric": metric_name(args.workload), "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (procedural Cornell box, counter-based RNG, no inputs to load)",
            "config": {
                "workload": f"demo scene {width}x{height}, {spp_rank * world} spp per frame"
                            + (f" ({spp_rank} per GPU)" if world > 1 else "") + ", 5 bounces, 4 AOV layers, plugin C ABI",
                "spp_per_gpu": spp_rank, "spp_total": spp_rank * world, "render_calls_per_step": spp_rank,
                "launches_per_step": launches_per_step,
                "parallelism": f"sample-sharded x{world}" + (
                    f", 1 {args.reduce} exchange per frame (rank-ordered sum, 1/N scale, frame into host memory)"
                    if world > 1 else ""),
                "l2": "no flush: a frame streams the 132.7 MB accumulation set (265 MB at 4K x 4 layers), larger than the 126 MB L2",
                "rays_per_step": rays_step,
            },
            "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s",
                    "ms_per_step": round(e2e_s / args.steps * 1e3, 4),
                    "h2d_bytes_per_step": param_bytes, "d2h_bytes_per_step": n_px * 64},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "parity": parity,
        }
        if world > 1:
            line["config"]["limiter"] = ("per frame: the slowest rank's render, then one local reduce kernel per rank and each "
                                         "rank's slice crossing its own PCIe link" if args.reduce == "push" else
                                         "per frame: the slowest rank's render, whose tiles are reduced and sent to the host "
                                         "by the render kernels themselves as their last copy arrives; then a device-side "
                                         "barrier" if args.reduce == "fused" else
                                         "per frame: the slowest rank's render, ncclReduce to rank 0, the scale pass and the "
                                         "whole frame crossing rank 0's PCIe link")
        if world == 1 and not args.no_cpu:
            _, _, desc = cpu_arm(budget_s=args.cpu_budget or 20.0, steps=1, warmup=0, workload=args.workload)
            line["cpu_baseline"] = desc

    if exchange is not None:
        exchange.close()
    host.close()

    if rank == 0:
        if world == 1 and args.workload == "demo" and not args.no_extra:
            line["config"]["cpp_host"] = cpp_host_leg(width, height, spp)
            try:
                line["extra"] = {"mesh": mesh_measurement(steps=3, warmup=3, builder="sah", cpu=False)}
            except Exception as e:
                line["extra"] = {"mesh": {"error": str(e)[:200]}}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------- config 3 (not the headline) ----

def mesh_measurement(steps, warmup, builder, cpu):
    """BASELINE configs[2]: the 1 M-triangle terrain scene (SAH BVH, or LBVH built on the GPU), 1920x1080, 16 spp,
    4 bounces, one GPU. Same conventions as the headline: a step is one 16-spp frame, `value` is device-timed,
    `e2e` includes the readback, the CPU baseline is the mesh oracle on every host core on a bounded sample."""
    import torch

    from quasi_b200.headless import headless_host

    width, height = 1920, 1080
    spp, bounces, grid = 16, 4, 707
    host = headless_host(width, height, single_stream=True,
                         env={"QUASI_CUDA_SCENE": "mesh", "QUASI_CUDA_MESH_GRID": grid, "QUASI_CUDA_MESH_SEED": 1,
                              "QUASI_CUDA_MAX_BOUNCES": bounces, "QUASI_CUDA_MESH_BUILDER": builder})
    stream = host.context.frame_stream()

    def step():
        host.render_frame(camera_dirty=True)
        for _ in range(spp - 1):
            host.render_frame(camera_dirty=False)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    s0 = host.stats()
    sampler = clock_sampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_step = ev0.elapsed_time(ev1) / steps
    s1 = host.stats()
    rays = (s1.rays - s0.rays) / steps
    nodes = (s1.nodes_visited - s0.nodes_visited) / steps
    tris = (s1.triangles_tested - s0.triangles_tested) / steps
    launches = s1.kernel_launches - s0.kernel_launches

    def e2e_step():
        step()
        rb = host.plugin.readback_aov()  # blocking D2H of the four layers into pinned host memory
        assert rb.buffers[0].data, "readback returned no data"
        host.plugin.readback_aov_free(rb)

    for _ in range(min(warmup, 3)):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / steps
    messages = [m for m in host.messages if "esh" in m]
    host.close()

    bytes_step = nodes * 32 + tris * 48  # nodes visited x 32 B + triangles tested x 48 B
    # The frame is 2 * bounces + 2 launches (mesh_wavefront.cu); the trace kernel's four launches are 84 % of it under
    # ncu (profiles/ncu_summary.json), and that share of this run's frame time is what its issue slots are counted over.
    prof = profiled_kernel("mesh_wave_trace_kernel")
    issue = None
    trace_ms = None
    if prof and clocks.get("sm_mhz"):
        share = prof["time_ms_under_ncu_per_frame"] / prof["frame_ms_under_ncu_all_kernels"]
        trace_ms = ms_step * share
        per_frame = prof["warp_instructions_per_frame"]
        slots = 148 * 4 * clocks["sm_mhz"] * 1e6 * trace_ms * 1e-3
        issue = {"warp_inst_per_frame": int(per_frame), "frac": round(per_frame / slots, 4),
                 "warp_inst_per_ray": round(per_frame / rays, 2),
                 "kernel_share_of_frame_ncu": round(share, 4),
                 "issue_slot_utilization_ncu": prof.get("issue_slot_utilization"),
                 "threads_per_instruction_ncu": prof.get("threads_per_instruction"),
                 "threads_per_instruction_per_wave_ncu": prof.get("threads_per_instruction_per_wave"),
                 "l1_hit_rate_ncu": prof.get("l1_hit_rate"), "l2_hit_rate_ncu": prof.get("l2_hit_rate"),
                 "l1_data_pipe_utilization_per_wave_ncu": prof.get("l1_data_pipe_utilization_per_wave"),
                 "dram_bytes_per_frame_ncu": (prof.get("dram_bytes_read", 0) + prof.get("dram_bytes_write", 0)) or None,
                 "source": prof.get("source")}
    line = {
        "metric": "Mrays/s @1080p 16spp (1 M-triangle mesh scene, 4 bounces, headless via plugin C ABI)",
        "value": round(rays / (ms_step * 1e-3) / 1e6, 2), "unit": "Mrays/s", "n_gpus": 1, "steps": steps,
        "warmup": warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (procedural terrain box from pcg_hash, no inputs to load)",
        "config": {"workload": f"terrain mesh grid {grid} ({12 + 2 * grid * grid} triangles), {width}x{height}, {spp} spp, "
                               f"{bounces} bounces, BVH builder {builder}",
                   "rays_per_step": rays, "nodes_per_ray": round(nodes / rays, 2),
                   "triangles_per_ray": round(tris / rays, 2), "plugin_messages": messages,
                   "l2": "no flush: the scene (19 MB of quantized nodes, 96 MB of triangles and shading records) is meant "
                         "to be served by L2; the ray queues of a 16-sample batch (GBs) are streamed through it"},
        "e2e": {"value": round(rays / e2e_s / 1e6, 2), "unit": "Mrays/s", "ms_per_step": round(e2e_s * 1e3, 4),
                "h2d_bytes_per_step": spp * (320 + 80), "d2h_bytes_per_step": width * height * 64},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "issue", "kernel": "mesh_wave_trace_kernel", "issue": issue,
                     "cache_served": {"achieved": round(bytes_step / (ms_step * 1e-3) / 1e9, 1), "unit": "GB/s",
                                      "algorithmic_bytes_per_frame": round(bytes_step),
                                      "what": "nodes visited x 32 B + triangles tested x 48 B (SURVEY.md 8d's unit; the "
                                              "quantized tree reads 16 B per node), served by L1 and L2, NOT HBM traffic"},
                     "kernel_ms_per_frame": round(trace_ms, 4) if trace_ms else None,
                     "launches_per_frame": launches / steps,
                     "note": "config 3 is a wavefront: ray generation, then per bounce a persistent trace kernel and a "
                             "streaming shade kernel, then the resolve (accumulate) pass. The trace kernel is bound by "
                             "instruction issue at 17 to 23 of 32 lanes (DESIGN.md 4.3); the shade and resolve kernels "
                             "stream the ray queues at HBM speed. No HBM fraction is quoted for the trace kernel because "
                             "the bytes above are served by L1/L2"},
    }
    if cpu:
        from oracle_lib import MeshOracle

        cw, ch, cframes = 960, 540, 4
        oracle = MeshOracle(grid)
        oracle.render_frames(cw, ch, 1, max_bounces=bounces)  # pages the library in
        t0 = time.perf_counter()
        _, st = oracle.render_frames(cw, ch, cframes, max_bounces=bounces)
        cpu_s = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": round(st["rays"] / cpu_s / 1e6, 3), "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{cw}x{ch}, {cframes} spp of the same scene and BVH (oracle/quasi_oracle_mesh.c, gcc -O2 "
                      f"-ffp-contract=off, one thread per core), {st['rays']} rays in {cpu_s:.2f} s"}
    return line


def run_mesh(args):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libquasi_cuda.so has no CPU fallback")
    print(json.dumps(mesh_measurement(args.steps, args.warmup, args.mesh_builder, not args.no_cpu)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--reduce", choices=["push", "fused", "nccl"], default="fused",
                    help="multi-GPU exchange: fused = inside the render kernel (tiles staged at the slice owners, counted, "
                         "reduced in rank order by the last arrival and stored into the shared host frame; the default); "
                         "push = tiles pushed by the render kernel + one local reduce kernel per rank (both bit-exact vs "
                         "the oracle); nccl = NCCL reduce to rank 0 + scale kernel + copy to the host")
    ap.add_argument("--scaling", choices=["strong", "weak"], default="strong",
                    help="N > 1: split the frame's samples over the GPUs (default) or give every GPU the whole 64")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the 320x180 comparison with the oracle")
    ap.add_argument("--no-extra", action="store_true", help="skip the C++ host and mesh-scene legs of the default run")
    ap.add_argument("--cpu-budget", type=float, default=0.0, help="seconds of CPU work for the CPU legs")
    ap.add_argument("--workload", choices=["demo", "demo4k", "mesh"], default="demo",
                    help="demo = the headline (BASELINE configs[1]); demo4k = configs[3], 4K x 1024 spp; "
                         "mesh = configs[2], the 1 M-triangle scene, 1 GPU")
    ap.add_argument("--mesh-builder", choices=["sah", "lbvh"], default="sah")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "mesh":
        run_mesh(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
