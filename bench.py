#!/usr/bin/env python
"""bench.py - the headline measurement: Mrays/s and ms/frame at 1080p, 64 spp.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the demo scene at 1920x1080, 64 samples per
pixel, rendered headless through the unchanged plugin C ABI. One STEP is one
whole frame: 64 Q_plugin_render calls (the first with camera_dirty set, so every
step restarts accumulation and renders RNG frame indices 0..63), then, with
N > 1 GPUs, one sum-reduce of the accumulation buffers to rank 0 and the 1/N
scale. A ray is one closest-hit query (trace_scene,
/root/reference/backends/metal/plugin.mm:199-212), counted on the device.

  value   rays per step / device time per step (CUDA events on the frame stream,
          max over ranks); nothing crosses PCIe inside the timed region.
  e2e     the same frames through the same C ABI, wall clock, including the
          per-call parameter blocks going to the device and Q_plugin_readback_aov
          bringing the four HDR layers (132.7 MB) back into host memory.
  roofline    the dominant kernel (pathtrace_accumulate_kernel) against measured
              HBM bandwidth with the algorithmic 128 B/pixel/call.
              The kernel is FP32-issue-bound, not HBM-bound: see DESIGN.md.
  cpu_baseline / --impl reference
          the reference's own shader text compiled for the CPU (oracle/_ref),
          or the plain-C port when that is absent, on every host core.

Multi-GPU (launched by torchrun, one rank per GPU): weak scaling. Rank r renders
RNG frame indices r, r+N, r+2N, ... (64 per rank, 64*N spp in total).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WIDTH, HEIGHT, SPP = 1920, 1080, 64
METRIC = "Mrays/s @1080p 64spp (demo scene, headless via plugin C ABI)"
ALGORITHMIC_BYTES_PER_PIXEL_PER_CALL = 128  # 4 layers x (16 B read + 16 B write), SURVEY.md 8d


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class clock_sampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU every 20 ms via NVML."""

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
            self._stop_evt.wait(0.02)

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

def cpu_arm(budget_s: float, steps: int, warmup: int):
    """Times the reference's CPU form of the path on all host cores.

    Returns (mrays_per_s, ms_per_step, descriptor). Each step renders
    `spp_step` of the 64 samples per pixel at the full 1920x1080 (frames
    0..spp_step-1), sized so that warmup+steps fit in `budget_s`.
    """
    from oracle_lib import Oracle, Reference

    oracle = Oracle()
    use_ref = Reference.available()
    impl = Reference() if use_ref else oracle
    cores = impl.cores()
    u = oracle.scene(WIDTH, HEIGHT)

    def run(n_frames):
        t0 = time.perf_counter()
        impl.render_frames(u, WIDTH, HEIGHT, n_frames)
        return time.perf_counter() - t0

    t1 = run(1)  # also pages the library in
    per_step_budget = budget_s / max(1, steps + warmup)
    spp_step = int(max(1, min(SPP, per_step_budget // max(t1, 1e-3))))
    # ray count of exactly those frames, from the bit-identical C port, outside the timed region
    _, rays = oracle.render_frames(u, WIDTH, HEIGHT, spp_step)
    for _ in range(warmup):
        run(spp_step)
    t = [run(spp_step) for _ in range(steps)]
    total = sum(t)
    mrays = rays * steps / total / 1e6
    desc = {
        "value": round(mrays, 3), "unit": "Mrays/s", "cores": cores,
        "kind": "reference" if use_ref else "port",
        "sample": f"{spp_step} of {SPP} spp per step at {WIDTH}x{HEIGHT} (frames 0..{spp_step - 1}), "
                  f"{steps} timed steps, {rays} rays per step; "
                  + ("reference shader text (oracle/_ref) compiled g++ -O2 -ffp-contract=off"
                     if use_ref else "plain-C port (oracle/quasi_oracle.c) gcc -O2 -ffp-contract=off"),
        "ms_per_full_frame_extrapolated": round(total / steps * 1e3 * SPP / spp_step, 1),
    }
    return mrays, total / steps * 1e3, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    mrays, ms, desc = cpu_arm(budget_s=args.cpu_budget or 120.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mrays, 3), "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (procedural Cornell box, counter-based RNG)",
        "config": {"workload": f"demo scene {WIDTH}x{HEIGHT}, {SPP} spp per frame, CPU, bounded sample per step"},
        "cpu_baseline": desc,
        "e2e": {"value": round(mrays, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- GPU arm ----

def run_ours(args):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libquasi_cuda.so has no CPU fallback")
    import quasi_b200
    from quasi_b200.headless import headless_host

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    host = headless_host(WIDTH, HEIGHT, device_ordinal=local_rank, shard_index=rank, shard_count=world,
                         single_stream=True)
    device = host.context.device
    stream = host.context.frame_stream()
    n_px = WIDTH * HEIGHT
    reducer = None
    if world > 1:
        from quasi_b200.multi_gpu import shard_reducer

        reducer = shard_reducer(host, dist, mode=args.reduce)
        pinned = torch.empty(n_px * 16, dtype=torch.float32).pin_memory() if rank == 0 else None

    def render_frame_set():
        host.render_frame(camera_dirty=True)
        for _ in range(SPP - 1):
            host.render_frame(camera_dirty=False)

    def step():
        render_frame_set()
        if reducer is not None:
            reducer.combine()  # one exchange per frame: sum over ranks, 1/N scale, result on rank 0

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
    reducer_launches0 = reducer.launches if reducer else 0
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
    launches_local = render_launches_local + (reducer.launches - reducer_launches0 if reducer else 0)

    # ---- end to end through the public API with host buffers (e2e) ----
    def e2e_step():
        step()
        if world == 1:
            rb = host.plugin.readback_aov()          # blocking D2H of the four layers
            assert rb.buffers[0].data, "readback returned no data"
            host.plugin.readback_aov_free(rb)
        elif rank == 0:
            from quasi_b200.multi_gpu import device_tensor

            with torch.cuda.stream(stream):          # the combined layers, D2H into pinned memory
                pinned.copy_(device_tensor(reducer.result_ptr, n_px * 16, device), non_blocking=True)
            stream.synchronize()

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

    if rank == 0:
        ms_step = ms_total / args.steps
        rays_step = rays_total / args.steps
        value = rays_step / (ms_step * 1e-3) / 1e6
        e2e_value = rays_step / (e2e_s / args.steps) / 1e6
        peak, peak_src = measured_peaks()
        # the step is back-to-back launches of the one kernel: SPP of them, or SPP / k under QUASI_CUDA_BATCH=k
        launches_per_step = max(1.0, render_launches_local / args.steps)
        kernel_ms = ms_step / launches_per_step
        achieved = ALGORITHMIC_BYTES_PER_PIXEL_PER_CALL * n_px / (kernel_ms * 1e-3) / 1e9
        traffic, issue_util = None, None
        summary = ROOT / "profiles" / "ncu_summary.json"
        if summary.exists():
            try:
                k = json.loads(summary.read_text()).get("pathtrace_accumulate_kernel", {})
                traffic, issue_util = k.get("dram_bytes_per_launch"), k.get("issue_slot_utilization")
            except Exception:
                traffic = None
        param_bytes = int(launches_per_step) * (7856 + 80)  # sizeof(scene_params) + sizeof(frame_params) per launch
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (procedural Cornell box, counter-based RNG, no inputs to load)",
            "config": {
                "workload": f"demo scene {WIDTH}x{HEIGHT}, {SPP} spp per GPU per frame "
                            f"({SPP * world} spp total), 5 bounces, 4 AOV layers, plugin C ABI",
                "spp_per_gpu": SPP, "spp_total": SPP * world, "render_calls_per_step": SPP,
                "launches_per_step": launches_per_step,
                "parallelism": f"sample-sharded x{world}" + (
                    f", 1 {args.reduce} combine/frame (sum over ranks, 1/N scale, result on rank 0)"
                    if world > 1 else ""),
                "l2": "no flush: each launch streams the 132.7 MB accumulation set, larger than the 126 MB L2",
                "rays_per_step": rays_step,
            },
            "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s",
                    "ms_per_step": round(e2e_s / args.steps * 1e3, 4),
                    "h2d_bytes_per_step": param_bytes, "d2h_bytes_per_step": n_px * 64},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "bound": "hbm", "kernel": "pathtrace_accumulate_kernel",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALGORITHMIC_BYTES_PER_PIXEL_PER_CALL * n_px,
                "kernel_ms": round(kernel_ms, 5),
                "issue_slot_utilization_ncu": issue_util,
                "note": "instruction-issue-bound kernel (16-quad scan per ray, ~36 warp instructions per ray): the "
                        "HBM fraction is low by design, the issue slots are the resource that is full "
                        "(ncu, profiles/ncu_summary.json); see DESIGN.md section 4.1",
            },
        }
        if world == 1 and not args.no_cpu:
            _, _, desc = cpu_arm(budget_s=args.cpu_budget or 20.0, steps=1, warmup=0)
            line["cpu_baseline"] = desc
        print(json.dumps(line), flush=True)

    if reducer is not None:
        dist.barrier()
        if rank != 0:
            reducer.close()   # unmap peers before their owners free
        dist.barrier()
        if rank == 0:
            reducer.close()
    host.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------- config 3 (not the headline) ----

def run_mesh(args):
    """BASELINE configs[2]: the 1 M-triangle terrain scene (SAH BVH, or LBVH built on the GPU with
    --mesh-builder lbvh), 1920x1080, 16 spp, 4 bounces, one GPU. Same conventions as the headline:
    a step is one 16-spp frame, `value` is device-timed, `e2e` includes the readback, the CPU
    baseline is the mesh oracle on every host core on a bounded sample."""
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libquasi_cuda.so has no CPU fallback")
    from quasi_b200.headless import headless_host

    spp, bounces, grid = 16, 4, 707
    host = headless_host(WIDTH, HEIGHT, single_stream=True,
                         env={"QUASI_CUDA_SCENE": "mesh", "QUASI_CUDA_MESH_GRID": grid, "QUASI_CUDA_MESH_SEED": 1,
                              "QUASI_CUDA_MAX_BOUNCES": bounces, "QUASI_CUDA_MESH_BUILDER": args.mesh_builder})
    stream = host.context.frame_stream()

    def step():
        host.render_frame(camera_dirty=True)
        for _ in range(spp - 1):
            host.render_frame(camera_dirty=False)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    s0 = host.stats()
    sampler = clock_sampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_step = ev0.elapsed_time(ev1) / args.steps
    s1 = host.stats()
    rays = (s1.rays - s0.rays) / args.steps
    nodes = (s1.nodes_visited - s0.nodes_visited) / args.steps
    tris = (s1.triangles_tested - s0.triangles_tested) / args.steps
    launches = s1.kernel_launches - s0.kernel_launches

    def e2e_step():
        step()
        rb = host.plugin.readback_aov()  # blocking D2H of the four layers into pinned host memory
        assert rb.buffers[0].data, "readback returned no data"
        host.plugin.readback_aov_free(rb)

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / args.steps
    messages = [m for m in host.messages if "esh" in m]
    host.close()

    peak, peak_src = measured_peaks()
    bytes_step = nodes * 32 + tris * 48  # nodes visited x 32 B + triangles tested x 48 B
    line = {
        "metric": "Mrays/s @1080p 16spp (1 M-triangle mesh scene, 4 bounces, headless via plugin C ABI)",
        "value": round(rays / (ms_step * 1e-3) / 1e6, 2), "unit": "Mrays/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (procedural terrain box from pcg_hash, no inputs to load)",
        "config": {"workload": f"terrain mesh grid {grid} ({12 + 2 * grid * grid} triangles), {WIDTH}x{HEIGHT}, {spp} spp, "
                               f"{bounces} bounces, BVH builder {args.mesh_builder}",
                   "rays_per_step": rays, "nodes_per_ray": round(nodes / rays, 2),
                   "triangles_per_ray": round(tris / rays, 2), "plugin_messages": messages,
                   "l2": "no flush: the scene (20-37 MB of nodes, 96 MB of triangles and shading records) is meant "
                         "to be served by L2; the 132.7 MB accumulation set is streamed every launch"},
        "e2e": {"value": round(rays / e2e_s / 1e6, 2), "unit": "Mrays/s", "ms_per_step": round(e2e_s * 1e3, 4),
                "h2d_bytes_per_step": spp * (320 + 80), "d2h_bytes_per_step": WIDTH * HEIGHT * 64},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "mesh_pathtrace_kernel",
                     "achieved": round(bytes_step / (ms_step * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(bytes_step / (ms_step * 1e-3) / 1e9 / peak, 4), "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": round(bytes_step / spp),
                     "kernel_ms": round(ms_step / spp, 5),
                     "note": "algorithmic bytes (nodes x 32 B + triangles x 48 B) are served by L1/L2, not HBM "
                             "(DRAM traffic is ~0.8 GB per launch, profiles/ncu_summary.json): the kernel is "
                             "bound by instruction issue at 14.6 of 32 lanes, see DESIGN.md section 4.3"},
    }
    if not args.no_cpu:
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
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--reduce", choices=["p2p", "p2p_host", "nccl"], default="nccl",
                    help="multi-GPU combine: NCCL reduce + scale kernel (default: 0.25 ms per frame faster at 8 "
                         "GPUs) or the fused peer-memory kernel (p2p: fixed summation order, bit-exact vs the oracle)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--cpu-budget", type=float, default=0.0, help="seconds of CPU work for the CPU legs")
    ap.add_argument("--workload", choices=["demo", "mesh"], default="demo",
                    help="demo = the headline (BASELINE configs[1]); mesh = configs[2], the 1 M-triangle scene, 1 GPU")
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
