#!/usr/bin/env python
"""BASELINE config 3: the 1 M-triangle terrain scene, 1920x1080, 16 spp, 4 bounces, one GPU.

    python tools/mesh_bench.py [--grid 707] [--spp 16] [--steps 5] [--warmup 2] [--json OUT]

Reports ms/frame (CUDA events on the frame stream), Mrays/s, nodes and triangles
per ray, and the algorithmic bytes per ray of SURVEY.md 8d:
nodes visited x 32 B + triangles tested x 48 B.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=707)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--bounces", type=int, default=4)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--json", default=None)
    ap.add_argument("--builder", choices=["sah", "lbvh"], default="sah",
                    help="BVH builder: binned SAH on the host (default) or LBVH on the GPU")
    ap.add_argument("--wide", type=int, default=0, help="1: the kernel walks the 4-wide tree (QUASI_CUDA_MESH_WIDE); 2: both forms, one after the other")
    ap.add_argument("--plugin", default=None, help="path of a libquasi_cuda.so build to measure")
    ap.add_argument("--env", action="append", default=[], help="KEY=VALUE for the plugin's environment (repeatable)")
    args = ap.parse_args()

    import torch

    from quasi_b200.headless import headless_host

    import quasi_b200

    results = []
    for wide in ([0, 1] if args.wide == 2 else [args.wide]):
        results.append(measure(args, torch, headless_host, quasi_b200, wide))
        print(json.dumps(results[-1]), flush=True)
    if args.json:
        Path(args.json).write_text(json.dumps(results[0] if len(results) == 1 else results, indent=1))


def measure(args, torch, headless_host, quasi_b200, wide):
    host = headless_host(args.width, args.height, single_stream=True,
                         plugin_path=args.plugin or quasi_b200.PLUGIN_PATH,
                         env={"QUASI_CUDA_SCENE": "mesh", "QUASI_CUDA_MESH_GRID": args.grid,
                              "QUASI_CUDA_MESH_SEED": 1, "QUASI_CUDA_MAX_BOUNCES": args.bounces,
                              "QUASI_CUDA_MESH_BUILDER": args.builder, "QUASI_CUDA_MESH_WIDE": wide,
                              **dict(kv.split("=", 1) for kv in args.env)})
    stream = host.context.frame_stream()

    def step():
        host.render_frame(camera_dirty=True)
        for _ in range(args.spp - 1):
            host.render_frame(camera_dirty=False)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    s0 = host.stats()
    times = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    s1 = host.stats()
    rays = (s1.rays - s0.rays) / args.steps
    nodes = (s1.nodes_visited - s0.nodes_visited) / args.steps
    tris = (s1.triangles_tested - s0.triangles_tested) / args.steps
    ms = sorted(times)[len(times) // 2]
    out = {
        "workload": f"terrain mesh grid {args.grid} ({12 + 2 * args.grid * args.grid} triangles), "
                    f"{args.width}x{args.height}, {args.spp} spp, {args.bounces} bounces",
        "ms_per_frame_median": round(ms, 3), "ms_per_frame_best": round(min(times), 3),
        "mrays_per_s": round(rays / (ms * 1e-3) / 1e6, 1), "rays_per_frame": rays,
        "nodes_per_ray": round(nodes / rays, 2), "triangles_per_ray": round(tris / rays, 2),
        "algorithmic_bytes_per_ray": round((nodes * 32 + tris * 48) / rays, 1),
        "algorithmic_gbs": round((nodes * 32 + tris * 48) / (ms * 1e-3) / 1e9, 1),
        "wide": wide, "messages": [m for m in host.messages if "esh" in m],
    }
    host.close()
    return out


if __name__ == "__main__":
    main()
