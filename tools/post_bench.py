#!/usr/bin/env python
"""tools/post_bench.py - the HBM-bound passes against measured HBM bandwidth.

Times, through the C ABI extension entry points of libquasi_cuda.so, the stand-alone forms of
  accumulate   (/root/reference/backends/metal/plugin.mm:313-330)   48 B/px: 2 x 16 B read + 16 B write
  tonemap      (plugin.mm:334-350)                                   20 B/px: 16 B read + 4 B write
  combine      (rank-ordered sum of G shard means, 1/G) 16 (G + 1) B/px [+ 4 with the fused tonemap]
at 1920x1080 and 3840x2160 with CUDA events. Every launch works on a different buffer set of a
ring larger than the 126 MB L2, so what is timed is HBM traffic, not L2 hits. Prints one JSON
object; with --ncu-friendly it launches each kernel a few times only (for a capture under ncu).

    python tools/post_bench.py [--out gpurun_out/post_bench.json] [--ncu-friendly]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

L2_BYTES = 126 * 1024 * 1024


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--ncu-friendly", action="store_true")
    ap.add_argument("--iters", type=int, default=40)
    args = ap.parse_args()

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("post_bench.py: no CUDA device; libquasi_cuda.so has no CPU fallback")
    import quasi_b200

    lib = quasi_b200.dynamic_library.open(quasi_b200.PLUGIN_PATH)
    acc = lib.get_symbol("Q_cuda_accumulate")
    acc.restype, acc.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
    tm = lib.get_symbol("Q_cuda_tonemap_bgra8")
    tm.restype, tm.argtypes = C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    cmb = lib.get_symbol("Q_cuda_combine_shards")
    cmb.restype = C.c_int
    cmb.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_uint64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]

    peaks = ROOT / "MEASURED_PEAKS.json"
    peak = float(json.loads(peaks.read_text())["hbm_gbs"]) if peaks.exists() else 6650.0
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    iters = 3 if args.ncu_friendly else args.iters
    results = []

    def timed(name, size, bytes_per_px, n_px, sets, launch):
        """launch(k) enqueues one launch on buffer set k % sets."""
        with torch.cuda.stream(stream):
            for k in range(min(3, iters)):
                launch(k)
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(iters):
                launch(k)
            e1.record(stream)
            stream.synchronize()
        ms = e0.elapsed_time(e1) / iters
        gbs = bytes_per_px * n_px / (ms * 1e-3) / 1e9
        results.append({"kernel": name, "size": size, "bytes_per_px": bytes_per_px, "ms": round(ms, 5),
                        "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4), "buffer_sets": sets})

    for (w, h) in ((1920, 1080), (3840, 2160)):
        n_px = w * h
        layer = n_px * 16
        size = f"{w}x{h}"
        s = stream.cuda_stream

        # accumulate: 3 layers per set
        sets = max(2, -(-2 * L2_BYTES // (3 * layer)))
        bufs = [[torch.rand(n_px * 4, device=dev) for _ in range(3)] for _ in range(sets)]
        def launch_acc(k, bufs=bufs, sets=sets, n_px=n_px):
            c, hh, o = bufs[k % sets]
            assert acc(c.data_ptr(), hh.data_ptr(), o.data_ptr(), n_px, 7, s) == 0
        timed("accumulate_kernel", size, 48, n_px, sets, launch_acc)
        del bufs

        # tonemap: one layer + one BGRA8 image per set
        sets = max(2, -(-2 * L2_BYTES // (layer + n_px * 4)))
        bufs = [(torch.rand(n_px * 4, device=dev), torch.empty(n_px * 4, dtype=torch.uint8, device=dev))
                for _ in range(sets)]
        def launch_tm(k, bufs=bufs, sets=sets, w=w, h=h):
            a, o = bufs[k % sets]
            assert tm(a.data_ptr(), w, h, o.data_ptr(), 4 * w, s) == 0
        timed("tonemap_kernel", size, 20, n_px, sets, launch_tm)
        del bufs

        # combine of G local shard layers (the single-GPU form of the multi-GPU exchange), no tonemap
        for g in (2, 4, 8):
            sets = max(2, -(-2 * L2_BYTES // ((g + 1) * layer)))
            bufs = [([torch.rand(n_px * 4, device=dev) for _ in range(g)], torch.empty(n_px * 4, device=dev))
                    for _ in range(sets)]
            tables = [(C.c_void_p * g)(*[t.data_ptr() for t in shards]) for shards, _ in bufs]
            def launch_cmb(k, bufs=bufs, tables=tables, sets=sets, g=g, n_px=n_px):
                assert cmb(tables[k % sets], g, n_px, 0.0, bufs[k % sets][1].data_ptr(), None, s) == 0
            timed(f"combine_shards_kernel G={g}", size, 16 * (g + 1), n_px, sets, launch_cmb)
            del bufs, tables
        torch.cuda.empty_cache()

    out = {"hbm_peak_GBs": peak, "peak_source": "MEASURED_PEAKS.json" if peaks.exists() else "fallback",
           "iters": iters, "passes": results}
    text = json.dumps(out, indent=1)
    print(text)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(text + "\n")


if __name__ == "__main__":
    main()
