"""The SOURCE of the path-tracing kernels (quasi_b200/csrc/pathtrace.cu and mesh_trace.cu with
device_common.cuh, scan_decide.hpp, sincos_table.hpp) compiled by g++ and run on the host under a
small warp/CTA emulator (tests/cuda_emulation/cuda_on_host.hpp: one std::thread per CUDA thread,
barriers for __syncthreads and the warp collectives), against the CPU oracle, bit for bit.

What this pins without a GPU: the control flow of the persistent kernel (device-wide work queue,
the per-warp primary and bounce queues in shared memory, partial iterations on ragged and tiny
images, and the batched form: k passes over the image inside one launch ordered by per-tile epochs,
entries put aside when the previous sample of their tile is still in flight, the export of every
finished tile) and the arithmetic as written in the source. What it cannot say is what nvcc and
ptxas make of that source; tests/test_cuda_parity.py does, on the device. The product never runs
this way: the emulator lives under tests/."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))

from oracle_lib import MeshOracle, Oracle, bits  # noqa: E402

LAYERS = ("beauty", "albedo", "normal", "depth")


def build(tmp_path, flags, harness="run_pathtrace"):
    exe = tmp_path / harness
    subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-pthread", "-I/usr/local/cuda/include",
                    f"-I{ROOT / 'quasi_b200' / 'csrc'}", f"-I{ROOT / 'include'}", *flags,
                    str(ROOT / "tests" / "cuda_emulation" / f"{harness}.cpp"), "-o", str(exe)], check=True)
    return exe


def check(exe, tmp_path, oracle, width, height, frames, batch, scan):
    out = tmp_path / "layers.bin"
    run = subprocess.run([str(exe), str(width), str(height), str(frames), str(batch), scan, str(out)],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr  # (6: a batched launch's export differs from the layers)
    words = run.stdout.split()
    want, rays = oracle.render_frames(oracle.scene(width, height), width, height, frames)
    got = np.fromfile(out, dtype=np.float32).reshape(4, height, width, 4)
    for name, g, w in zip(LAYERS, got, want):
        bad = int((bits(g) != bits(w)).any(axis=-1).sum())
        assert bad == 0, f"{name}: {bad} pixels differ ({width}x{height}, {frames} frames, batch {batch}, {scan})"
    assert int(words[words.index("rays") + 1]) == rays
    assert int(words[words.index("launches") + 1]) == -(-frames // batch)


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


CASES = [(64, 36, 3, 1, "quads"), (61, 35, 5, 1, "boxes"), (97, 61, 2, 1, "generic"),     # one launch per render call
         (1, 1, 3, 1, "quads"), (9, 5, 2, 1, "quads"),                                      # partial iterations only
         # QUASI_CUDA_BATCH: full batches, a partial last one, a batch as long as the run, every scan form
         (64, 36, 7, 3, "quads"), (64, 36, 8, 8, "quads"), (61, 35, 5, 2, "boxes"), (61, 35, 5, 4, "generic"),
         (64, 36, 16, 16, "quads"), (97, 61, 12, 6, "generic"),
         # images of one or two tiles: a warp meets the next sample of a tile it has not finished (entries put aside)
         (9, 5, 6, 4, "quads"), (8, 4, 5, 5, "quads"), (17, 9, 9, 4, "quads"), (33, 17, 5, 5, "quads")]


def test_kernel_source_renders_the_oracle_image(tmp_path, oracle):
    exe = build(tmp_path, [])
    for case in CASES:
        check(exe, tmp_path, oracle, *case)


def test_mesh_kernel_source_renders_the_mesh_oracle_image(tmp_path):
    """mesh_trace.cu the same way (per-lane state machines, shared-memory stacks, descent runs and
    triangle batches voted by the warp), on the binary tree and on the 4-wide tree collapsed from it,
    against oracle/quasi_oracle_mesh.c."""
    exe = build(tmp_path, [], harness="run_mesh")
    for grid, width, height, frames in ((3, 64, 36, 2), (24, 80, 45, 2), (3, 61, 35, 5), (24, 48, 27, 4)):
        want, stats = MeshOracle(grid).render_frames(width, height, frames, max_bounces=4, use_bvh=True)
        for wide in (0, 1):
            out = tmp_path / "layers.bin"
            run = subprocess.run([str(exe), str(grid), str(width), str(height), str(frames), str(wide), str(out)],
                                 capture_output=True, text=True, timeout=900)
            assert run.returncode == 0, run.stdout + run.stderr
            got = np.fromfile(out, dtype=np.float32).reshape(4, height, width, 4)
            for name, g, w in zip(LAYERS, got, want):
                bad = int((bits(g) != bits(w)).any(axis=-1).sum())
                assert bad == 0, f"{name}: {bad} pixels differ (grid {grid}, wide {wide})"
            words = run.stdout.split()
            assert int(words[words.index("rays") + 1]) == stats["rays"]


def test_mesh_wavefront_source_renders_the_mesh_oracle_image(tmp_path):
    """mesh_wavefront.cu (ray generation, the trace kernel that refills its lanes from a queue, shade with its
    compaction, resolve in sample order) in the launch sequence of launch_mesh_wavefront: batches that divide the
    frame count and batches that do not, both tree forms, a ragged image."""
    exe = build(tmp_path, [], harness="run_mesh_wave")
    for grid, width, height, frames, wide, batch in ((3, 64, 36, 2, 0, 2), (24, 80, 45, 2, 0, 1), (3, 61, 35, 5, 0, 3),
                                                     (24, 48, 27, 4, 1, 4), (24, 33, 17, 3, 0, 16)):
        want, stats = MeshOracle(grid).render_frames(width, height, frames, max_bounces=4, use_bvh=True)
        out = tmp_path / "layers.bin"
        run = subprocess.run([str(exe), str(grid), str(width), str(height), str(frames), str(wide), str(out), str(batch)],
                             capture_output=True, text=True, timeout=900)
        assert run.returncode == 0, run.stdout + run.stderr
        got = np.fromfile(out, dtype=np.float32).reshape(4, height, width, 4)
        for name, g, w in zip(LAYERS, got, want):
            bad = int((bits(g) != bits(w)).any(axis=-1).sum())
            assert bad == 0, f"{name}: {bad} pixels differ (grid {grid}, wide {wide}, batch {batch})"
        words = run.stdout.split()
        assert int(words[words.index("rays") + 1]) == stats["rays"]


def test_kernel_source_fused_reduce_is_the_oracles_combine(tmp_path, oracle):
    """The multi-GPU form on one host: G shards rendered one after the other, each batched launch storing its
    tiles into the owners' staging slots and counting arrivals; the launch that brings a unit's G-th arrival
    sums the copies in rank order (batch_params' fused reduce). The finished frame must be the oracle's
    rank-ordered combine of the G shard images, bit for bit, on ragged sizes and with slices that end inside
    a layer."""
    exe = build(tmp_path, [])
    for width, height, samples, ranks in ((64, 36, 3, 2), (61, 35, 2, 3), (33, 17, 4, 8), (97, 61, 2, 5)):
        out = tmp_path / "frame.bin"
        run = subprocess.run([str(exe), str(width), str(height), str(samples), str(samples), "quads", str(out), str(ranks)],
                             capture_output=True, text=True, timeout=600)
        assert run.returncode == 0, run.stdout + run.stderr
        u = oracle.scene(width, height)
        shards, rays = [], 0
        for g in range(ranks):
            layers, n = oracle.render_frames(u, width, height, samples, frame_first=g, frame_stride=ranks)
            shards.append(layers)
            rays += n
        got = np.fromfile(out, dtype=np.float32).reshape(4, height, width, 4)
        for l, name in enumerate(LAYERS):
            want = oracle.combine_shards([s[l] for s in shards])
            bad = int((bits(got[l]) != bits(want)).any(axis=-1).sum())
            assert bad == 0, f"{name}: {bad} pixels differ ({width}x{height}, {samples} samples, {ranks} ranks)"
        words = run.stdout.split()
        assert int(words[words.index("rays") + 1]) == rays
