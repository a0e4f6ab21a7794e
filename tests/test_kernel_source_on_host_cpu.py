"""The SOURCE of the quad path-tracing kernel (quasi_b200/csrc/pathtrace.cu with device_common.cuh,
scan_decide.hpp, sincos_table.hpp) compiled by g++ and run on the host under a small warp/CTA
emulator (tests/cuda_emulation/cuda_on_host.hpp: one std::thread per CUDA thread, barriers for
__syncthreads and the warp collectives), against the CPU oracle, bit for bit.

What this pins without a GPU: the control flow of the persistent kernel (device-wide tile queue,
ballot/popc path regeneration, ragged edge tiles, the batched form in which a lane gives its pixel
several consecutive samples, and the sample-major one with its per-tile epochs, where running the
source this way found a self-deadlock and a lane-dependent readiness test before any GPU saw them)
and the arithmetic of every build-time option, as written in the source. What it cannot say is what nvcc and ptxas make of that source; tests/test_cuda_parity.py
does, on the device. The product never runs this way: the emulator lives under tests/."""
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


def check(exe, tmp_path, oracle, width, height, frames, batch, scan, order=""):
    """`order` "s": the sample-major form of the batched launch."""
    out = tmp_path / "layers.bin"
    run = subprocess.run([str(exe), str(width), str(height), str(frames), f"{batch}{order}", scan, str(out)],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
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


CASES = [(64, 36, 3, 1, "quads"), (61, 35, 5, 1, "boxes"), (97, 61, 2, 1, "generic"),     # one launch per frame
         (64, 36, 7, 3, "quads"), (64, 36, 8, 8, "quads"), (61, 35, 5, 2, "boxes"),        # QUASI_CUDA_BATCH
         (61, 35, 5, 4, "generic"), (9, 5, 6, 4, "quads"),
         # QUASI_CUDA_BATCH_ORDER=sample; on the small images a warp meets its own unfinished tile again
         (64, 36, 8, 8, "quads", "s"), (61, 35, 6, 3, "boxes", "s"), (61, 35, 5, 2, "generic", "s"),
         (97, 61, 4, 4, "quads", "s"), (9, 5, 6, 4, "quads", "s"), (8, 4, 5, 5, "quads", "s")]


def test_kernel_source_renders_the_oracle_image(tmp_path, oracle):
    """The default build: every scan form, ragged sizes, and the batched form (full batches, a
    partial last one, a batch as long as the run, an image smaller than two tiles)."""
    exe = build(tmp_path, [])
    for case in CASES:
        check(exe, tmp_path, oracle, *case)


@pytest.mark.parametrize("flags", [["-DQUASI_SCAN_RECORDS=1"], ["-DQUASI_SINCOS_LOOKUP=1"], ["-DQUASI_SINCOS_TABLE=1"],
                                   ["-DQUASI_SCAN_RECORDS=1", "-DQUASI_SINCOS_LOOKUP=1"]],
                         ids=["scan-records", "sincos-lookup", "sincos-table", "records+lookup"])
def test_build_time_options_render_the_oracle_image(tmp_path, oracle, flags):
    """The candidates of DESIGN.md section 9, each against the oracle in the plain and the batched form."""
    exe = build(tmp_path, flags)
    for case in [(64, 36, 4, 1, "quads"), (61, 35, 6, 4, "quads"), (61, 35, 3, 1, "generic"), (61, 35, 3, 1, "boxes"),
                 (61, 35, 6, 3, "quads", "s")]:
        check(exe, tmp_path, oracle, *case)


def test_mesh_kernel_source_renders_the_mesh_oracle_image(tmp_path):
    """mesh_trace.cu the same way (per-lane state machines, shared-memory stacks, descent runs and
    triangle batches voted by the warp), on the binary tree and on the 4-wide tree collapsed from it,
    launch by launch and in the batched form, against oracle/quasi_oracle_mesh.c."""
    exe = build(tmp_path, [], harness="run_mesh")
    for grid, width, height, frames, batch in ((3, 64, 36, 2, 1), (24, 80, 45, 2, 1), (3, 61, 35, 5, 2), (24, 48, 27, 4, 4)):
        want, stats = MeshOracle(grid).render_frames(width, height, frames, max_bounces=4, use_bvh=True)
        for wide in (0, 1):
            out = tmp_path / "layers.bin"
            run = subprocess.run([str(exe), str(grid), str(width), str(height), str(frames), str(wide), str(out), str(batch)],
                                 capture_output=True, text=True, timeout=900)
            assert run.returncode == 0, run.stdout + run.stderr
            got = np.fromfile(out, dtype=np.float32).reshape(4, height, width, 4)
            for name, g, w in zip(LAYERS, got, want):
                bad = int((bits(g) != bits(w)).any(axis=-1).sum())
                assert bad == 0, f"{name}: {bad} pixels differ (grid {grid}, wide {wide})"
            words = run.stdout.split()
            assert int(words[words.index("rays") + 1]) == stats["rays"]
