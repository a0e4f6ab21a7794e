"""The pruning arithmetic of the mesh traversal (quasi_b200/csrc/slab_test.hpp, the code the CUDA
kernel compiles) run on the host on the backend's own scene and BVH: the closest hit found
through the tree must be the brute-force closest hit over all triangles, for rays that
include exact zeros and denormals in the direction and origins exactly on box planes.
The same for the 4-wide tree collapsed from the binary one (wide_bvh.hpp: the collapse, with and
without a binding stack limit, and the visit of a wide node as the kernel compiles it)."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("flags", [(), ("-DQUASI_SCAN_CHECK_PERTURB",)], ids=["exact-reciprocal", "reciprocal-2ulp-off"])
def test_tree_traversal_finds_the_brute_force_hit(tmp_path, flags):
    exe = tmp_path / "mesh_prune_check"
    subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-pthread", *flags,
                    f"-I{ROOT / 'quasi_b200' / 'csrc'}", str(ROOT / "tests" / "mesh_prune_check.cpp"), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe), "32", "80000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "mismatches 0" in out.stdout
    counts = dict(zip(out.stdout.split()[0::2], out.stdout.split()[1::2]))
    assert int(counts["hits"]) > 40000 and float(counts["boxes_per_ray"]) < 120   # the tree does prune


@pytest.mark.parametrize("flags", [(), ("-DQUASI_SCAN_CHECK_PERTURB",)], ids=["exact-reciprocal", "reciprocal-2ulp-off"])
def test_wide_tree_is_the_binary_tree_regrouped_and_finds_the_brute_force_hit(tmp_path, flags):
    exe = tmp_path / "mesh_prune_check"
    subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-pthread", *flags,
                    f"-I{ROOT / 'quasi_b200' / 'csrc'}", str(ROOT / "tests" / "mesh_prune_check.cpp"), "-o", str(exe)],
                   check=True)
    # stack limit 31 (the kernel's) does not bind at this size; 14 does: nodes on the longest paths stay narrow
    for grid, limit, most_visits in (("32", "31", 16.0), ("32", "14", 24.0), ("3", "31", 8.0)):
        out = subprocess.run([str(exe), grid, "60000", "wide", limit], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout[-2000:]
        assert "structure violations 0" in out.stdout and "mismatches 0" in out.stdout
        words = out.stdout.replace("(", " ").replace(")", " ").replace(",", " ").split()
        bound = int(words[words.index("bound") + 1])
        deepest = int(words[words.index("deepest") + 2])
        visits = float(words[words.index("visits_per_ray") + 1])
        assert deepest <= bound <= int(limit)
        assert visits < most_visits


@pytest.mark.parametrize("flags", [(), ("-DQUASI_SCAN_CHECK_PERTURB",)], ids=["exact-reciprocal", "reciprocal-2ulp-off"])
def test_quantized_tree_holds_the_float_tree_and_finds_the_brute_force_hit(tmp_path, flags):
    """quantized_bvh.hpp as the wavefront trace kernel compiles it: 16-bit boxes on a grid, plane distances by one
    fused multiply-add from the index in the mantissa of 2^23, near and far plane chosen by the ray's sign."""
    exe = tmp_path / "mesh_prune_check"
    subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-pthread", *flags,
                    f"-I{ROOT / 'quasi_b200' / 'csrc'}", str(ROOT / "tests" / "mesh_prune_check.cpp"), "-o", str(exe)],
                   check=True)
    for grid in ("32", "3"):
        out = subprocess.run([str(exe), grid, "80000", "quant"], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout[-2000:]
        assert "structure violations 0" in out.stdout and "mismatches 0" in out.stdout
