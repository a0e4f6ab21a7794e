"""The pruning arithmetic of the mesh traversal (quasi_b200/csrc/slab_test.hpp, the code the CUDA
kernel compiles) run on the host on the backend's own scene and BVH: the closest hit found
through the tree must be the brute-force closest hit over all triangles, for rays that
include exact zeros and denormals in the direction and origins exactly on box planes."""
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
