"""BASELINE config 3 on the CPU: the mesh oracle against its own brute force, and the
backend's host-side mesh generator + SAH BVH builder against the oracle's independent
implementation, bit for bit (triangles, node array, leaf order)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import LAYERS
from oracle_lib import MeshOracle, bits

ROOT = Path(__file__).resolve().parent.parent
TOOL = ROOT / "quasi_b200" / "bin" / "tests" / "mesh_dump_tool"


def backend_build(grid, seed, tmp_path, task_limit=None):
    subprocess.run(["make", "-s", "-C", str(ROOT), str(TOOL.relative_to(ROOT))], check=True)
    out = tmp_path / f"mesh_{grid}.bin"
    subprocess.run([str(TOOL), str(grid), str(seed), str(out)] + ([str(task_limit)] if task_limit else []), check=True)
    raw = out.read_bytes()
    nt, nn, depth, _ = np.frombuffer(raw, dtype=np.uint32, count=4)
    o = 16
    tris = raw[o:o + 48 * nt]; o += 48 * nt
    nodes = raw[o:o + 32 * nn]; o += 32 * nn
    order = np.frombuffer(raw, dtype=np.uint32, count=nt, offset=o)
    return int(nt), int(nn), int(depth), tris, nodes, order


@pytest.mark.parametrize("grid,seed", [(1, 1), (7, 3), (24, 1), (100, 9)])
def test_backend_builder_equals_oracle_builder(grid, seed, tmp_path):
    nt, nn, depth, tris, nodes, order = backend_build(grid, seed, tmp_path)
    m = MeshOracle(grid, seed)
    assert nt == m.n_triangles == 12 + 2 * grid * grid and nn == m.n_nodes
    assert tris == m.triangles().tobytes()
    assert nodes == m.nodes().tobytes()          # the bit-exact BVH build-order contract
    assert np.array_equal(order, m.leaf_order())
    assert sorted(order.tolist()) == list(range(nt))


@pytest.mark.parametrize("task_limit", [16, 500, 4000])
def test_concurrent_build_is_the_sequential_build(task_limit, tmp_path):
    """The backend builds ranges above `task_limit` triangles by splitting them and building the
    halves on different threads (mesh_scene.hpp:build_parallel); whatever the limit, the node
    array and the leaf order are the oracle's, which is built by one thread, byte for byte."""
    nt, nn, depth, tris, nodes, order = backend_build(60, 5, tmp_path, task_limit)
    m = MeshOracle(60, 5)
    assert nn == m.n_nodes and nodes == m.nodes().tobytes() and np.array_equal(order, m.leaf_order())


def test_million_triangle_build_matches_and_fits_the_traversal_stack(tmp_path):
    nt, nn, depth, tris, nodes, order = backend_build(707, 1, tmp_path)
    m = MeshOracle(707, 1)
    assert nt == 999_710 and nn == m.n_nodes
    assert nodes == m.nodes().tobytes() and np.array_equal(order, m.leaf_order())
    assert depth <= 32
    n = m.nodes()
    leaves = n[n["count"] > 0]
    assert leaves["count"].max() <= 4 and int(leaves["count"].sum()) == nt


@pytest.mark.parametrize("grid", [3, 24])
def test_bvh_traversal_equals_brute_force(grid):
    m = MeshOracle(grid)
    a, sa = m.render_frames(64, 36, 2, max_bounces=4, use_bvh=True)
    b, sb = m.render_frames(64, 36, 2, max_bounces=4, use_bvh=False)
    assert sa["rays"] == sb["rays"]
    for name, x, y in zip(LAYERS, a, b):
        assert np.array_equal(bits(x), bits(y)), name
    assert sa["triangles_tested"] < sb["triangles_tested"]


def test_scene_is_lit_and_bounded():
    m = MeshOracle(16)
    layers, st = m.render_frames(48, 27, 4, max_bounces=4)
    assert layers[0][..., :3].mean() > 0.01 and np.isfinite(layers[0]).all()
    assert np.all(layers[0][..., 3] == 1.0)
    t = m.triangles()
    hi = (t["v0"] + np.maximum(t["e1"], 0) + np.maximum(t["e2"], 0))[12:, 1].max()
    assert 0.03 <= t["v0"][12:, 1].min() and hi <= 0.61  # terrain heights stay in [0.03, 0.6]
