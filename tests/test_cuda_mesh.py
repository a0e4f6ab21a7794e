"""BASELINE config 3 on the GPU: the mesh path tracer (BVH traversal kernel) through the
plugin C ABI against oracle/quasi_oracle_mesh.c. Bit-exact, like the quad scene."""
import ctypes as C

import numpy as np
import pytest

from conftest import LAYERS
from oracle_lib import MeshOracle, bits

pytestmark = pytest.mark.gpu


def mesh_host(width, height, grid, bounces=4, seed=1, builder="sah", wide=0, **env):
    from quasi_b200.headless import headless_host

    return headless_host(width, height, env={"QUASI_CUDA_SCENE": "mesh", "QUASI_CUDA_MESH_GRID": grid,
                                             "QUASI_CUDA_MESH_SEED": seed, "QUASI_CUDA_MAX_BOUNCES": bounces,
                                             "QUASI_CUDA_MESH_BUILDER": builder, "QUASI_CUDA_MESH_WIDE": wide,
                                             # (these tests render two to five calls: as waves, however few they are)
                                             "QUASI_CUDA_MESH_WAVEFRONT_MIN": 1, **env})


def render(width, height, frames, grid, bounces=4, builder="sah", wide=0, **env):
    h = mesh_host(width, height, grid, bounces, builder=builder, wide=wide, **env)
    h.render(frames)
    layers = h.readback_aov()
    stats = h.stats()
    h.close()
    return layers, stats


def assert_bit_exact(got, want):
    for name, g, w in zip(LAYERS, got, want):
        bad = int((bits(g) != bits(w)).any(axis=-1).sum())
        assert bad == 0, f"{name}: {bad} pixels differ"


@pytest.mark.parametrize("grid,size,frames", [(3, (64, 36), 3), (24, (160, 90), 3), (100, (97, 61), 2)])
def test_small_meshes_vs_oracle_and_brute_force(grid, size, frames):
    w, h = size
    got, stats = render(w, h, frames, grid)
    m = MeshOracle(grid)
    want, st = m.render_frames(w, h, frames, max_bounces=4, use_bvh=True)
    assert_bit_exact(got, want)
    assert stats.rays == st["rays"]
    if grid <= 24:
        brute, _ = m.render_frames(w, h, frames, max_bounces=4, use_bvh=False)
        assert_bit_exact(got, brute)
    assert stats.nodes_visited > 0 and stats.triangles_tested > 0


@pytest.mark.parametrize("env", [{"QUASI_CUDA_MESH_WAVEFRONT": 0}, {"QUASI_CUDA_MESH_BATCH": 1},
                                 {"QUASI_CUDA_MESH_WAVEFRONT_MIN": 4},  # the default: 5 calls as waves, then 2 and 2 one by one
                                 {"QUASI_CUDA_MESH_BATCH": 2}, {"QUASI_CUDA_MESH_BATCH": 5},
                                 {"QUASI_CUDA_MESH_BATCH": 3, "QUASI_CUDA_MESH_WIDE": 1},
                                 {"QUASI_CUDA_MESH_BATCH": 4, "QUASI_CUDA_MAX_BOUNCES": 7}])
def test_both_forms_and_every_batch_length_render_the_oracle_image(env):
    """The one-kernel path tracer (QUASI_CUDA_MESH_WAVEFRONT=0: mesh_trace.cu, one launch per call) and the wavefront
    form (mesh_wavefront.cu) with batches that divide the frame's calls and batches that do not, on a ragged
    image: the same bits as the oracle, the same number of rays."""
    grid, w, h, frames = 24, 157, 83, 5
    bounces = int(env.get("QUASI_CUDA_MAX_BOUNCES", 4))
    host = mesh_host(w, h, grid, **env)
    host.render(frames)
    got, stats = host.readback_aov(), host.stats()
    want, st = MeshOracle(grid).render_frames(w, h, frames, max_bounces=bounces, use_bvh=True)
    assert_bit_exact(got, want)
    assert stats.rays == st["rays"] and stats.paths == frames * w * h
    # accumulation goes on across a readback (a batch cut short by it) and restarts with camera_dirty
    host.render_frame(camera_dirty=False)
    host.render_frame(camera_dirty=False)
    more, _ = MeshOracle(grid).render_frames(w, h, frames + 2, max_bounces=bounces, use_bvh=True)
    assert_bit_exact(host.readback_aov(), more)
    host.render_frame(camera_dirty=True)
    host.render_frame(camera_dirty=False)
    again, _ = MeshOracle(grid).render_frames(w, h, 2, max_bounces=bounces, use_bvh=True)
    assert_bit_exact(host.readback_aov(), again)
    host.close()


def test_plugin_built_bvh_equals_oracle_bvh():
    grid = 64
    h = mesh_host(16, 16, grid)
    v = h.plugin.cuda_mesh_view()
    m = MeshOracle(grid)
    assert v.n_triangles == m.n_triangles and v.n_nodes == m.n_nodes and v.depth <= 32
    assert C.string_at(v.nodes, 32 * v.n_nodes) == m.nodes().tobytes()
    assert C.string_at(v.triangles, 48 * v.n_triangles) == m.triangles().tobytes()
    order = np.frombuffer(C.string_at(v.leaf_order, 4 * v.n_triangles), dtype=np.uint32)
    assert np.array_equal(order, m.leaf_order())
    assert any("Mesh scene:" in s for s in h.messages)
    h.close()
    from quasi_b200.headless import headless_host

    q = headless_host(8, 8)           # the demo scene has no mesh view
    assert q.plugin.cuda_mesh_view() is None
    q.close()


def test_million_triangles_vs_oracle():
    """The full 999 710-triangle scene at reduced resolution, 2 spp, 4 bounces."""
    w, h, frames = 480, 270, 2
    got, stats = render(w, h, frames, 707)
    m = MeshOracle(707)
    want, st = m.render_frames(w, h, frames, max_bounces=4, use_bvh=True)
    assert_bit_exact(got, want)
    assert stats.rays == st["rays"]


def test_config3_full_size_properties():
    """1920x1080, 16 spp, 4 bounces on the 1 M-triangle scene: size-independent properties."""
    got, stats = render(1920, 1080, 16, 707)
    again, stats2 = render(1920, 1080, 16, 707)
    assert stats.paths == 16 * 1920 * 1080 and stats.rays == stats2.rays
    assert stats.paths <= stats.rays <= 4 * stats.paths
    for a, b in zip(got, again):
        assert np.array_equal(bits(a), bits(b))          # deterministic
        assert np.all(a[..., 3] == 1.0) and np.isfinite(a).all() and a.min() >= 0.0
    assert got[0][:, :300, :3].max() == 0.0              # the margins beside the box stay black
    assert got[0][..., :3].mean() > 0.02


@pytest.mark.parametrize("grid,size,frames", [(3, (64, 36), 2), (24, (160, 90), 3), (100, (97, 61), 2),
                                              (707, (480, 270), 2)])
def test_gpu_built_bvh_renders_the_oracle_image(grid, size, frames):
    """The tree built on the GPU (lbvh_build.cu) is a different tree from the oracle's SAH
    tree; the closest hit does not depend on the tree, so the image is the same bits."""
    w, h = size
    got, stats = render(w, h, frames, grid, builder="lbvh")
    want, st = MeshOracle(grid).render_frames(w, h, frames, max_bounces=4, use_bvh=True)
    assert_bit_exact(got, want)
    assert stats.rays == st["rays"]
    assert stats.nodes_visited > 0 and stats.triangles_tested > 0


@pytest.mark.parametrize("grid,size,frames,builder", [(1, (64, 36), 2, "sah"), (3, (64, 36), 2, "sah"),
                                                      (24, (160, 90), 3, "sah"), (100, (97, 61), 2, "sah"),
                                                      (100, (97, 61), 2, "lbvh"), (707, (480, 270), 2, "sah")])
def test_wide_tree_renders_the_oracle_image(grid, size, frames, builder):
    """QUASI_CUDA_MESH_WIDE=1: the kernel walks the 4-wide tree collapsed from the binary one
    (wide_bvh.hpp); it prunes what the binary tree prunes, so the image is the same bits, from
    either builder's tree."""
    w, h = size
    got, stats = render(w, h, frames, grid, builder=builder, wide=1)
    want, st = MeshOracle(grid).render_frames(w, h, frames, max_bounces=4, use_bvh=True)
    assert_bit_exact(got, want)
    assert stats.rays == st["rays"]
    assert stats.nodes_visited > 0 and stats.triangles_tested > 0


def test_gpu_built_bvh_is_a_tree_over_all_triangles():
    grid = 40
    h = mesh_host(16, 16, grid, builder="lbvh")
    assert any("built on the GPU" in s for s in h.messages), h.messages
    v = h.plugin.cuda_mesh_view()
    n = v.n_triangles
    assert n == 12 + 2 * grid * grid and v.n_nodes == 2 * n - 1
    nodes = np.frombuffer(C.string_at(v.nodes, 32 * v.n_nodes), dtype=np.dtype(
        [("lo", "<f4", 3), ("left_first", "<u4"), ("hi", "<f4", 3), ("count", "<u4")]))
    tris = np.frombuffer(C.string_at(v.triangles, 48 * n), dtype=np.dtype(
        [("v0", "<f4", 3), ("m", "<u4"), ("e1", "<f4", 3), ("p0", "<u4"), ("e2", "<f4", 3), ("p1", "<u4")]))
    order = np.frombuffer(C.string_at(v.leaf_order, 4 * n), dtype=np.uint32)
    assert np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))     # a permutation
    corners = np.stack([tris["v0"], tris["v0"] + tris["e1"], tris["v0"] + tris["e2"]], axis=1)  # (n, 3, 3)
    seen = np.zeros(n, dtype=np.int32)
    depth = 0
    stack = [(0, 1, None, None)]
    while stack:
        i, d, plo, phi = stack.pop()
        node = nodes[i]
        depth = max(depth, d)
        if plo is not None:   # a child's box lies inside its parent's (both padded)
            assert np.all(node["lo"] >= plo - 1e-4) and np.all(node["hi"] <= phi + 1e-4)
        if node["count"] == 0:
            left = int(node["left_first"])
            stack.append((left, d + 1, node["lo"], node["hi"]))
            stack.append((left + 1, d + 1, node["lo"], node["hi"]))
        else:
            assert 1 <= node["count"] <= 4
            for slot in range(int(node["left_first"]), int(node["left_first"] + node["count"])):
                t = int(order[slot])
                seen[t] += 1
                assert np.all(corners[t] >= node["lo"]) and np.all(corners[t] <= node["hi"])
                assert np.all(corners[t] - node["lo"] >= 9e-6) and np.all(node["hi"] - corners[t] >= 9e-6)  # padding
    assert np.all(seen == 1)
    assert depth == v.depth <= 32
    h.close()
