"""The plain-C oracle against fixtures produced by the reference itself.

CPU only. The fixtures under tests/golden/ were written by
tests/golden/make_golden.py from oracle/_ref (the reference's own shader text,
/root/reference/backends/metal/plugin.mm:26-351, and its own make_cornell_box).
Integer paths and, because both sides use the float semantics stated in
oracle/quasi_oracle.h, float paths too must match BIT-EXACTLY.
"""
import hashlib

import numpy as np
import pytest

from conftest import LAYERS
from oracle_lib import DEFAULT_CAMERA, bits


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_pcg_known_answers(oracle, golden):
    for x, want in golden.kat["pcg_hash"].items():
        assert oracle.pcg_hash(int(x)) == want
    # vectors quoted in SURVEY.md section 8c, derived from plugin.mm:98-102 by hand
    assert oracle.pcg_hash(0) == 0x07BB2FE2
    assert oracle.pcg_hash(0xFFFFFFFF) == 0xE62A4902


def test_seed_known_answers_and_collisions(oracle, golden):
    for px, py, f, want in golden.kat["seed"]:
        assert oracle.seed(px, py, f) == want
    # plugin.mm:295 hard-codes 1920x1080: above that size seeds collide, and the
    # uint32 product wraps after frame 2071.
    for f in (0, 7, 1023):
        assert oracle.seed(1920, 0, f) == oracle.seed(0, 1, f)
        assert oracle.seed(0, 1080, f) == oracle.seed(0, 0, f + 1)
    assert oracle.seed(0, 0, 2072) == oracle.pcg_hash((2072 * 1920 * 1080) % 2**32)


def test_random_float_stream(oracle, golden):
    for key, want in golden.kat["draws"].items():
        px, py, f = map(int, key.split(","))
        states, floats = oracle.random_floats(oracle.seed(px, py, f), 8)
        assert states == want["states"]
        assert [int(np.float32(x).view(np.uint32)) for x in floats] == want["float_bits"]
    # float(0xFFFFFFFF) rounds to 2^32, so the stream is [0, 1] INCLUSIVE
    # (plugin.mm:106): find a state that hashes to the top of the range.
    states, floats = oracle.random_floats(oracle.seed(0, 0, 0), 4)
    assert floats[0] == np.float32(0.190399364) and floats[3] == np.float32(0.214354739)
    assert np.float32(0xFFFFFFFF) / np.float32(0xFFFFFFFF) == np.float32(1.0)


@pytest.mark.parametrize("tag", ["aspect_1", "aspect_16_9"])
def test_scene_block_matches_reference_make_cornell_box(oracle, golden, tag):
    g = golden.scene[tag]
    u = oracle.scene(g["width"], g["height"])
    assert u.tobytes().hex() == g["uniforms_hex"]
    assert int(u["quad_count"]) == 16 and int(u["light_index"]) == 5


def test_single_sample_48x48(oracle, golden):
    g = golden.npz("ref_48x48_f0.npz")
    layers, rays = oracle.render_sample(oracle.scene(48, 48), 48, 48)
    for name, got in zip(LAYERS, layers):
        assert np.array_equal(bits(got), bits(g[name])), name
    assert 48 * 48 <= rays <= 5 * 48 * 48
    assert np.array_equal(oracle.tonemap_bgra8(layers[0]), g["tonemap_bgra8"])


def test_accumulated_frames_64x36(oracle, golden):
    g = golden.npz("ref_64x36_acc4.npz")
    layers, _ = oracle.render_frames(oracle.scene(64, 36), 64, 36, 4)
    for name, got in zip(LAYERS, layers):
        assert np.array_equal(bits(got), bits(g[name])), name
        assert np.all(got[..., 3] == 1.0)  # alpha stays exactly 1 under mix()


def test_sample_shard_frames_64x36(oracle, golden):
    g = golden.npz("ref_64x36_shard.npz")
    layers, _ = oracle.render_frames(oracle.scene(64, 36), 64, 36, 3, frame_first=1, frame_stride=2)
    for name, got in zip(LAYERS, layers):
        assert np.array_equal(bits(got), bits(g[name])), name


@pytest.mark.parametrize("tag", ["config1_512x512_f0", "512x512_acc3", "config2_1920x1080_f0"])
def test_full_size_checksums(oracle, golden, tag):
    g = golden.checksums[tag]
    w, h = g["width"], g["height"]
    layers, rays = oracle.render_frames(oracle.scene(w, h), w, h, g["frames"], frame_first=g["frame_first"])
    for name, got in zip(LAYERS, layers):
        assert sha(got) == g[name], name
    assert rays >= w * h * g["frames"]


def test_thread_count_cannot_change_the_image(oracle):
    u = oracle.scene(40, 30)
    a, ra = oracle.render_sample(u, 40, 30, threads=1)
    b, rb = oracle.render_sample(u, 40, 30, threads=5)
    assert ra == rb
    for x, y in zip(a, b):
        assert np.array_equal(bits(x), bits(y))


def test_first_bounce_miss_writes_zero_normal_not_half(oracle):
    # plugin.mm:226-228, 303-307: at 16:9 the leftmost column looks past the box.
    u = oracle.scene(192, 108)
    frag, rays = oracle.pathtrace_pixel(u, 0, 54, 192, 108)
    assert rays == 1
    assert frag[0].tolist() == [0, 0, 0, 1]
    assert frag[1].tolist() == [0, 0, 0, 1]
    assert frag[2].tolist() == [0, 0, 0, 1]
    assert frag[3].tolist() == [0, 0, 0, 1]


def test_light_is_seen_directly_and_ends_the_path(oracle):
    # Look straight up at the light from just below it.
    cam = dict(position=(0.0, 1.0, 0.0), target=(0.0, 2.0, 0.0), up=(0.0, 0.0, -1.0), fov=10.0)
    u = oracle.scene(9, 9, camera=cam)
    frag, rays = oracle.pathtrace_pixel(u, 4, 4, 9, 9)
    assert rays == 1
    assert frag[0].tolist() == [15, 15, 15, 1]
    assert frag[1].tolist() == [15, 15, 15, 1]  # albedo AOV = emission on lights (plugin.mm:243-246)


def test_accumulate_is_running_mean(oracle):
    rng = np.random.default_rng(1)
    hist = rng.random((4, 5, 4), dtype=np.float32)
    cur = rng.random((4, 5, 4), dtype=np.float32)
    out0 = oracle.accumulate(cur, hist, 0)
    assert np.array_equal(out0, hist + (cur - hist) * np.float32(1.0))
    out3 = oracle.accumulate(cur, hist, 3)
    assert np.array_equal(out3, hist + (cur - hist) * (np.float32(1.0) / np.float32(4.0)))


def test_combine_shards_is_rank_ordered_mean(oracle):
    rng = np.random.default_rng(2)
    shards = [rng.random((3, 3, 4), dtype=np.float32) for _ in range(4)]
    want = (((shards[0] + shards[1]) + shards[2]) + shards[3]) * np.float32(0.25)
    assert np.array_equal(oracle.combine_shards(shards), want)
