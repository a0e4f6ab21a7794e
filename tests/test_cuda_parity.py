"""Parity of libquasi_cuda.so against the CPU oracle and the reference goldens.

GPU tests: every render goes through the plugin C ABI (loader -> Q_plugin_create
-> Q_plugin_render -> Q_plugin_readback_aov), the way the reference host drives
a backend (/root/reference/src/quasi/host/main.cpp:218-321). Integer AND float
results must be bit-exact: both sides evaluate the same IEEE binary32 operations
in the same order (DESIGN.md "Numerics"). The tolerance BASELINE.json allows
(mean relative error <= 1e-3) is asserted too, as the weaker published gate.
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

from conftest import LAYERS
from oracle_lib import bits

pytestmark = pytest.mark.gpu

MEAN_REL_ERR_GATE = 1e-3  # BASELINE.json north_star


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def host(width, height, **kw):
    from quasi_b200.headless import headless_host

    return headless_host(width, height, **kw)


def render(width, height, frames, **kw):
    h = host(width, height, **kw)
    h.render(frames)
    layers = h.readback_aov()
    stats = h.stats()
    h.close()
    return layers, stats


def mean_rel_err(got, want):
    g = got[..., :3].astype(np.float64)
    w = want[..., :3].astype(np.float64)
    return float(np.mean(np.abs(g - w) / np.maximum(np.abs(w), 1e-3)))


def assert_bit_exact(got_layers, want_layers):
    for name, g, w in zip(LAYERS, got_layers, want_layers):
        bad = int((bits(g) != bits(w)).any(axis=-1).sum())
        assert bad == 0, f"{name}: {bad} pixels differ"
        assert mean_rel_err(g, w) <= MEAN_REL_ERR_GATE


def test_config1_512x512_1spp_bit_exact(torch_cuda, oracle, golden):
    """BASELINE config 1: the oracle image, and the reference checksum of it."""
    got, stats = render(512, 512, 1)
    want, rays = oracle.render_frames(oracle.scene(512, 512), 512, 512, 1)
    assert_bit_exact(got, want)
    assert stats.rays == rays and stats.paths == 512 * 512 and stats.frame_count == 1
    g = golden.checksums["config1_512x512_f0"]
    for name, layer in zip(LAYERS, got):
        assert sha(layer) == g[name]


def test_golden_48x48_and_tonemapped_drawable(torch_cuda, golden):
    g = golden.npz("ref_48x48_f0.npz")
    h = host(48, 48, present=True)
    h.render(1)
    layers = h.readback_aov()
    shown = h.context.drawable_tensor().cpu().numpy()
    h.close()
    assert_bit_exact(layers, [g[n] for n in LAYERS])
    # The drawable is display-only (not parity-gated): GPU powf is within a few
    # ulp of the correctly rounded pow, so a channel may land one LSB away.
    diff = np.abs(shown.astype(np.int16) - g["tonemap_bgra8"].astype(np.int16))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3
    assert np.all(shown[..., 3] == 255)


def test_golden_accumulated_frames(torch_cuda, golden):
    g = golden.npz("ref_64x36_acc4.npz")
    got, stats = render(64, 36, 4)
    assert_bit_exact(got, [g[n] for n in LAYERS])
    assert stats.frame_count == 4 and stats.kernel_launches == 1  # four collected calls, one batched launch


def test_golden_sample_shard(torch_cuda, golden):
    """Shard 1 of 2 renders RNG frame indices 1, 3, 5."""
    g = golden.npz("ref_64x36_shard.npz")
    got, _ = render(64, 36, 3, shard_index=1, shard_count=2)
    assert_bit_exact(got, [g[n] for n in LAYERS])


@pytest.mark.parametrize("size,frames", [((97, 61), 2), ((33, 17), 5), ((1, 1), 1), ((16, 16), 1)])
def test_ragged_sizes_vs_oracle(torch_cuda, oracle, size, frames):
    w, h = size
    got, stats = render(w, h, frames)
    want, rays = oracle.render_frames(oracle.scene(w, h), w, h, frames)
    assert_bit_exact(got, want)
    assert stats.rays == rays


def test_moved_camera_vs_oracle(torch_cuda, oracle):
    h = host(160, 120)
    h.camera.azimuth, h.camera.elevation, h.camera.distance, h.camera.fov = 0.6, 0.35, 2.8, 55.0
    h.camera.update_position()
    h.render(3)
    got = h.readback_aov()
    cam = dict(position=tuple(h.camera.position), target=tuple(h.camera.target), up=(0, 1, 0), fov=55.0)
    h.close()
    want, _ = oracle.render_frames(oracle.scene(160, 120, camera=cam), 160, 120, 3)
    assert_bit_exact(got, want)


def test_config2_1080p_first_frame_matches_reference_checksum(torch_cuda, golden):
    g = golden.checksums["config2_1920x1080_f0"]
    got, stats = render(1920, 1080, 1)
    for name, layer in zip(LAYERS, got):
        assert sha(layer) == g[name], name
    assert stats.paths == 1920 * 1080


def test_config2_1080p_8spp_vs_oracle(torch_cuda, oracle):
    got, stats = render(1920, 1080, 8)
    want, rays = oracle.render_frames(oracle.scene(1920, 1080), 1920, 1080, 8)
    assert_bit_exact(got, want)
    assert stats.rays == rays


def test_config4_4k_colliding_seeds_match_reference_checksum(torch_cuda, golden):
    """Above 1920x1080 the reference's seed collides between pixels (plugin.mm:295)."""
    g = golden.checksums["config4_3840x2160_f1"]
    got, _ = render(3840, 2160, 1, shard_index=1, shard_count=2)  # first frame index = 1
    for name, layer in zip(LAYERS, got):
        assert sha(layer) == g[name], name


def test_full_size_properties_1080p_64spp(torch_cuda, golden):
    """BASELINE config 2 at full size: the reference's own checksum of the finished frame (oracle/_ref, written
    by tests/golden/make_golden.py), then size-independent properties."""
    got, stats = render(1920, 1080, 64)
    g = golden.checksums["config2_1920x1080_64spp"]
    for name, layer in zip(LAYERS, got):
        assert sha(layer) == g[name], name
    beauty, albedo, normal, depth = got
    assert stats.frame_count == 64 and stats.paths == 64 * 1920 * 1080
    assert 64 * 1920 * 1080 <= stats.rays <= 5 * 64 * 1920 * 1080
    for layer in got:
        assert np.all(layer[..., 3] == 1.0)          # mix() keeps alpha exactly 1
        assert np.isfinite(layer).all() and layer.min() >= 0.0
    assert np.array_equal(depth[..., 0], depth[..., 1]) and depth.max() <= 1.0
    # determinism: a second run is bit-identical
    again, stats2 = render(1920, 1080, 64)
    assert stats2.rays == stats.rays
    for a, b in zip(got, again):
        assert np.array_equal(bits(a), bits(b))
    # mirror symmetry of the empty margins: columns that look past the box stay black
    assert beauty[:, :300, :3].max() == 0.0 and beauty[:, -300:, :3].max() == 0.0


def test_sharded_frames_recombine_to_the_single_gpu_sample_set(torch_cuda, oracle):
    """Two shards of 4 frames hold exactly frames 0..7; their rank-ordered mean
    matches the oracle's combine and is within float tolerance of the 8-frame mean."""
    w, h = 96, 54
    s0, _ = render(w, h, 4, shard_index=0, shard_count=2)
    s1, _ = render(w, h, 4, shard_index=1, shard_count=2)
    single, _ = render(w, h, 8)
    for l in range(4):
        want = oracle.combine_shards([s0[l], s1[l]])
        o0, _ = oracle.render_frames(oracle.scene(w, h), w, h, 4, frame_first=0, frame_stride=2)
        assert np.array_equal(bits(s0[l]), bits(o0[l]))
        assert np.allclose(want, single[l], rtol=2e-6, atol=2e-6)


def test_camera_dirty_and_resize_restart_accumulation(torch_cuda, oracle):
    h = host(64, 48)
    h.render(3)
    h.render_frame(camera_dirty=True)  # plugin.mm:643-646: restart, this call is sample 0
    one = h.readback_aov()
    assert h.stats().frame_count == 1
    want, _ = oracle.render_frames(oracle.scene(64, 48), 64, 48, 1)
    assert_bit_exact(one, want)
    h.context.resize(40, 30)           # plugin.mm:638-640
    h.render(2)
    two = h.readback_aov()
    want, _ = oracle.render_frames(oracle.scene(40, 30), 40, 30, 2)
    assert_bit_exact(two, want)
    h.close()


def test_readback_before_any_render_returns_null(torch_cuda):
    h = host(32, 32)
    assert h.readback_aov() is None and h.plugin.readback_numpy() is None
    h.render(1)
    beauty = h.plugin.readback_numpy()   # Q_plugin_readback: beauty only
    assert beauty.shape == (32, 32, 4) and np.array_equal(bits(beauty), bits(h.readback_aov()[0]))
    assert "Cornell Box path tracer initialized (CUDA)" in h.messages
    assert "AOV readback complete" in h.messages and "HDR readback complete" in h.messages
    h.close()
    assert h.messages[-1] == "Path tracer destroyed"


def test_null_frame_and_null_stream_are_ignored(torch_cuda):
    from quasi_b200 import abi

    h = host(16, 16)
    h.plugin.render(None)
    f = abi.Q_render_frame()
    f.width = f.height = 16
    f.command_buffer = None            # plugin.mm:632: no command buffer, no work
    h.plugin.render(f)
    assert h.stats().frame_count == 0
    h.close()


def test_standalone_passes_vs_oracle(torch_cuda, oracle):
    torch = torch_cuda
    h = host(8, 8)
    fn = h.plugin.fn
    rng = np.random.default_rng(5)
    w, ht = 61, 47
    cur = (rng.random((ht, w, 4), dtype=np.float32) * 20).astype(np.float32)
    hist = (rng.random((ht, w, 4), dtype=np.float32) * 20).astype(np.float32)
    dcur, dhist = torch.from_numpy(cur).cuda(), torch.from_numpy(hist).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    for fc in (0, 1, 7, 1023):
        dout = torch.empty_like(dcur)
        assert fn("Q_cuda_accumulate")(dcur.data_ptr(), dhist.data_ptr(), dout.data_ptr(), w * ht, fc, stream) == 0
        torch.cuda.synchronize()
        assert np.array_equal(bits(dout.cpu().numpy()), bits(oracle.accumulate(cur, hist, fc)))
    # in place (output aliases history), as the fused kernel does
    dalias = dhist.clone()
    assert fn("Q_cuda_accumulate")(dcur.data_ptr(), dalias.data_ptr(), dalias.data_ptr(), w * ht, 3, stream) == 0
    torch.cuda.synchronize()
    assert np.array_equal(bits(dalias.cpu().numpy()), bits(oracle.accumulate(cur, hist, 3)))
    # tonemap
    dimg = torch.zeros((ht, w, 4), dtype=torch.uint8, device="cuda")
    assert fn("Q_cuda_tonemap_bgra8")(dcur.data_ptr(), w, ht, dimg.data_ptr(), 4 * w, stream) == 0
    torch.cuda.synchronize()
    diff = np.abs(dimg.cpu().numpy().astype(np.int16) - oracle.tonemap_bgra8(cur).astype(np.int16))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3
    # combine (rank-ordered mean) with and without the fused tonemap
    shards = [torch.from_numpy((rng.random((ht, w, 4), dtype=np.float32) * 4).astype(np.float32)).cuda()
              for _ in range(5)]
    ptrs = (C.c_void_p * 5)(*[s.data_ptr() for s in shards])
    dout = torch.empty_like(shards[0])
    assert fn("Q_cuda_combine_shards")(ptrs, 5, w * ht, 0.0, dout.data_ptr(), dimg.data_ptr(), stream) == 0
    torch.cuda.synchronize()
    want = oracle.combine_shards([s.cpu().numpy() for s in shards])
    assert np.array_equal(bits(dout.cpu().numpy()), bits(want))
    diff = np.abs(dimg.cpu().numpy().astype(np.int16) - oracle.tonemap_bgra8(want).astype(np.int16))
    assert diff.max() <= 1
    h.close()


def test_sincos_probe_matches_correctly_rounded_values(torch_cuda, oracle):
    """The kernel's own binary64 sin/cos, rounded once, against the oracle's on
    2^24 arguments phi = 2*pi*r1 drawn the way the shader draws them, plus the
    ends of the range and every quadrant boundary neighbourhood."""
    torch = torch_cuda
    h = host(8, 8)
    rng = np.random.default_rng(11)
    state = rng.integers(0, 2**32, size=1 << 24, dtype=np.uint64).astype(np.uint32)
    r1 = state.astype(np.float32) * np.float32(2.0**-32)
    phi = np.float32(6.28318548202514648) * r1
    edges = np.array([0.0, 6.2831855, np.pi / 2, np.pi, 3 * np.pi / 2, np.pi / 4, 1e-30, 1e-10], dtype=np.float32)
    near = np.concatenate([np.nextafter(edges, np.float32(10), dtype=np.float32),
                           np.nextafter(edges, np.float32(-1), dtype=np.float32)])
    phi = np.concatenate([phi, edges, near[(near >= 0) & (near <= np.float32(6.2831855))]]).astype(np.float32)
    d = torch.from_numpy(phi).cuda()
    ds, dc = torch.empty_like(d), torch.empty_like(d)
    rc = h.plugin.fn("Q_cuda_probe_sincos")(d.data_ptr(), ds.data_ptr(), dc.data_ptr(), phi.size,
                                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    want_s, want_c = oracle.sincos(phi)
    bad_s = int((bits(ds.cpu().numpy()) != bits(want_s)).sum())
    bad_c = int((bits(dc.cpu().numpy()) != bits(want_c)).sum())
    h.close()
    # 2^-28 per call would allow ~0.1 expected mismatches over 3.4e7 evaluations
    assert bad_s + bad_c <= 1, (bad_s, bad_c)


def test_sincos_correctly_rounded_over_the_rng_grid(torch_cuda, oracle):
    """cos/sin of phi = 2*pi*r1 decide bounce directions; the kernel's must be
    the correctly rounded binary32 values the oracle defines. Exercised through
    rendering: many bounces, many pixels, zero mismatches is the evidence; this
    test adds a long accumulation at small size to reach deep into the stream."""
    got, stats = render(48, 27, 96)
    want, rays = oracle.render_frames(oracle.scene(48, 27), 48, 27, 96)
    assert_bit_exact(got, want)
    assert stats.rays == rays


def test_accumulation_survives_a_plugin_reload_when_the_host_keeps_it(torch_cuda, oracle):
    """with a Q_cuda_persistent_state in Q_gpu_context::layer the
    destroyed instance parks its layers with the host and the next instance continues at
    the same sample index: 5 + reload + 5 frames are the bits of 10 uninterrupted frames.
    Without it a reload restarts at sample 0, as in the reference (plugin.mm:594-617)."""
    from quasi_b200.headless import headless_host

    w, h = 200, 120
    want10, _ = oracle.render_frames(oracle.scene(w, h), w, h, 10)
    want5, _ = oracle.render_frames(oracle.scene(w, h), w, h, 5)

    host = headless_host(w, h, keep_accumulation=True)
    host.render(5)
    host.reload()
    assert any("adopted" in m for m in host.messages)
    assert host.context.persistent.valid == 0          # the block is the plugin's again
    kept = host.readback_aov()                          # readable before the first new frame
    for g, wv in zip(kept, want5):
        assert np.array_equal(bits(g), bits(wv))
    host.render(5)
    for name, g, wv in zip(LAYERS, host.readback_aov(), want10):
        assert np.array_equal(bits(g), bits(wv)), name
    assert host.stats().frame_count == 10

    # a parked block that does not fit (other viewport) is released, not adopted
    host.context.synchronize()
    host.plugin.destroy()
    assert host.context.persistent.valid == 1 and host.context.persistent.frame_count == 10
    host.context.persistent.width += 1
    host.library.close()
    host._load()
    assert host.context.persistent.valid == 0 and host.plugin.readback_aov_numpy() is None
    host.render_frame(camera_dirty=False)
    host.render(4)
    for g, wv in zip(host.readback_aov(), want5):
        assert np.array_equal(bits(g), bits(wv))
    host.close()

    plain = headless_host(w, h)
    plain.render(5)
    plain.reload()
    assert plain.plugin.readback_aov_numpy() is None    # nothing rendered yet in the new instance
    plain.render(5)
    for g, wv in zip(plain.readback_aov(), want5):
        assert np.array_equal(bits(g), bits(wv))
    plain.close()


@pytest.mark.parametrize("scan", ["generic", "boxes", "generic-boxes"])
def test_every_scan_variant_is_bit_exact(torch_cuda, oracle, scan):
    """QUASI_CUDA_SCAN selects the any-scene kernel (one loop per quad shape instead of the
    code unrolled for the demo scene) and/or the grouping of each block's five faces into one
    slab test. All of them decide the same closest hit."""
    from quasi_b200.headless import headless_host

    w, h, frames = 211, 119, 4
    host = headless_host(w, h, env={"QUASI_CUDA_SCAN": scan})
    host.render(frames)
    got, stats = host.readback_aov(), host.stats()
    host.close()
    want, rays = oracle.render_frames(oracle.scene(w, h), w, h, frames)
    for name, g, wv in zip(LAYERS, got, want):
        assert np.array_equal(bits(g), bits(wv)), f"{scan}: {name}"
    assert stats.rays == rays


@pytest.mark.parametrize("mirror", [0, 1], ids=["copy-after", "mirrored-during"])
@pytest.mark.parametrize("batch,frames,scan", [(1, 5, ""), (4, 10, ""), (64, 64, ""), (3, 7, "boxes"), (8, 8, "generic"), (64, 9, "")])
def test_batched_launches_render_the_same_bits(torch_cuda, oracle, batch, frames, scan, mirror):
    """QUASI_CUDA_BATCH=k (default 64) collects render calls and gives every pixel k samples in one launch,
    as k passes over the image ordered by per-tile epochs (pathtrace.cu, Batch): launch by launch (k = 1),
    full batches, a partial one flushed by the readback, a batch longer than the run, and a camera_dirty
    restart in the middle of a batch must all give the oracle's bits. With QUASI_CUDA_MIRROR=1 (the
    default) the batched launch exports every finished tile into the pinned buffers the readback then
    returns; a second readback of the same frame, with the buffers gone, copies from the device."""
    from quasi_b200.headless import headless_host

    w, h = 211, 119
    host = headless_host(w, h, env={"QUASI_CUDA_BATCH": batch, "QUASI_CUDA_SCAN": scan, "QUASI_CUDA_MIRROR": mirror})
    host.render(frames)
    got, stats = host.readback_aov(), host.stats()
    mirrored = any("mirrored during the render" in m for m in host.messages)
    last_launch = frames % batch if frames % batch else min(batch, frames)  # samples in the launch before the readback
    assert mirrored == bool(mirror and last_launch > 1)
    want, rays = oracle.render_frames(oracle.scene(w, h), w, h, frames)
    for name, g, wv in zip(LAYERS, got, want):
        assert np.array_equal(bits(g), bits(wv)), f"batch {batch}: {name}"
    assert stats.rays == rays
    assert stats.kernel_launches == -(-frames // batch)
    again = host.readback_aov()  # the mirror has been handed over: this one is a device-to-host copy of the layers
    for name, g, wv in zip(LAYERS, again, want):
        assert np.array_equal(bits(g), bits(wv)), f"second readback: {name}"
    beauty = host.plugin.readback_numpy()
    assert np.array_equal(bits(beauty), bits(want[0]))
    # restart in the middle of a batch: two calls, then a dirty camera, then three more
    host.render_frame(camera_dirty=True)
    host.render_frame(camera_dirty=False)
    host.render_frame(camera_dirty=True)
    host.render(2)
    beauty = host.plugin.readback_numpy()   # the beauty layer alone, from the mirror when there is one
    got = host.readback_aov()
    host.close()
    want3, _ = oracle.render_frames(oracle.scene(w, h), w, h, 3)
    assert np.array_equal(bits(beauty), bits(want3[0]))
    for name, g, wv in zip(LAYERS, got, want3):
        assert np.array_equal(bits(g), bits(wv)), f"after restart: {name}"


def test_two_instances_with_different_export_settings_do_not_meet(torch_cuda, oracle):
    """The export target of a batched launch travels as a kernel parameter of that launch. (Round 1 kept it in
    a device global: an instance without a mirror then wrote into the freed buffers of an earlier one.)"""
    from quasi_b200.headless import headless_host

    w, h, frames = 160, 90, 6
    a = headless_host(w, h, env={"QUASI_CUDA_BATCH": 3, "QUASI_CUDA_MIRROR": 1})
    b = headless_host(w, h, env={"QUASI_CUDA_BATCH": 2, "QUASI_CUDA_MIRROR": 0})
    a.render(frames)
    b.render(frames)
    got_a = a.readback_aov()
    a.close()
    b.render(2)          # a's buffers are back in the pool or freed by now
    got_b = b.readback_aov()
    b.close()
    want6, _ = oracle.render_frames(oracle.scene(w, h), w, h, frames)
    want8, _ = oracle.render_frames(oracle.scene(w, h), w, h, frames + 2)
    for name, g, wv in zip(LAYERS, got_a, want6):
        assert np.array_equal(bits(g), bits(wv)), name
    for name, g, wv in zip(LAYERS, got_b, want8):
        assert np.array_equal(bits(g), bits(wv)), name
