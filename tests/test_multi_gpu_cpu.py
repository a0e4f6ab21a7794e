"""Host-side logic of the sample-sharded multi-GPU path, on CPU with gloo.

World size 2, no GPU: each rank renders its shard of frame indices with the CPU
oracle (standing in for its GPU), then the ranks run the same combine protocol
bench.py uses (quasi_b200.multi_gpu.reduce_mean_to_root). Checks the shard plan
(union of samples == the single-GPU sample set), the combine against the
oracle's rank-ordered mean, and the slice partition of the push exchange.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle_lib import Oracle
    from quasi_b200.multi_gpu import frame_indices, reduce_mean_to_root

    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = Oracle()
    w, h, per_rank = 48, 27, 3
    idx = frame_indices(rank, world, per_rank)
    assert idx == [rank + i * world for i in range(per_rank)]
    layers, _ = oracle.render_frames(oracle.scene(w, h), w, h, per_rank, frame_first=rank, frame_stride=world,
                                     threads=2)
    block = torch.from_numpy(np.concatenate([l.reshape(-1) for l in layers]))
    result = reduce_mean_to_root(block, dist, world, dst=0)
    gathered = [None] * world
    dist.all_gather_object(gathered, block.numpy())
    if rank == 0:
        np.save(Path(out_dir) / "result.npy", result.numpy())
        np.save(Path(out_dir) / "shards.npy", np.stack(gathered))
    else:
        assert result is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_combine_matches_oracle(tmp_path, oracle):
    import torch.multiprocessing as mp

    world, port = 2, 29517 + os.getpid() % 500
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    result = np.load(tmp_path / "result.npy")
    shards = np.load(tmp_path / "shards.npy")
    want = oracle.combine_shards([shards[0], shards[1]])
    assert np.array_equal(result.view(np.uint32), want.view(np.uint32))  # a+b is order-free for two ranks
    # the union of both shards' samples is frames 0..5: within float tolerance of the 6-frame mean
    w, h = 48, 27
    single, _ = oracle.render_frames(oracle.scene(w, h), w, h, 6)
    single = np.concatenate([l.reshape(-1) for l in single])
    assert np.allclose(result, single, rtol=2e-6, atol=2e-6)


def test_shard_plan_covers_every_frame_index_once():
    from quasi_b200.multi_gpu import frame_indices

    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in frame_indices(r, world, 64))
        assert seen == list(range(64 * world))


def test_slice_partition_is_exact():
    from quasi_b200.multi_gpu import slice_bounds

    for total in (0, 1, 7, 8 * 1920 * 1080, 4 * 1920 * 1080 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [slice_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi >= lo for lo, hi in spans)


def test_tile_row_slices_give_every_tile_of_a_layer_one_owner():
    """The fused exchange's cut (multi_gpu.py:tile_row_slices, the same formula in host/src/main.cpp and in the
    emulator harness): the slices cover the four layers exactly, are as even as rows of tiles allow, and every
    boundary falls on a row of 8x4 tiles of one layer, so the 32 pixels of a tile in a layer never straddle two
    owners (the kernel counts arrivals per tile and layer at ONE owner)."""
    sys.path.insert(0, str(ROOT))
    from quasi_b200.multi_gpu import tile_row_slices

    for w, h in ((1920, 1080), (3840, 2160), (320, 180), (61, 35), (33, 17), (8, 4), (1, 1)):
        for world in (2, 3, 4, 5, 8, 16):
            lo = tile_row_slices(w, h, world)
            assert len(lo) == world + 1 and lo[0] == 0 and lo[-1] == 4 * w * h
            assert all(a <= b for a, b in zip(lo, lo[1:]))
            tiles_y = (h + 3) // 4
            for x in lo[:-1]:
                assert x % w == 0 and ((x // w) % h) % 4 == 0           # starts a row of tiles of one layer
            sizes = [b - a for a, b in zip(lo, lo[1:])]
            rows = [-(-(s + 1) * 4 * tiles_y // world) - (-(-s * 4 * tiles_y // world)) for s in range(world)]
            assert max(rows) - min(rows) <= 1                          # rows of tiles per slice: as even as it gets
            # every unit (tile, layer) lies inside one slice
            for layer in range(4):
                for ty in range(tiles_y):
                    first = (layer * h + ty * 4) * w
                    last = (layer * h + min(h, ty * 4 + 4)) * w - 1
                    owner = max(s for s in range(world) if lo[s] <= first)
                    assert last < lo[owner + 1] or sizes[owner] == 0, (w, h, world, layer, ty)
