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
