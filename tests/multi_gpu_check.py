#!/usr/bin/env python
"""Multi-GPU parity check, launched by torchrun on a box with N >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_check.py

Each rank renders its sample shard through the plugin C ABI; the shards are
combined with the fused peer-memory kernel (p2p) and with NCCL. Rank 0 checks:
every shard bit-exact against the oracle, the p2p result bit-exact against the
oracle's rank-ordered combine, the NCCL result within float tolerance.
(test_cuda_multi_gpu.py runs this under pytest when >= 2 GPUs are visible.)
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    from oracle_lib import Oracle, bits
    from quasi_b200.headless import headless_host
    from quasi_b200.multi_gpu import shard_reducer

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h, per_rank = 320, 180, 5
    host = headless_host(w, h, device_ordinal=local, shard_index=rank, shard_count=world, single_stream=True)
    results = {}
    for mode in ("p2p", "p2p_host", "nccl"):
        reducer = shard_reducer(host, dist, mode=mode)
        host.render_frame(camera_dirty=True)
        for _ in range(per_rank - 1):
            host.render_frame(camera_dirty=False)
        mine = host.readback_aov()           # this rank's shard mean (before the in-place NCCL sum)
        reducer.combine()
        results[mode] = reducer.result_numpy()
        if mode == "p2p":
            # back-to-back frames with no host synchronisation in between: the in-kernel
            # flags alone must order render -> combine -> next render across the GPUs
            for _ in range(3):
                host.render_frame(camera_dirty=True)
                for _ in range(per_rank - 1):
                    host.render_frame(camera_dirty=False)
                reducer.combine()
            again = reducer.result_numpy()
            if rank == 0 and not all(np.array_equal(bits(a), bits(b)) for a, b in zip(again, results[mode])):
                print("FAIL p2p: repeated frames without host barriers differ")
                results["p2p_repeat_failed"] = True
        dist.barrier()
        if rank != 0:
            reducer.close()
        dist.barrier()
        if rank == 0:
            reducer.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, [m.copy() for m in mine])
    ok = True
    if rank == 0:
        oracle = Oracle()
        u = oracle.scene(w, h)
        for l in range(4):
            shards = []
            for g in range(world):
                want, _ = oracle.render_frames(u, w, h, per_rank, frame_first=g, frame_stride=world)
                if not np.array_equal(bits(gathered[g][l]), bits(want[l])):
                    print(f"FAIL shard {g} layer {l} differs from the oracle")
                    ok = False
                shards.append(want[l])
            combined = oracle.combine_shards(shards)
            if not np.array_equal(bits(results["p2p"][l]), bits(combined)):
                print(f"FAIL p2p combine layer {l}: {(bits(results['p2p'][l]) != bits(combined)).sum()} floats differ")
                ok = False
            if not np.array_equal(bits(results["p2p_host"][l]), bits(combined)):
                print(f"FAIL p2p_host combine layer {l}")
                ok = False
            if not np.allclose(results["nccl"][l], combined, rtol=2e-6, atol=2e-6):
                print(f"FAIL nccl combine layer {l}")
                ok = False
        ok = ok and not results.get("p2p_repeat_failed", False)
        print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", f"world={world} {w}x{h} {per_rank} spp/rank")
    host.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
