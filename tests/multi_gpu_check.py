#!/usr/bin/env python
"""Multi-GPU parity check, launched by torchrun on a box with N >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_check.py

Each rank renders its sample shard through the plugin C ABI; the shards are combined by the
push exchange (quasi_b200/multi_gpu.py: tiles pushed to the slice owners by the batched render
kernel, one local reduce kernel per rank, the frame landing in host memory shared by all ranks)
by the fused exchange (the reduce inside the render kernel: arrival counters at the owners, the last
arrival sums) and by NCCL. Rank 0 checks: every shard bit-exact against the oracle; the push result bit-exact
against the oracle's rank-ordered combine, with batched launches (export fused into the render),
with one launch per sample (Q_cuda_push_slices after the fact), and over back-to-back frames of
different length with no host synchronisation in between; the NCCL result within float tolerance,
and the plugin's own layers untouched by it.
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
    from quasi_b200.multi_gpu import shard_exchange

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h = 320, 180
    failures = []

    def render(host, samples):
        host.render_frame(camera_dirty=True)
        for _ in range(samples - 1):
            host.render_frame(camera_dirty=False)

    oracle = Oracle() if rank == 0 else None
    expected = {}

    def want(samples):  # rank 0: the oracle's shards and their rank-ordered combine for `samples` calls per rank
        if samples not in expected:
            u = oracle.scene(w, h)
            shards = [oracle.render_frames(u, w, h, samples, frame_first=g, frame_stride=world)[0] for g in range(world)]
            expected[samples] = (shards, [oracle.combine_shards([s[l] for s in shards]) for l in range(4)])
        return expected[samples]

    def check(name, got, samples, exact=True):
        if rank != 0:
            return
        _, combined = want(samples)
        for l in range(4):
            if exact and not np.array_equal(bits(got[l]), bits(combined[l])):
                failures.append(f"{name}: layer {l}: {(bits(got[l]) != bits(combined[l])).sum()} floats differ from the oracle")
            if not exact and not np.allclose(got[l], combined[l], rtol=2e-6, atol=2e-6):
                failures.append(f"{name}: layer {l} outside the float tolerance")

    for label, env in (("batched", {}), ("one launch per sample", {"QUASI_CUDA_BATCH": 1})):
        host = headless_host(w, h, device_ordinal=local, shard_index=rank, shard_count=world, single_stream=True, env=env)
        # every shard on its own, before any exchange
        render(host, 5)
        mine = host.readback_aov()
        gathered = [None] * world
        dist.all_gather_object(gathered, [m.copy() for m in mine])
        if rank == 0:
            shards, _ = want(5)
            for g in range(world):
                for l in range(4):
                    if not np.array_equal(bits(gathered[g][l]), bits(shards[g][l])):
                        failures.append(f"{label}: shard {g} layer {l} differs from the oracle")

        exchange = shard_exchange(host, dist, mode="push")
        exchange.begin_frame()
        render(host, 5)
        exchange.combine()
        check(f"push, {label}", exchange.result_numpy(), 5)
        # back-to-back frames of different length, nothing between them on the host but the enqueue calls: the
        # device-side flags and the two staging buffers alone order push -> reduce -> next push across the GPUs
        for samples in (3, 7, 2, 5):
            exchange.begin_frame()
            render(host, samples)
            exchange.combine()
        check(f"push, {label}, after four unsynchronised frames", exchange.result_numpy(), 5)
        exchange.begin_frame()
        render(host, 1)  # a frame of ONE sample: nothing is exported by a single-sample launch
        exchange.combine()
        check(f"push, {label}, one-sample frame", exchange.result_numpy(), 1)
        exchange.close()

        # the exchange fused into the render kernel: staged tiles counted at their owners, the last arrival reduces
        exchange = shard_exchange(host, dist, mode="fused")
        exchange.begin_frame(render_calls=5)
        render(host, 5)
        exchange.combine()
        check(f"fused, {label}", exchange.result_numpy(), 5)
        for samples in (3, 7, 2, 5):  # back to back: the device-side barrier alone keeps a buffer from early reuse
            exchange.begin_frame(to_host=samples == 5, render_calls=samples)
            render(host, samples)
            exchange.combine()
        check(f"fused, {label}, after four unsynchronised frames", exchange.result_numpy(), 5)
        exchange.begin_frame(render_calls=1)
        render(host, 1)  # nothing is exported by a single-sample launch: push + reduce with the same buffers
        exchange.combine()
        check(f"fused, {label}, one-sample frame", exchange.result_numpy(), 1)
        exchange.close()

        exchange = shard_exchange(host, dist, mode="nccl")
        exchange.begin_frame()
        render(host, 5)
        exchange.combine()
        check(f"nccl, {label}", exchange.result_numpy(), 5, exact=False)
        after = host.readback_aov()  # the reduce worked on a copy: the plugin's layers are still this rank's mean
        if not all(np.array_equal(bits(a), bits(b)) for a, b in zip(after, mine)):
            failures.append(f"nccl, {label}: rank {rank}'s own layers changed under the combine")
        exchange.close()
        host.close()

    # a frame of several launches (5 calls, 2 per launch, then 1): only the launch with the frame's last call may
    # export; here that is a single-sample launch, so the frame falls back to push + reduce. And 6 calls, 3 per
    # launch: the second launch alone exports and reduces.
    host = headless_host(w, h, device_ordinal=local, shard_index=rank, shard_count=world, single_stream=True,
                         env={"QUASI_CUDA_BATCH": 2})
    exchange = shard_exchange(host, dist, mode="fused")
    for samples in (5, 2, 5):
        exchange.begin_frame(render_calls=samples)
        render(host, samples)
        exchange.combine()
    check("fused, frames of several launches (batch 2)", exchange.result_numpy(), 5)
    exchange.begin_frame(render_calls=2)
    render(host, 2)
    exchange.combine()
    check("fused, batch 2, a frame of one launch after them", exchange.result_numpy(), 2)
    exchange.close()
    host.close()

    all_failures = [None] * world
    dist.all_gather_object(all_failures, failures)
    ok = not any(all_failures)
    if rank == 0:
        for f in sum(all_failures, []):
            print("FAIL", f)
        print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", f"world={world} {w}x{h}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
