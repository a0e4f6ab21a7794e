"""Sample-sharded multi-GPU rendering: one process per GPU, one exchange per frame.

Every (pixel, frame index) sample depends only on its seed and the immutable
scene (/root/reference/backends/metal/plugin.mm:295), so GPU r of N renders the
RNG frame indices r, r+N, r+2N, ... for every pixel into its own full-frame
layers (the plugin does this itself from Q_cuda_device_desc.shard_index/count).
The only exchange is the combine: result = (sum over ranks of the per-rank
means) * (1/N). Two implementations of that exchange:

  nccl  torch.distributed.reduce(SUM) of the contiguous 4-layer block to rank 0
        on the frame stream, then the 1/N scale kernel on rank 0. The baseline.
  p2p   the layer blocks of all ranks are mapped into every process with CUDA
        IPC. Each rank runs ONE kernel (Q_cuda_combine_shards_sync on peer
        pointers) that waits on device-side flags until every rank's render has
        finished, loads its 1/N slice of every rank's layers over NVLink in rank
        order, scales, stores the slice straight into rank 0's result buffer and
        signals completion: reduce-scatter + gather fused, no staging copy, no
        host barrier, and unlike NCCL's ring/tree the summation order is fixed,
        so the result is bit-identical to the oracle's rank-ordered combine.
  p2p_host  the same data path with Q_cuda_combine_shards and two host barriers
        per frame instead of the flags (the first form; kept for comparison).

The reference has no multi-GPU path; the oracle for the combine
is oracle/quasi_oracle.c:qo_combine_shards.
"""
from __future__ import annotations

import ctypes as C

from . import abi


def frame_indices(rank: int, world: int, count: int):
    """RNG frame indices rendered by `rank` in its first `count` render calls."""
    return [rank + i * world for i in range(count)]


def slice_bounds(total: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of `total` items owned by `rank`."""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def reduce_mean_to_root(tensor, dist, world: int, dst: int = 0):
    """Backend-agnostic combine on a tensor (gloo on CPU, nccl on GPU):
    sum-reduce to `dst`, then scale by 1/world there. Returns the result on
    `dst`, None elsewhere. Used by the CPU protocol tests and as the semantics
    both GPU paths must reproduce."""
    import torch

    work = tensor.clone()
    dist.reduce(work, dst=dst, op=dist.ReduceOp.SUM)
    if dist.get_rank() != dst:
        return None
    return work * torch.tensor(1.0 / world, dtype=torch.float32).to(work.device)


def device_tensor(ptr: int, n_floats: int, device):
    """A torch view of plugin-owned device memory (CUDA array interface, no copy)."""
    import torch

    class _holder:
        __cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False),
                                    "version": 2, "strides": None}

    return torch.as_tensor(_holder(), device=device)


class shard_reducer:
    """Combines the per-rank accumulation layers into rank 0's result buffer."""

    def __init__(self, host, dist, mode: str = "p2p"):
        import torch

        self.torch, self.host, self.dist, self.mode = torch, host, dist, mode
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = host.context.device
        self.stream = host.context.frame_stream()
        self.fn = host.plugin.fn
        self.n_px = host.context.width * host.context.height
        self.total_float4 = self.n_px * 4  # four layers, contiguous
        self.result_ptr = None            # rank 0: the combined layers (own memory)
        self._opened = []
        self.launches = 0
        self.epoch = 0
        self.flag_ptrs = None
        self._own_flags = None
        if mode in ("p2p", "p2p_host"):
            self._setup_p2p()
        elif mode == "nccl":
            self._setup_nccl()
        else:
            raise ValueError(mode)

    # ---- setup ----
    def _layer_block(self):
        layers = self.host.plugin.cuda_layers(self.stream.cuda_stream)
        return layers.device_ptr[0]

    def _setup_nccl(self):
        self.result = self.torch.empty(self.total_float4 * 4, dtype=self.torch.float32, device=self.device)
        self.result_ptr = self.result.data_ptr()

    def _setup_p2p(self):
        # The plugin allocates its layers at create; render once so the block exists
        # even if a host resized, then export it.
        block = self._layer_block()
        mine = abi.Q_cuda_ipc_handle()
        if self.fn("Q_cuda_ipc_export")(block, C.byref(mine)) != 0:
            raise RuntimeError("Q_cuda_ipc_export(layers) failed")
        result_handle = None
        if self.rank == 0:
            p = C.c_void_p()
            if self.fn("Q_cuda_alloc")(self.total_float4 * 16, C.byref(p)) != 0:
                raise RuntimeError("Q_cuda_alloc(result) failed")
            self.result_ptr = p.value
            rh = abi.Q_cuda_ipc_handle()
            if self.fn("Q_cuda_ipc_export")(self.result_ptr, C.byref(rh)) != 0:
                raise RuntimeError("Q_cuda_ipc_export(result) failed")
            result_handle = bytes(rh.bytes)
        # this rank's flag block for the in-kernel synchronisation, zeroed before anyone can see it
        p = C.c_void_p()
        if self.fn("Q_cuda_alloc")(256, C.byref(p)) != 0:
            raise RuntimeError("Q_cuda_alloc(flags) failed")
        self._own_flags = p.value
        if self.fn("Q_cuda_memset_async")(self._own_flags, 0, 256, self.stream.cuda_stream) != 0:
            raise RuntimeError("Q_cuda_memset_async(flags) failed")
        self.stream.synchronize()
        fh = abi.Q_cuda_ipc_handle()
        if self.fn("Q_cuda_ipc_export")(self._own_flags, C.byref(fh)) != 0:
            raise RuntimeError("Q_cuda_ipc_export(flags) failed")
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, (bytes(mine.bytes), result_handle, bytes(fh.bytes)))
        self.block_ptrs, self.flag_ptrs = [], []
        for g, (layer_handle, _, flag_handle) in enumerate(gathered):
            if g == self.rank:
                self.block_ptrs.append(block)
                self.flag_ptrs.append(self._own_flags)
            else:
                self.block_ptrs.append(self._open(layer_handle))
                self.flag_ptrs.append(self._open(flag_handle))
        self.out_ptr = self.result_ptr if self.rank == 0 else self._open(gathered[0][1])

    def _open(self, raw: bytes) -> int:
        h = abi.Q_cuda_ipc_handle()
        C.memmove(h.bytes, raw, 64)
        p = C.c_void_p()
        if self.fn("Q_cuda_ipc_open")(C.byref(h), C.byref(p)) != 0:
            raise RuntimeError("Q_cuda_ipc_open failed (no peer access between these GPUs?)")
        self._opened.append(p.value)
        return p.value

    # ---- per frame ----
    def combine(self):
        if self.mode == "nccl":
            self._combine_nccl()
        elif self.mode == "p2p":
            self._combine_p2p()
        else:
            self._combine_p2p_host()

    def _combine_nccl(self):
        torch, dist = self.torch, self.dist
        block = device_tensor(self._layer_block(), self.total_float4 * 4, self.device)
        with torch.cuda.stream(self.stream):
            dist.reduce(block, dst=0, op=dist.ReduceOp.SUM)
        if self.rank == 0:
            ptrs = (C.c_void_p * 1)(block.data_ptr())
            rc = self.fn("Q_cuda_combine_shards")(ptrs, 1, self.total_float4, 1.0 / self.world,
                                                  self.result_ptr, None, self.stream.cuda_stream)
            assert rc == 0
            self.launches += 1

    def _combine_p2p(self):
        """One launch per rank, enqueue only: the kernel itself waits for the peers' renders
        and for their slices (combine_shards_sync_kernel)."""
        self.epoch += 1
        lo, hi = slice_bounds(self.total_float4, self.world, self.rank)
        off = lo * 16
        ptrs = (C.c_void_p * self.world)(*[p + off for p in self.block_ptrs])
        flags = (C.c_void_p * self.world)(*self.flag_ptrs)
        rc = self.fn("Q_cuda_combine_shards_sync")(ptrs, flags, self.world, self.rank, self.epoch, hi - lo, 0.0,
                                                   self.out_ptr + off, None, self.stream.cuda_stream)
        assert rc == 0
        self.launches += 1

    def _combine_p2p_host(self):
        # All ranks must have finished rendering before anyone reads their layers,
        # and all slices must have landed before rank 0 uses the result (or anyone
        # restarts accumulation). Host-side barriers; the kernels are one per rank.
        self.stream.synchronize()
        self.dist.barrier()
        lo, hi = slice_bounds(self.total_float4, self.world, self.rank)
        if hi > lo:
            off = lo * 16
            ptrs = (C.c_void_p * self.world)(*[p + off for p in self.block_ptrs])
            rc = self.fn("Q_cuda_combine_shards")(ptrs, self.world, hi - lo, 0.0, self.out_ptr + off,
                                                  None, self.stream.cuda_stream)
            assert rc == 0
            self.launches += 1
        self.stream.synchronize()
        self.dist.barrier()

    def result_numpy(self):
        """Rank 0: the four combined layers as (H, W, 4) float32 arrays."""
        if self.rank != 0:
            return None
        self.stream.synchronize()
        t = device_tensor(self.result_ptr, self.total_float4 * 4, self.device).cpu().numpy()
        h, w = self.host.context.height, self.host.context.width
        return [t[i * self.n_px * 4:(i + 1) * self.n_px * 4].reshape(h, w, 4).copy() for i in range(4)]

    def close(self):
        for p in self._opened:
            self.fn("Q_cuda_ipc_close")(p)
        self._opened = []
        if self.mode != "nccl" and self.rank == 0 and self.result_ptr:
            self.fn("Q_cuda_free")(self.result_ptr)
            self.result_ptr = None
        if self._own_flags:
            self.fn("Q_cuda_free")(self._own_flags)
            self._own_flags = None
