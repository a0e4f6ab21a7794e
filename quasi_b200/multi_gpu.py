"""Sample-sharded multi-GPU rendering: one process per GPU, one exchange per frame.

Every (pixel, frame index) sample depends only on its seed and the immutable
scene (/root/reference/backends/metal/plugin.mm:295), so GPU r of N renders the
RNG frame indices r, r+N, r+2N, ... for every pixel into its own full-frame
layers (the plugin does this itself from Q_cuda_device_desc.shard_index/count).
The only exchange is the combine: result = (sum over ranks of the per-rank
means) * (1/N). Three implementations of that exchange:

  push  reduce-scatter by peer stores, fused into the render. The contiguous
        4-layer block is cut into N slices; rank s owns slice s and a staging area of N
        slots (CUDA IPC, mapped by every rank). A rank's batched render kernel sends every
        8x4 tile that has received its last sample to the owners of its four layers' slices
        over NVLink, 128 bytes at a time, while the rest of the image is still being
        rendered (Q_cuda_plugin_set_export; Q_cuda_push_slices does the same after the
        fact for a frame whose last launch was a single sample). What follows the render
        is ONE local kernel per rank (Q_cuda_reduce_slices_sync): it waits on device-side
        flags until every rank's frame is complete, sums the N staged copies of its slice
        in rank order, scales, and stores the slice straight into pinned host memory that
        all ranks share, so the finished frame reaches the host over N PCIe links at once
        instead of through rank 0's. No host barrier, no staging copy, no NCCL, and unlike
        NCCL's ring/tree the summation order is fixed: the result is bit-identical to the
        oracle's rank-ordered combine.
  fused (default) the whole exchange inside the render kernel (Q_cuda_plugin_set_reduce). Slices are cut along rows of
        8x4 tiles, so that a tile's 32 pixels in one layer (a unit) have one owner. A warp that has stored a
        finished tile at the owners adds one to the units' arrival counters there (system-scope atomics over
        NVLink); the warp whose arrival is the N-th for a unit, on whichever GPU it runs, loads the N staged
        copies, sums them in rank order, scales, stores the unit into the shared host frame and puts the counter back. The reduce and
        the frame's way to the host overlap the render of every rank; what follows the render is only a
        device-side barrier (so that a staging buffer is not reused while a peer still reads it). Bit-identical
        to the oracle's combine like `push`. A frame whose last launch is a single sample falls back to
        push + reduce with the same buffers.
  nccl  torch.distributed.reduce(SUM) of a copy of the block to rank 0 on the frame
        stream, the 1/N scale kernel on rank 0, then rank 0 copies the frame to the
        host. The baseline.

The reference has no multi-GPU path; the oracle for the combine
is oracle/quasi_oracle.c:qo_combine_shards.
"""
from __future__ import annotations

import ctypes as C
import time

from . import abi


def frame_indices(rank: int, world: int, count: int):
    """RNG frame indices rendered by `rank` in its first `count` render calls."""
    return [rank + i * world for i in range(count)]


def slice_bounds(total: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of `total` items owned by `rank`."""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def tile_row_slices(width: int, height: int, world: int):
    """slice_lo[0..world] of the fused exchange: the 4-layer block (float4 index = (layer * height + y) * width + x)
    cut into `world` slices along rows of 8x4 tiles, as even as that allows (4 * ceil(height / 4) rows of tiles
    over the four layers; slice s starts at row ceil(s * rows / world))."""
    tiles_y = (height + 3) // 4
    rows = 4 * tiles_y
    lo = []
    for s in range(world + 1):
        row = -(-s * rows // world)
        lo.append(4 * width * height if s == world else ((row // tiles_y) * height + (row % tiles_y) * 4) * width)
    return lo


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


class host_frame:
    """The finished frame in host memory every rank has mapped: a POSIX shared-memory segment created
    by rank 0 (4 layers of H x W x 4 float32, then one uint32 per rank and one for rank 0's reader:
    'my slice of frame e has landed' / 'frame e has been looked at')."""

    def __init__(self, dist, n_floats: int):
        from multiprocessing import shared_memory

        import numpy as np

        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.data_bytes = n_floats * 4
        size = self.data_bytes + 4096
        name = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=size)
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=0)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
        self.floats = np.frombuffer(self.shm.buf, dtype=np.float32, count=n_floats)
        self.flags = np.frombuffer(self.shm.buf, dtype=np.uint32, count=64, offset=self.data_bytes)
        if self.rank == 0:
            self.flags[:] = 0
        dist.barrier()
        self._anchor = C.c_char.from_buffer(self.shm.buf)
        self.address = C.addressof(self._anchor)

    def landed(self, epoch: int):
        """This rank's slice of frame `epoch` is in the segment."""
        self.flags[self.rank] = epoch

    def wait_all_landed(self, epoch: int, timeout_s: float = 60.0):
        """Rank 0: spins until every rank's slice of frame `epoch` is in the segment."""
        t0 = time.perf_counter()
        while int(self.flags[:self.world].min()) < epoch:
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError(f"frame {epoch}: slices landed {self.flags[:self.world].tolist()}")

    def consumed(self, epoch: int):
        """Rank 0: whoever reads the segment is done with frame `epoch`; its memory may be overwritten."""
        self.flags[32] = epoch

    def wait_consumed(self, epoch: int, timeout_s: float = 60.0):
        t0 = time.perf_counter()
        while int(self.flags[32]) < epoch:
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError(f"frame {epoch} was never released by rank 0")

    def close(self):
        self.floats = self.flags = self._anchor = None
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass


class shard_exchange:
    """Combines the per-rank accumulation layers of one frame: begin_frame() before the frame's render
    calls, combine() after them (enqueue only), finish_frame() when the frame is wanted on the host."""

    def __init__(self, host, dist, mode: str = "fused", host_path: str | None = None):
        """`host_path` (push mode): how a reduced slice reaches the shared host frame: "store", the reduce kernel
        writes it there itself (zero-copy stores over PCIe), or "dma", the kernel writes device memory and a
        copy engine moves the slice (the default: QUASI_EXCHANGE_HOST_PATH overrides)."""
        import os

        import torch

        self.host_path = host_path or os.environ.get("QUASI_EXCHANGE_HOST_PATH", "dma")

        if mode not in ("push", "fused", "nccl"):
            raise ValueError(mode)
        self.torch, self.host, self.dist, self.mode = torch, host, dist, mode
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = host.context.device
        self.stream = host.context.frame_stream()
        self.fn = host.plugin.fn
        self.handle = host.plugin._handle
        self.n_px = host.context.width * host.context.height
        self.total_float4 = self.n_px * 4  # four layers, contiguous
        self.per = (self.total_float4 + self.world - 1) // self.world
        self.lo, self.hi = slice_bounds(self.total_float4, self.world, self.rank)
        self.launches = 0
        self.epoch = 0
        self._handed_out = 0  # the frame finish_frame() last returned and release_frame() has not yet been called for
        self._opened = []
        self._own = []
        self.frame = None
        if mode == "push":
            self._setup_push()
        elif mode == "fused":
            self._setup_fused()
        else:
            self._setup_nccl()

    # ---- setup ----
    def _layer_block(self):
        """Device pointer of this rank's contiguous 4-layer block; launches what the plugin has collected."""
        layers = self.host.plugin.cuda_layers(self.stream.cuda_stream)
        return layers.device_ptr[0]

    def _alloc(self, n_bytes: int) -> int:
        p = C.c_void_p()
        if self.fn("Q_cuda_alloc")(n_bytes, C.byref(p)) != 0:
            raise RuntimeError(f"Q_cuda_alloc({n_bytes}) failed")
        self._own.append(p.value)
        return p.value

    def _export(self, ptr: int) -> bytes:
        h = abi.Q_cuda_ipc_handle()
        if self.fn("Q_cuda_ipc_export")(ptr, C.byref(h)) != 0:
            raise RuntimeError("Q_cuda_ipc_export failed")
        return bytes(h.bytes)

    def _open(self, raw: bytes) -> int:
        h = abi.Q_cuda_ipc_handle()
        C.memmove(h.bytes, raw, 64)
        p = C.c_void_p()
        if self.fn("Q_cuda_ipc_open")(C.byref(h), C.byref(p)) != 0:
            raise RuntimeError("Q_cuda_ipc_open failed (no peer access between these GPUs?)")
        self._opened.append(p.value)
        return p.value

    def _setup_nccl(self):
        self.result = self.torch.empty(self.total_float4 * 4, dtype=self.torch.float32, device=self.device)
        self.pinned = self.torch.empty(self.total_float4 * 4, dtype=self.torch.float32).pin_memory() \
            if self.rank == 0 else None

    def _setup_push(self):
        # staging: two frames (parity) x `world` slots x `per` float4; flags: one block, zeroed before anyone sees it
        slot_bytes = self.per * 16
        self.staging = self._alloc(2 * self.world * slot_bytes)
        self.device_slice = self._alloc(slot_bytes)  # combine(to_host=False): the reduced slice, kept on this GPU
        self.flags = self._alloc(256)
        if self.fn("Q_cuda_memset_async")(self.flags, 0, 256, self.stream.cuda_stream) != 0:
            raise RuntimeError("Q_cuda_memset_async(flags) failed")
        self.stream.synchronize()
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, (self._export(self.staging), self._export(self.flags)))
        self.staging_ptrs, self.flag_ptrs = [], []
        for g, (staging_handle, flag_handle) in enumerate(gathered):
            mine = g == self.rank
            self.staging_ptrs.append(self.staging if mine else self._open(staging_handle))
            self.flag_ptrs.append(self.flags if mine else self._open(flag_handle))
        # where rank g's staging expects THIS rank's copy of slice g, per parity
        self.tables = []
        for parity in range(2):
            t = abi.Q_cuda_export_table()
            t.struct_size = C.sizeof(abi.Q_cuda_export_table)
            t.slices, t.float4_per_slice = self.world, self.per
            for g in range(self.world):
                t.base[g] = self.staging_ptrs[g] + (parity * self.world + self.rank) * slot_bytes
            self.tables.append(t)
        # the finished frame: host memory shared by all ranks, pinned by each
        self.frame = host_frame(self.dist, self.total_float4 * 4)
        dev = C.c_void_p()
        if self.fn("Q_cuda_host_register")(self.frame.address, self.frame.data_bytes, C.byref(dev)) != 0:
            raise RuntimeError("Q_cuda_host_register(frame) failed")
        self.frame_device_ptr = dev.value

    def _setup_fused(self):
        w, h = self.host.context.width, self.host.context.height
        self.slice_lo = tile_row_slices(w, h, self.world)
        self.lo, self.hi = self.slice_lo[self.rank], self.slice_lo[self.rank + 1]
        self.per = max(1, max(b - a for a, b in zip(self.slice_lo, self.slice_lo[1:])))  # slot stride, float4
        slot_bytes = self.per * 16
        n_tiles = ((w + 7) // 8) * ((h + 3) // 4)
        counter_bytes = 4 * n_tiles * 4
        self.staging = self._alloc(2 * self.world * slot_bytes)
        self.counters = self._alloc(2 * counter_bytes)
        self.device_slice = self._alloc(slot_bytes)  # begin_frame(to_host=False): the finished slice stays on its owner
        self.flags = self._alloc(256)
        for ptr, n in ((self.flags, 256), (self.counters, 2 * counter_bytes)):
            if self.fn("Q_cuda_memset_async")(ptr, 0, n, self.stream.cuda_stream) != 0:
                raise RuntimeError("Q_cuda_memset_async failed")
        self.stream.synchronize()
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, tuple(self._export(p) for p in (self.staging, self.counters, self.flags,
                                                                              self.device_slice)))
        mapped = []
        for g, handles in enumerate(gathered):
            own = (self.staging, self.counters, self.flags, self.device_slice)
            mapped.append(own if g == self.rank else tuple(self._open(x) for x in handles))
        self.staging_ptrs = [m[0] for m in mapped]
        self.flag_ptrs = [m[2] for m in mapped]
        self.frame = host_frame(self.dist, self.total_float4 * 4)
        dev = C.c_void_p()
        if self.fn("Q_cuda_host_register")(self.frame.address, self.frame.data_bytes, C.byref(dev)) != 0:
            raise RuntimeError("Q_cuda_host_register(frame) failed")
        self.frame_device_ptr = dev.value
        self.reduce_tables = {}
        for parity in range(2):
            for to_host in (False, True):
                t = abi.Q_cuda_reduce_table()
                t.struct_size = C.sizeof(abi.Q_cuda_reduce_table)
                t.ranks, t.rank, t.slot_stride = self.world, self.rank, self.per
                for s in range(self.world + 1):
                    t.slice_lo[s] = self.slice_lo[s]
                for s in range(self.world):
                    t.staging[s] = mapped[s][0] + parity * self.world * slot_bytes
                    t.counters[s] = mapped[s][1] + parity * counter_bytes
                    t.out[s] = self.frame_device_ptr + self.slice_lo[s] * 16 if to_host else mapped[s][3]
                self.reduce_tables[parity, to_host] = t

    # ---- per frame ----
    def begin_frame(self, to_host: bool = True, render_calls: int = 0):
        """Before the frame's first render call: from here on the batched launches push their tiles. Fused mode,
        where the frame leaves during the render: `to_host`, whether the finished frame goes to the shared host
        frame or stays on the GPUs, slice by slice (see combine); `render_calls`, how many render calls the frame
        will make on every rank (needed when they may not fit one batched launch: only the last launch exports)."""
        self.epoch += 1
        if self.mode == "fused":
            self._to_host = to_host
            if to_host and self._handed_out:
                self.frame.wait_consumed(self._handed_out)  # the segment still holds a frame rank 0 is looking at
                self._handed_out = 0
            table = self.reduce_tables[self.epoch & 1, to_host]
            table.render_calls = render_calls
            rc = self.fn("Q_cuda_plugin_set_reduce")(self.handle, C.byref(table))
            assert rc == 0
        elif self.mode == "push":
            rc = self.fn("Q_cuda_plugin_set_export")(self.handle, C.byref(self.tables[self.epoch & 1]))
            assert rc == 0
        elif self.epoch == 1:
            # No rank reads its own layers back in this mode: a table of no slices keeps the plugin's readback mirror
            # from sending every rank's whole frame to the host during the render (eight GPUs doing that share one
            # host memory: 9 ms per 1080p frame on the 8 x B200 box).
            none = abi.Q_cuda_export_table()
            none.struct_size = C.sizeof(abi.Q_cuda_export_table)
            rc = self.fn("Q_cuda_plugin_set_export")(self.handle, C.byref(none))
            assert rc == 0

    def combine(self, to_host: bool = True):
        """After the frame's last render call; enqueue only. `to_host` False: the reduced slice stays in this GPU's
        memory (nothing crosses PCIe: what bench.py's device-timed figure wants); True: it goes to the shared
        host frame, which finish_frame() then waits for."""
        if self.mode == "nccl":
            return self._combine_nccl()
        if self.mode == "fused":
            return self._combine_fused()
        slot_bytes = self.per * 16
        exported = self.fn("Q_cuda_plugin_flush")(self.handle, self.stream.cuda_stream)
        assert exported >= 0
        if exported == 0:  # the frame's last launch was a single sample: send the layers now
            block = self._layer_block()
            rc = self.fn("Q_cuda_push_slices")(block, self.total_float4, C.byref(self.tables[self.epoch & 1]),
                                               self.stream.cuda_stream)
            assert rc == 0
            self.launches += 1
        if to_host and self._handed_out:
            self.frame.wait_consumed(self._handed_out)  # the segment still holds a frame rank 0 is looking at
            self._handed_out = 0
        flags = (C.c_void_p * self.world)(*self.flag_ptrs)
        mine = self.staging + (self.epoch & 1) * self.world * slot_bytes
        direct = to_host and self.host_path == "store"
        out = self.frame_device_ptr + self.lo * 16 if direct else self.device_slice
        rc = self.fn("Q_cuda_reduce_slices_sync")(mine, self.per, flags, self.world, self.rank, self.epoch,
                                                  self.hi - self.lo, 0.0, out, self.stream.cuda_stream)
        assert rc == 0
        self.launches += 1
        if to_host and not direct and self.hi > self.lo:
            rc = self.fn("Q_cuda_memcpy_async")(self.frame.address + self.lo * 16, self.device_slice,
                                                (self.hi - self.lo) * 16, self.stream.cuda_stream)
            assert rc == 0

    def _combine_fused(self):
        """The frame was reduced, and sent where it is wanted, by the render launches themselves; what is enqueued here
        is the device-side barrier between frames (Q_cuda_reduce_slices_sync over no elements: arrive, wait for every
        rank), which keeps frame e + 2 out of the staging buffer and counters of frame e while a peer's launch of
        frame e may still read them. A frame whose last launch exported nothing (a single sample) goes the `push`
        way instead, with the same staging buffer: Q_cuda_push_slices_reduce, then the reduce proper."""
        table = self.reduce_tables[self.epoch & 1, self._to_host]
        exported = self.fn("Q_cuda_plugin_flush")(self.handle, self.stream.cuda_stream)
        assert exported >= 0
        flags = (C.c_void_p * self.world)(*self.flag_ptrs)
        mine = self.staging + (self.epoch & 1) * self.world * self.per * 16
        n = 0
        if exported == 0:
            rc = self.fn("Q_cuda_push_slices_reduce")(self._layer_block(), self.total_float4, C.byref(table),
                                                      self.stream.cuda_stream)
            assert rc == 0
            self.launches += 1
            n = self.hi - self.lo
        out = self.frame_device_ptr + self.lo * 16 if self._to_host else self.device_slice
        rc = self.fn("Q_cuda_reduce_slices_sync")(mine, self.per, flags, self.world, self.rank, self.epoch, n, 0.0, out,
                                                  self.stream.cuda_stream)
        assert rc == 0
        self.launches += 1

    def _combine_nccl(self):
        torch, dist = self.torch, self.dist
        block = device_tensor(self._layer_block(), self.total_float4 * 4, self.device)
        with torch.cuda.stream(self.stream):
            self.result.copy_(block)  # the plugin's layers stay what they are: the next call may accumulate onto them
            dist.reduce(self.result, dst=0, op=dist.ReduceOp.SUM)
        if self.rank == 0:
            ptrs = (C.c_void_p * 1)(self.result.data_ptr())
            rc = self.fn("Q_cuda_combine_shards")(ptrs, 1, self.total_float4, 1.0 / self.world,
                                                  self.result.data_ptr(), None, self.stream.cuda_stream)
            assert rc == 0
            self.launches += 1

    def finish_frame(self):
        """Blocks until this rank's part of the frame is on the host; on rank 0, until the whole frame is.
        Returns, on rank 0, the four layers as one float32 array (a view: valid until the next combine)."""
        if self.mode == "nccl":
            if self.rank != 0:
                self.stream.synchronize()
                return None
            with self.torch.cuda.stream(self.stream):
                self.pinned.copy_(self.result, non_blocking=True)
            self.stream.synchronize()
            return self.pinned.numpy()
        self.stream.synchronize()
        self.frame.landed(self.epoch)
        self._handed_out = self.epoch
        if self.rank != 0:
            return None
        self.frame.wait_all_landed(self.epoch)
        return self.frame.floats

    def release_frame(self):
        """Rank 0 is done with what finish_frame returned."""
        if self.mode in ("push", "fused") and self.rank == 0:
            self.frame.consumed(self.epoch)

    def result_numpy(self):
        """Rank 0: the four combined layers as (H, W, 4) float32 arrays (copies). Collective: every rank calls it."""
        flat = self.finish_frame()
        out = None
        if self.rank == 0:
            h, w = self.host.context.height, self.host.context.width
            out = [flat[i * self.n_px * 4:(i + 1) * self.n_px * 4].reshape(h, w, 4).copy() for i in range(4)]
        self.release_frame()
        return out

    def close(self):
        self.fn("Q_cuda_plugin_set_export")(self.handle, None)
        if self.mode == "fused":
            self.fn("Q_cuda_plugin_set_reduce")(self.handle, None)
        peer = self.mode in ("push", "fused")
        if peer:
            self.stream.synchronize()
            self.fn("Q_cuda_host_unregister")(self.frame.address)
            self.dist.barrier()  # nobody still pushes into, or waits on, memory that is about to go
        for p in self._opened:
            self.fn("Q_cuda_ipc_close")(p)
        self._opened = []
        if peer:
            self.dist.barrier()
        for p in self._own:
            self.fn("Q_cuda_free")(p)
        self._own = []
        if self.frame is not None:
            self.frame.close()
            self.frame = None
