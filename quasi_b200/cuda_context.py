"""Python mirror of the host-side GPU context for the CUDA backend.

Shape follows Q::gpu::metal::context
(/root/reference/src/quasi/gpu/metal/context.hpp:54-103): create / gpu() /
begin_frame / end_frame / resize. It owns the device descriptor, the primary
stream, one stream per frame in flight (3, context.mm:16) and, optionally, a
BGRA8 presentation buffer. torch is used only for streams, events and device
memory; no pixel is computed here. The C++ twin is quasi_b200/host/.
"""
from __future__ import annotations

import ctypes as C
import math

from . import abi

K_MAX_FRAMES_IN_FLIGHT = 3  # context.mm:16


class orbit_camera:
    """Host camera controller defaults, /root/reference/src/quasi/host/main.cpp:37-116."""

    def __init__(self):
        self.target = [0.0, 1.0, 0.0]
        self.distance = 3.5
        self.azimuth = 0.0
        self.elevation = 0.0
        self.fov = 40.0
        self.position = [0.0, 1.0, 3.5]
        self.dirty = True

    def update_position(self):  # main.cpp:52-63
        f32 = lambda x: C.c_float(x).value
        max_elev = 1.5
        self.elevation = max(-max_elev, min(max_elev, self.elevation))
        cos_elev = f32(math.cos(self.elevation))
        self.position[0] = f32(self.target[0] + self.distance * f32(math.sin(self.azimuth)) * cos_elev)
        self.position[1] = f32(self.target[1] + self.distance * f32(math.sin(self.elevation)))
        self.position[2] = f32(self.target[2] + self.distance * f32(math.cos(self.azimuth)) * cos_elev)

    def fill_camera(self, cam: abi.Q_camera):  # main.cpp:103-115
        for i in range(3):
            cam.position[i] = self.position[i]
            cam.target[i] = self.target[i]
        cam.up[0], cam.up[1], cam.up[2] = 0.0, 1.0, 0.0
        cam.fov = self.fov


class cuda_context:
    """Owns device, streams and the presentable framebuffer for one GPU."""

    def __init__(self):
        self._gpu = abi.Q_gpu_context()
        self._desc = abi.Q_cuda_device_desc()
        self._streams = []
        self._events = []
        self._frame_slot = 0
        self._drawable = None
        self._drawable_tensor = None
        self.width = 0
        self.height = 0
        self.device = None

    @classmethod
    def create(cls, width: int, height: int, device_ordinal: int = 0, shard_index: int = 0,
               shard_count: int = 1, present: bool = False,
               single_stream: bool = False, keep_accumulation: bool = False) -> "cuda_context":
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device not found")  # context_error::device_not_found
        self = cls()
        self._torch = torch
        self.device = torch.device("cuda", device_ordinal)
        torch.cuda.set_device(self.device)
        self._desc.struct_size = C.sizeof(abi.Q_cuda_device_desc)
        self._desc.device_ordinal = device_ordinal
        self._desc.shard_index = shard_index
        self._desc.shard_count = shard_count
        self._queue = torch.cuda.Stream(self.device)
        if single_stream:
            # every frame on one stream (still at most 3 in flight): lets a caller
            # bracket a run of frames with CUDA events on that stream
            self._streams = [torch.cuda.Stream(self.device)] * K_MAX_FRAMES_IN_FLIGHT
        else:
            self._streams = [torch.cuda.Stream(self.device) for _ in range(K_MAX_FRAMES_IN_FLIGHT)]
        self._events = [None] * K_MAX_FRAMES_IN_FLIGHT
        self._gpu.backend = abi.Q_GPU_BACKEND_CUDA
        self._gpu.device = C.cast(C.pointer(self._desc), C.c_void_p)
        self._gpu.queue = self._queue.cuda_stream
        # Q_cuda_persistent_state: lets the accumulation outlive a plugin instance
        self.persistent = abi.Q_cuda_persistent_state()
        self.persistent.struct_size = C.sizeof(abi.Q_cuda_persistent_state)
        self._gpu.layer = C.cast(C.pointer(self.persistent), C.c_void_p) if keep_accumulation else None
        self._present = present
        self.resize(width, height)
        return self

    def gpu(self) -> abi.Q_gpu_context:
        return self._gpu

    def queue(self):
        return self._queue

    def resize(self, width: int, height: int):
        self.width, self.height = width, height
        if self._present:
            torch = self._torch
            self._drawable_tensor = torch.zeros((height, width, 4), dtype=torch.uint8, device=self.device)
            self._drawable = abi.Q_cuda_drawable(self._drawable_tensor.data_ptr(), 4 * width, 0)

    def begin_frame(self) -> abi.Q_render_frame:
        """Waits until fewer than 3 frames are in flight, then hands out a stream."""
        slot = self._frame_slot
        ev = self._events[slot]
        if ev is not None:
            ev.synchronize()  # the semaphore wait of context.mm:138
        frame = abi.Q_render_frame()
        frame.drawable = C.cast(C.pointer(self._drawable), C.c_void_p) if self._drawable else None
        frame.command_buffer = self._streams[slot].cuda_stream
        frame.width = self.width
        frame.height = self.height
        frame.camera_dirty = 0
        frame._slot = slot
        return frame

    def end_frame(self, frame: abi.Q_render_frame):
        """Marks the end of the frame's GPU work (presentDrawable + commit in Metal)."""
        slot = self._frame_slot
        ev = self._torch.cuda.Event()
        ev.record(self._streams[slot])
        self._events[slot] = ev
        self._frame_slot = (slot + 1) % K_MAX_FRAMES_IN_FLIGHT

    def frame_stream(self, slot: int = 0):
        return self._streams[slot]

    def drawable_tensor(self):
        return self._drawable_tensor

    def synchronize(self):
        self._torch.cuda.synchronize(self.device)

    def make_plugin_context(self, log=None, request_shutdown=None) -> abi.Q_plugin_context:
        ctx = abi.Q_plugin_context()
        ctx.viewport_width = self.width
        ctx.viewport_height = self.height
        ctx.host_data = None
        ctx.gpu = C.pointer(self._gpu)
        self._log_cb = abi.LOG_FN(log if log else (lambda _h, _m: None))
        self._shutdown_cb = abi.SHUTDOWN_FN(request_shutdown if request_shutdown else (lambda _h: None))
        ctx.log = self._log_cb
        ctx.request_shutdown = self._shutdown_cb
        return ctx
