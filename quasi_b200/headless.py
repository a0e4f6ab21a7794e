"""Headless host loop over the plugin C ABI, in Python.

Mirrors the reference host's batch mode (`--render N`,
/root/reference/src/quasi/host/main.cpp:120-327): open the backend library,
create the GPU context, load the plugin, then per frame update -> begin_frame ->
fill camera -> render -> end_frame, and finally read the AOVs back. The C++
headless host under quasi_b200/host/ is the deliverable; this twin exists so
that pytest and bench.py can drive the same ABI without a subprocess.
"""
from __future__ import annotations

import ctypes as C

from . import PLUGIN_PATH, abi
from .cuda_context import cuda_context, orbit_camera
from .loader import dynamic_library, loader


class headless_host:
    def __init__(self, width: int, height: int, plugin_path=PLUGIN_PATH, device_ordinal: int = 0,
                 shard_index: int = 0, shard_count: int = 1, present: bool = False, log=None,
                 single_stream: bool = False, env: dict | None = None, keep_accumulation: bool = False):
        """`env`: out-of-band plugin options (QUASI_CUDA_SCENE, QUASI_CUDA_MAX_BOUNCES, ...), set in
        the process environment only while Q_plugin_create runs (the ABI has no slot for them)."""
        self.messages = []
        self._user_log = log

        def _log(_host_data, message):
            text = message.decode() if message else ""
            self.messages.append(text)
            if self._user_log:
                self._user_log(text)

        self.context = cuda_context.create(width, height, device_ordinal, shard_index, shard_count,
                                           present=present, single_stream=single_stream,
                                           keep_accumulation=keep_accumulation)
        self._plugin_path, self._env = plugin_path, env
        self.plugin_context = self.context.make_plugin_context(log=_log)
        self._load()
        self.camera = orbit_camera()
        self.camera.update_position()
        self.frames_rendered = 0

    def _load(self):
        import os

        self.library = dynamic_library.open(self._plugin_path)
        env = self._env
        saved = {k: os.environ.get(k) for k in (env or {})}
        os.environ.update({k: str(v) for k, v in (env or {}).items()})
        try:
            self.plugin = loader.load(self.library, self.plugin_context)
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    def reload(self):
        """What Q::plugin::manager does on a hot reload (manager.hpp:209-268): destroy the
        instance, close the library, open it again, create a new instance on the same context."""
        self.context.synchronize()
        self.plugin.destroy()
        self.library.close()
        self._load()

    def render_frame(self, camera_dirty=None):
        """One iteration of the main loop, main.cpp:243-270. Enqueue only."""
        self.plugin.update(0.0)
        frame = self.context.begin_frame()
        self.camera.fill_camera(frame.camera)
        dirty = self.camera.dirty if camera_dirty is None else camera_dirty
        frame.camera_dirty = 1 if dirty else 0
        self.camera.dirty = False
        self.plugin.render(frame)
        self.context.end_frame(frame)
        self.frames_rendered += 1
        return frame

    def render(self, frames: int):
        for _ in range(frames):
            self.render_frame()

    def readback_aov(self):
        return self.plugin.readback_aov_numpy()

    def stats(self) -> abi.Q_cuda_stats:
        return self.plugin.cuda_stats()

    def close(self):
        self.plugin.destroy()
        p = self.context.persistent
        if p.valid and p.layers:  # a block parked by the destroyed instance belongs to the host
            self.library.get_symbol("Q_cuda_free")(C.c_void_p(p.layers))
            p.valid, p.layers = 0, None
        self.library.close()
