"""Python mirror of Q::plugin::dynamic_library and Q::plugin::loader.

Same names, same resolution order and the same error cases as
/root/reference/src/quasi/plugin/dynamic_library.hpp:67-178 and
/root/reference/src/quasi/plugin/loader.hpp:35-283, so that the parity tests
drive libquasi_cuda.so exactly the way the reference host would. std::expected
becomes an exception carrying the same error enumerator.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from pathlib import Path

import numpy as np

from . import abi


class library_error(enum.Enum):  # dynamic_library.hpp:19-24
    file_not_found = "file not found"
    load_failed = "failed to load library"
    symbol_not_found = "symbol not found"
    not_loaded = "library not loaded"


class loader_error(enum.Enum):  # loader.hpp:38-42
    abi_mismatch = "ABI version mismatch"
    symbol_not_found = "Required symbol not found"
    create_failed = "Plugin creation failed"


class plugin_error(RuntimeError):
    def __init__(self, code):
        super().__init__(code.value)
        self.code = code


class dynamic_library:
    """dlopen(RTLD_NOW | RTLD_LOCAL) wrapper, dynamic_library.hpp:100-117."""

    def __init__(self):
        self._lib = None
        self._path = None

    @classmethod
    def open(cls, path) -> "dynamic_library":
        path = Path(path)
        if not path.exists():
            raise plugin_error(library_error.file_not_found)
        try:
            lib = C.CDLL(str(path), mode=os.RTLD_NOW | os.RTLD_LOCAL)
        except OSError as e:
            err = plugin_error(library_error.load_failed)
            err.detail = str(e)
            raise err
        self = cls()
        self._lib, self._path = lib, path
        return self

    def get_symbol(self, name: str):
        if self._lib is None:
            raise plugin_error(library_error.not_loaded)
        try:
            return getattr(self._lib, name)
        except AttributeError:
            raise plugin_error(library_error.symbol_not_found)

    def has_symbol(self, name: str) -> bool:
        try:
            self.get_symbol(name)
            return True
        except plugin_error:
            return False

    def is_loaded(self) -> bool:
        return self._lib is not None

    def path(self):
        return self._path

    def close(self):
        # ctypes never dlcloses; dropping the handle matches what tests need.
        self._lib = None
        self._path = None


class loader:
    """RAII-style wrapper over a plugin instance, loader.hpp:35-283."""

    def __init__(self):
        self._handle = None
        self._fn = {}
        self._ctx = None  # keeps the Q_plugin_context alive: the plugin stores the pointer

    @classmethod
    def load(cls, library: dynamic_library, context: abi.Q_plugin_context) -> "loader":
        p = cls()
        # required symbols, in the reference's order (loader.hpp:72-106)
        for name in abi.REQUIRED_SYMBOLS:
            try:
                p._fn[name] = library.get_symbol(name)
            except plugin_error:
                raise plugin_error(loader_error.symbol_not_found)
        p._bind_required()
        plugin_abi = p._fn["Q_plugin_abi_version"]()
        if plugin_abi < 1 or plugin_abi > abi.K_PLUGIN_ABI_VERSION:  # loader.hpp:109-112
            raise plugin_error(loader_error.abi_mismatch)
        for name in abi.OPTIONAL_SYMBOLS:  # loader.hpp:115-128
            if library.has_symbol(name):
                p._fn[name] = library.get_symbol(name)
        for name in abi.CUDA_EXTENSION_SYMBOLS:  # not part of the reference ABI
            if library.has_symbol(name):
                p._fn[name] = library.get_symbol(name)
        p._bind_optional()
        p._ctx = context
        handle = p._fn["Q_plugin_create"](C.byref(context) if context is not None else None)
        if not handle:
            raise plugin_error(loader_error.create_failed)
        p._handle = handle
        return p

    def _bind_required(self):
        f = self._fn
        f["Q_plugin_abi_version"].restype = C.c_uint32
        f["Q_plugin_abi_version"].argtypes = []
        f["Q_plugin_get_info"].restype = abi.Q_plugin_info
        f["Q_plugin_get_info"].argtypes = []
        f["Q_plugin_create"].restype = C.c_void_p
        f["Q_plugin_create"].argtypes = [C.POINTER(abi.Q_plugin_context)]
        f["Q_plugin_destroy"].restype = None
        f["Q_plugin_destroy"].argtypes = [C.c_void_p]
        f["Q_plugin_update"].restype = None
        f["Q_plugin_update"].argtypes = [C.c_void_p, C.c_float]
        f["Q_plugin_render"].restype = None
        f["Q_plugin_render"].argtypes = [C.c_void_p, C.POINTER(abi.Q_render_frame)]

    def _bind_optional(self):
        f = self._fn
        sig = {
            "Q_plugin_readback": (abi.Q_readback_result, [C.c_void_p]),
            "Q_plugin_readback_free": (None, [C.POINTER(abi.Q_readback_result)]),
            "Q_plugin_readback_aov": (abi.Q_readback_aov_result, [C.c_void_p]),
            "Q_plugin_readback_aov_free": (None, [C.POINTER(abi.Q_readback_aov_result)]),
            "Q_cuda_plugin_stats": (C.c_int, [C.c_void_p, C.POINTER(abi.Q_cuda_stats)]),
            "Q_cuda_plugin_layers": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(abi.Q_cuda_layers)]),
            "Q_cuda_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                            C.c_uint32, C.c_void_p]),
            "Q_cuda_tonemap_bgra8": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                               C.c_uint32, C.c_void_p]),
            "Q_cuda_combine_shards": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, C.c_uint64,
                                                C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
            "Q_cuda_plugin_set_export": (C.c_int, [C.c_void_p, C.POINTER(abi.Q_cuda_export_table)]),
            "Q_cuda_plugin_set_reduce": (C.c_int, [C.c_void_p, C.POINTER(abi.Q_cuda_reduce_table)]),
            "Q_cuda_push_slices_reduce": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(abi.Q_cuda_reduce_table), C.c_void_p]),
            "Q_cuda_plugin_flush": (C.c_int, [C.c_void_p, C.c_void_p]),
            "Q_cuda_push_slices": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(abi.Q_cuda_export_table), C.c_void_p]),
            "Q_cuda_reduce_slices_sync": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32,
                                                    C.c_uint32, C.c_uint64, C.c_float, C.c_void_p, C.c_void_p]),
            "Q_cuda_host_register": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
            "Q_cuda_host_unregister": (C.c_int, [C.c_void_p]),
            "Q_cuda_memcpy_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
            "Q_cuda_memset_async": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]),
            "Q_cuda_alloc": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p)]),
            "Q_cuda_free": (C.c_int, [C.c_void_p]),
            "Q_cuda_ipc_export": (C.c_int, [C.c_void_p, C.POINTER(abi.Q_cuda_ipc_handle)]),
            "Q_cuda_ipc_open": (C.c_int, [C.POINTER(abi.Q_cuda_ipc_handle), C.POINTER(C.c_void_p)]),
            "Q_cuda_ipc_close": (C.c_int, [C.c_void_p]),
            "Q_cuda_plugin_mesh_view": (C.c_int, [C.c_void_p, C.POINTER(abi.Q_cuda_mesh_view)]),
            "Q_cuda_probe_sincos": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
        }
        for name, (res, args) in sig.items():
            if name in f:
                f[name].restype = res
                f[name].argtypes = args

    # ---- loader.hpp:182-234 ----
    def update(self, delta_time: float):
        if self._handle:
            self._fn["Q_plugin_update"](self._handle, delta_time)

    def render(self, frame: abi.Q_render_frame):
        if self._handle:
            self._fn["Q_plugin_render"](self._handle, C.byref(frame) if frame is not None else None)

    def supports_readback(self) -> bool:
        return "Q_plugin_readback" in self._fn and "Q_plugin_readback_free" in self._fn

    def readback(self) -> abi.Q_readback_result:
        if self._handle and "Q_plugin_readback" in self._fn:
            return self._fn["Q_plugin_readback"](self._handle)
        return abi.Q_readback_result()

    def readback_free(self, result):
        if "Q_plugin_readback_free" in self._fn and result is not None:
            self._fn["Q_plugin_readback_free"](C.byref(result))

    def supports_readback_aov(self) -> bool:
        return "Q_plugin_readback_aov" in self._fn and "Q_plugin_readback_aov_free" in self._fn

    def readback_aov(self) -> abi.Q_readback_aov_result:
        if self._handle and "Q_plugin_readback_aov" in self._fn:
            return self._fn["Q_plugin_readback_aov"](self._handle)
        return abi.Q_readback_aov_result()

    def readback_aov_free(self, result):
        if "Q_plugin_readback_aov_free" in self._fn and result is not None:
            self._fn["Q_plugin_readback_aov_free"](C.byref(result))

    def info(self) -> abi.Q_plugin_info:
        return self._fn["Q_plugin_get_info"]()

    def abi_version(self) -> int:
        return self._fn["Q_plugin_abi_version"]() if "Q_plugin_abi_version" in self._fn else 0

    def is_valid(self) -> bool:
        return bool(self._handle)

    __bool__ = is_valid

    def destroy(self):
        if self._handle:
            self._fn["Q_plugin_destroy"](self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # ---- conveniences on top of the C ABI (numpy copies of readback buffers) ----
    def readback_aov_numpy(self):
        """Four (H, W, 4) float32 arrays (beauty, albedo, normal, depth) or None."""
        rb = self.readback_aov()
        try:
            out = []
            for i in range(abi.Q_AOV_COUNT):
                b = rb.buffers[i]
                if not b.data:
                    return None
                a = np.ctypeslib.as_array(b.data, shape=(b.height, b.width, b.channels))
                out.append(a.copy())
            return out
        finally:
            self.readback_aov_free(rb)

    def readback_numpy(self):
        rb = self.readback()
        try:
            if not rb.data:
                return None
            return np.ctypeslib.as_array(rb.data, shape=(rb.height, rb.width, rb.channels)).copy()
        finally:
            self.readback_free(rb)

    # ---- CUDA-only extensions (quasi_cuda_backend.h) ----
    def cuda_stats(self) -> abi.Q_cuda_stats:
        s = abi.Q_cuda_stats()
        if self._fn["Q_cuda_plugin_stats"](self._handle, C.byref(s)) != 0:
            raise RuntimeError("Q_cuda_plugin_stats failed")
        return s

    def cuda_layers(self, stream_ptr: int) -> abi.Q_cuda_layers:
        l = abi.Q_cuda_layers()
        if self._fn["Q_cuda_plugin_layers"](self._handle, stream_ptr, C.byref(l)) != 0:
            raise RuntimeError("Q_cuda_plugin_layers failed")
        return l

    def cuda_mesh_view(self):
        v = abi.Q_cuda_mesh_view()
        if self._fn["Q_cuda_plugin_mesh_view"](self._handle, C.byref(v)) != 0:
            return None
        return v

    def fn(self, name: str):
        return self._fn[name]
