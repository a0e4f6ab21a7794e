"""ctypes view of include/quasi_plugin_abi.h and include/quasi_cuda_backend.h.

Field order and types follow the C headers, which in turn follow
/root/reference/src/quasi/plugin/plugin_interface.hpp:27-92 and
/root/reference/src/quasi/gpu/types.hpp:14-70.
"""
import ctypes as C

Q_GPU_BACKEND_NONE, Q_GPU_BACKEND_METAL, Q_GPU_BACKEND_VULKAN, Q_GPU_BACKEND_WEBGPU = 0, 1, 2, 3
Q_GPU_BACKEND_CUDA = 4
Q_AOV_BEAUTY, Q_AOV_ALBEDO, Q_AOV_NORMAL, Q_AOV_DEPTH, Q_AOV_COUNT = 0, 1, 2, 3, 4
K_PLUGIN_ABI_VERSION = 3  # plugin_interface.hpp:183

LOG_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)
SHUTDOWN_FN = C.CFUNCTYPE(None, C.c_void_p)


class Q_gpu_context(C.Structure):
    _fields_ = [("backend", C.c_uint32), ("device", C.c_void_p), ("queue", C.c_void_p),
                ("layer", C.c_void_p)]


class Q_camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("target", C.c_float * 3), ("up", C.c_float * 3),
                ("fov", C.c_float)]


class Q_render_frame(C.Structure):
    _fields_ = [("drawable", C.c_void_p), ("command_buffer", C.c_void_p), ("width", C.c_uint32),
                ("height", C.c_uint32), ("camera", Q_camera), ("camera_dirty", C.c_uint32)]


class Q_plugin_version(C.Structure):
    _fields_ = [("major", C.c_uint32), ("minor", C.c_uint32), ("patch", C.c_uint32)]


class Q_plugin_info(C.Structure):
    _fields_ = [("name", C.c_char_p), ("version", Q_plugin_version), ("description", C.c_char_p),
                ("author", C.c_char_p)]


class Q_plugin_context(C.Structure):
    _fields_ = [("viewport_width", C.c_uint32), ("viewport_height", C.c_uint32),
                ("host_data", C.c_void_p), ("gpu", C.POINTER(Q_gpu_context)),
                ("log", LOG_FN), ("request_shutdown", SHUTDOWN_FN)]


class Q_readback_result(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("width", C.c_uint32), ("height", C.c_uint32),
                ("channels", C.c_uint32)]


class Q_aov_buffer(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("width", C.c_uint32), ("height", C.c_uint32),
                ("channels", C.c_uint32)]


class Q_readback_aov_result(C.Structure):
    _fields_ = [("buffers", Q_aov_buffer * Q_AOV_COUNT)]


# ---- quasi_cuda_backend.h ----

class Q_cuda_device_desc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device_ordinal", C.c_int32),
                ("shard_index", C.c_uint32), ("shard_count", C.c_uint32)]


class Q_cuda_persistent_state(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("valid", C.c_uint32), ("layers", C.c_void_p),
                ("bytes", C.c_uint64), ("width", C.c_uint32), ("height", C.c_uint32),
                ("frame_count", C.c_uint32), ("scene_tag", C.c_uint32), ("shard_index", C.c_uint32),
                ("shard_count", C.c_uint32), ("device_ordinal", C.c_int32), ("reserved", C.c_uint32)]


class Q_cuda_drawable(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("pitch_bytes", C.c_uint32), ("reserved", C.c_uint32)]


class Q_cuda_stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("paths", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("frame_count", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("reserved", C.c_uint32), ("nodes_visited", C.c_uint64), ("triangles_tested", C.c_uint64)]


class Q_cuda_mesh_view(C.Structure):
    _fields_ = [("triangles", C.c_void_p), ("nodes", C.c_void_p), ("leaf_order", C.c_void_p),
                ("n_triangles", C.c_uint32), ("n_nodes", C.c_uint32), ("depth", C.c_uint32),
                ("reserved", C.c_uint32)]


class Q_cuda_layers(C.Structure):
    _fields_ = [("device_ptr", C.c_void_p * Q_AOV_COUNT), ("width", C.c_uint32),
                ("height", C.c_uint32), ("frame_count", C.c_uint32), ("reserved", C.c_uint32)]


class Q_cuda_ipc_handle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 64)]


Q_CUDA_MAX_EXPORT_SLICES = 16


class Q_cuda_export_table(C.Structure):  # quasi_cuda_backend.h
    _fields_ = [("struct_size", C.c_uint32), ("slices", C.c_uint32), ("float4_per_slice", C.c_uint32),
                ("reserved", C.c_uint32), ("base", C.c_void_p * Q_CUDA_MAX_EXPORT_SLICES)]


class Q_cuda_reduce_table(C.Structure):  # quasi_cuda_backend.h
    _fields_ = [("struct_size", C.c_uint32), ("ranks", C.c_uint32), ("rank", C.c_uint32), ("slot_stride", C.c_uint32),
                ("slice_lo", C.c_uint32 * (Q_CUDA_MAX_EXPORT_SLICES + 1)), ("render_calls", C.c_uint32),
                ("staging", C.c_void_p * Q_CUDA_MAX_EXPORT_SLICES),
                ("counters", C.c_void_p * Q_CUDA_MAX_EXPORT_SLICES),
                ("out", C.c_void_p * Q_CUDA_MAX_EXPORT_SLICES)]


# Sizes on x86-64 Linux (probed from the reference headers).
assert C.sizeof(Q_plugin_context) == 40
assert C.sizeof(Q_render_frame) == 72
assert C.sizeof(Q_gpu_context) == 32
assert C.sizeof(Q_camera) == 40
assert C.sizeof(Q_readback_result) == 24
assert C.sizeof(Q_readback_aov_result) == 96

REQUIRED_SYMBOLS = (  # loader.hpp:72-106
    "Q_plugin_abi_version", "Q_plugin_get_info", "Q_plugin_create", "Q_plugin_destroy",
    "Q_plugin_update", "Q_plugin_render",
)
OPTIONAL_SYMBOLS = (  # loader.hpp:115-128
    "Q_plugin_readback", "Q_plugin_readback_free", "Q_plugin_readback_aov",
    "Q_plugin_readback_aov_free",
)
CUDA_EXTENSION_SYMBOLS = (  # quasi_cuda_backend.h
    "Q_cuda_plugin_stats", "Q_cuda_plugin_layers", "Q_cuda_accumulate", "Q_cuda_tonemap_bgra8",
    "Q_cuda_combine_shards", "Q_cuda_alloc", "Q_cuda_free", "Q_cuda_ipc_export", "Q_cuda_ipc_open",
    "Q_cuda_ipc_close", "Q_cuda_probe_sincos", "Q_cuda_plugin_mesh_view", "Q_cuda_memset_async",
    "Q_cuda_plugin_set_export", "Q_cuda_plugin_flush", "Q_cuda_push_slices", "Q_cuda_reduce_slices_sync",
    "Q_cuda_host_register", "Q_cuda_host_unregister", "Q_cuda_memcpy_async", "Q_cuda_plugin_set_reduce",
    "Q_cuda_push_slices_reduce",
)
