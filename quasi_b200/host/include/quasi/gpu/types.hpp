/// @file types.hpp
/// @brief C++ names for the GPU context types of the plugin C ABI.
///
/// The struct definitions live in include/quasi_plugin_abi.h (layout-identical to
/// /root/reference/src/quasi/gpu/types.hpp:14-70); this header adds the Q::gpu
/// aliases the reference exposes (types.hpp:74-90) plus the CUDA backend constant.

#pragma once

#include <quasi_cuda_backend.h>

namespace Q::gpu {

using gpu_backend = Q_gpu_backend;
using gpu_context = Q_gpu_context;
using camera_data = Q_camera;
using render_frame = Q_render_frame;

inline constexpr gpu_backend k_backend_none = Q_GPU_BACKEND_NONE;
inline constexpr gpu_backend k_backend_metal = Q_GPU_BACKEND_METAL;
inline constexpr gpu_backend k_backend_vulkan = Q_GPU_BACKEND_VULKAN;
inline constexpr gpu_backend k_backend_webgpu = Q_GPU_BACKEND_WEBGPU;
inline constexpr gpu_backend k_backend_cuda = Q_GPU_BACKEND_CUDA;

}  // namespace Q::gpu
