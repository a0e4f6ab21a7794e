/// @file gpu.hpp
/// @brief Umbrella header for the GPU layer.
///
/// Mirrors /root/reference/src/quasi/gpu/gpu.hpp so that code written against the reference
/// tree compiles here: it pulls in the backend-neutral POD types of the plugin ABI
/// (types.hpp, with Q_GPU_BACKEND_CUDA appended). The CUDA context that fills those types
/// (devices, streams, frames in flight, pinned drawables) is <quasi/gpu/cuda/context.hpp>;
/// it is not included from here because it needs the CUDA runtime headers, which a unit
/// test of the ABI types should not.

#pragma once

#include <quasi/gpu/types.hpp>
