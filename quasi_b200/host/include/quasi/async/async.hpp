/// @file async.hpp
/// @brief Umbrella header for the coroutine runtime.
///
/// Same role and names as /root/reference/src/quasi/async/async.hpp: lazy tasks (task.hpp), the
/// single-threaded cooperative scheduler that the headless host ticks (scheduler.hpp), the
/// awaitables (awaitables.hpp) and the polling file watcher behind hot reload
/// (file_watcher.hpp). What waits for the GPU in this tree (gpu/cuda/context.hpp: a frame slot
/// becoming free, wait_idle) is built on the same awaitables and polls cudaEventQuery, so a
/// tick never blocks on the device.

#pragma once

#include <quasi/async/awaitables.hpp>
#include <quasi/async/file_watcher.hpp>
#include <quasi/async/scheduler.hpp>
#include <quasi/async/task.hpp>

namespace Q::async {

// version of the coroutine module, as the reference numbers it
inline constexpr int k_version_major = 0, k_version_minor = 1, k_version_patch = 0;

}  // namespace Q::async
