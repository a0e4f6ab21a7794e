/// @file async.hpp
/// @brief Umbrella header for the coroutine runtime.

#pragma once

#include <quasi/async/awaitables.hpp>
#include <quasi/async/file_watcher.hpp>
#include <quasi/async/scheduler.hpp>
#include <quasi/async/task.hpp>

namespace Q::async {

inline constexpr int k_version_major = 0;
inline constexpr int k_version_minor = 1;
inline constexpr int k_version_patch = 0;

}  // namespace Q::async
