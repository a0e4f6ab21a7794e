/// @file plugin.hpp
/// @brief Umbrella header for the plugin system.

#pragma once

#include <quasi/plugin/dynamic_library.hpp>
#include <quasi/plugin/loader.hpp>
#include <quasi/plugin/manager.hpp>
#include <quasi/plugin/plugin_interface.hpp>

namespace Q::plugin {

inline constexpr int k_version_major = 0;
inline constexpr int k_version_minor = 1;
inline constexpr int k_version_patch = 0;

}  // namespace Q::plugin
