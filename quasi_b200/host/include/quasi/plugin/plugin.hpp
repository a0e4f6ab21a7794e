/// @file plugin.hpp
/// @brief Umbrella header for the plugin system.
///
/// Same role and names as /root/reference/src/quasi/plugin/plugin.hpp: the C ABI a backend
/// exports (plugin_interface.hpp), the dlopen wrapper (dynamic_library.hpp), the symbol
/// binder with its ABI-version check (loader.hpp) and the hot-reload manager (manager.hpp),
/// which the headless host in this repository actually drives (the reference's own host
/// never wired it in).

#pragma once

#include <quasi/plugin/dynamic_library.hpp>
#include <quasi/plugin/loader.hpp>
#include <quasi/plugin/manager.hpp>
#include <quasi/plugin/plugin_interface.hpp>

namespace Q::plugin {

// version of the host-side plugin module, as the reference numbers it
inline constexpr int k_version_major = 0, k_version_minor = 1, k_version_patch = 0;

}  // namespace Q::plugin
