/// @file plugin_interface.hpp
/// @brief C++ view of the plugin C ABI.
///
/// The ABI itself is include/quasi_plugin_abi.h. This header adds what
/// /root/reference/src/quasi/plugin/plugin_interface.hpp:152-197 adds on top of
/// it for C++ hosts: Q::plugin aliases, dlsym names, the ABI version constant and
/// version comparison.

#pragma once

#include <quasi/gpu/types.hpp>

#include <cstddef>
#include <cstdint>

namespace Q::plugin {

using plugin_handle = Q_plugin_handle;
using plugin_version = Q_plugin_version;
using plugin_info = Q_plugin_info;
using plugin_context = Q_plugin_context;
using readback_result = Q_readback_result;
using aov_type = Q_aov_type;
using aov_buffer = Q_aov_buffer;
using readback_aov_result = Q_readback_aov_result;

inline constexpr const char* k_symbol_abi_version = "Q_plugin_abi_version";
inline constexpr const char* k_symbol_get_info = "Q_plugin_get_info";
inline constexpr const char* k_symbol_create = "Q_plugin_create";
inline constexpr const char* k_symbol_destroy = "Q_plugin_destroy";
inline constexpr const char* k_symbol_update = "Q_plugin_update";
inline constexpr const char* k_symbol_render = "Q_plugin_render";
inline constexpr const char* k_symbol_readback = "Q_plugin_readback";
inline constexpr const char* k_symbol_readback_free = "Q_plugin_readback_free";
inline constexpr const char* k_symbol_readback_aov = "Q_plugin_readback_aov";
inline constexpr const char* k_symbol_readback_aov_free = "Q_plugin_readback_aov_free";

inline constexpr uint32_t k_plugin_abi_version = Q_PLUGIN_ABI_VERSION;

[[nodiscard]] constexpr bool operator==(plugin_version a, plugin_version b) noexcept {
    return a.major == b.major && a.minor == b.minor && a.patch == b.patch;
}

[[nodiscard]] constexpr bool operator<(plugin_version a, plugin_version b) noexcept {
    if (a.major != b.major) return a.major < b.major;
    if (a.minor != b.minor) return a.minor < b.minor;
    return a.patch < b.patch;
}

}  // namespace Q::plugin
