/// @file exr_writer.hpp
/// @brief Writes readback buffers as OpenEXR files.
///
/// Same API, channel names and pixel types as
/// /root/reference/src/quasi/io/exr_writer.hpp:15-52 and exr_writer.cpp:17-185.
/// The reference links OpenEXR 3.4.7, which is not available to this build, so
/// the file format is written directly (scanline image, ZIP compression, zlib).

#pragma once

#include <quasi/plugin/plugin_interface.hpp>

#include <expected>
#include <filesystem>

namespace Q::io {

enum class exr_error {
    invalid_data,     ///< Null data or zero dimensions.
    write_failed,     ///< Encoding or file I/O failed.
    directory_error,  ///< Could not create the output directory.
};

[[nodiscard]] constexpr const char* to_string(exr_error e) noexcept {
    switch (e) {
        case exr_error::invalid_data: return "Invalid readback data";
        case exr_error::write_failed: return "EXR write failed";
        case exr_error::directory_error: return "Could not create output directory";
    }
    return "Unknown error";
}

/// RGBA float32 readback -> half R, G, B, A.
[[nodiscard]] std::expected<void, exr_error> write_exr(const std::filesystem::path& path,
                                                       const Q_readback_result& result);

/// AOV readback -> half R, G, B, A (beauty), half albedo.R/G/B, half normal.X/Y/Z,
/// float depth.Z (channel 0 of the depth buffer). Absent layers are skipped.
[[nodiscard]] std::expected<void, exr_error> write_exr(const std::filesystem::path& path,
                                                       const Q_readback_aov_result& result);

/// directory / quasi_YYYYMMDD_HHMMSS.exr (local time).
[[nodiscard]] std::filesystem::path make_timestamped_path(const std::filesystem::path& directory = ".");

}  // namespace Q::io
