/// @file exr_writer.cpp
/// @brief Self-contained OpenEXR 2.0 scanline writer (ZIP, 16 lines per chunk).
///
/// Layout written: magic 20000630, version 2 (single-part scanline, short names),
/// header attributes channels / compression / dataWindow / displayWindow /
/// lineOrder / pixelAspectRatio / screenWindowCenter / screenWindowWidth, the
/// chunk offset table, then one chunk per 16 scanlines. Within a chunk each
/// scanline stores its channels in alphabetical order, planar per line. ZIP
/// chunks are byte-deinterleaved, delta-predicted and deflated (zlib); a chunk
/// that does not shrink is stored raw, as the format prescribes.
/// float -> half is round-to-nearest-even, as Imath's half (exr_writer.cpp:44-51).

#include <quasi/io/exr_writer.hpp>

#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <string>
#include <system_error>
#include <vector>

namespace Q::io {

namespace {

enum class pixel_type : int32_t { half_float = 1, full_float = 2 };

/// One output channel: where its samples come from in an interleaved RGBA float buffer.
struct channel {
    std::string name;
    pixel_type type;
    const float* source;  ///< first sample; stride is 4 floats per pixel
};

uint16_t float_to_half(float value) {
    uint32_t x;
    std::memcpy(&x, &value, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return static_cast<uint16_t>(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0u));
    if (x >= 0x477ff000u) return static_cast<uint16_t>(sign | 0x7c00u);  // rounds to infinity
    if (x < 0x33000001u) return static_cast<uint16_t>(sign);             // rounds to zero
    if (x < 0x38800000u) {                                               // subnormal half
        const uint32_t shift = 126u - (x >> 23);
        const uint32_t mant = (x & 0x7fffffu) | 0x800000u;
        uint32_t h = mant >> shift;
        const uint32_t rem = mant & ((1u << shift) - 1u);
        const uint32_t half_way = 1u << (shift - 1u);
        if (rem > half_way || (rem == half_way && (h & 1u))) h++;
        return static_cast<uint16_t>(sign | h);
    }
    uint32_t h = ((x - 0x38000000u) >> 13);
    const uint32_t rem = x & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return static_cast<uint16_t>(sign | h);
}

class byte_sink {
public:
    void bytes(const void* p, size_t n) {
        const auto* b = static_cast<const uint8_t*>(p);
        data_.insert(data_.end(), b, b + n);
    }
    void u8(uint8_t v) { data_.push_back(v); }
    void i32(int32_t v) { bytes(&v, 4); }  // little-endian hosts only (x86-64, aarch64)
    void u64(uint64_t v) { bytes(&v, 8); }
    void f32(float v) { bytes(&v, 4); }
    void str(const std::string& s) { bytes(s.c_str(), s.size() + 1); }
    void attribute(const std::string& name, const std::string& type, const std::vector<uint8_t>& value) {
        str(name);
        str(type);
        i32(static_cast<int32_t>(value.size()));
        bytes(value.data(), value.size());
    }
    [[nodiscard]] std::vector<uint8_t>& data() { return data_; }

private:
    std::vector<uint8_t> data_;
};

std::vector<uint8_t> zip_chunk(const std::vector<uint8_t>& raw) {
    const size_t n = raw.size();
    std::vector<uint8_t> tmp(n);
    // deinterleave: even bytes first, odd bytes second
    uint8_t* t1 = tmp.data();
    uint8_t* t2 = tmp.data() + (n + 1) / 2;
    for (size_t i = 0; i < n; i++) {
        if (i & 1u) {
            *t2++ = raw[i];
        } else {
            *t1++ = raw[i];
        }
    }
    // delta predictor
    int p = tmp.empty() ? 0 : tmp[0];
    for (size_t i = 1; i < n; i++) {
        const int d = static_cast<int>(tmp[i]) - p + (128 + 256);
        p = tmp[i];
        tmp[i] = static_cast<uint8_t>(d);
    }
    uLongf out_size = compressBound(static_cast<uLong>(n));
    std::vector<uint8_t> out(out_size);
    if (compress2(out.data(), &out_size, tmp.data(), static_cast<uLong>(n), Z_DEFAULT_COMPRESSION) != Z_OK) {
        return {};
    }
    out.resize(out_size);
    return out;
}

std::expected<void, exr_error> ensure_parent(const std::filesystem::path& path) {
    const auto parent = path.parent_path();
    if (!parent.empty()) {
        std::error_code ec;
        std::filesystem::create_directories(parent, ec);
        if (ec) return std::unexpected{exr_error::directory_error};
    }
    return {};
}

std::expected<void, exr_error> write_scanline_file(const std::filesystem::path& path, uint32_t width,
                                                   uint32_t height, std::vector<channel> channels) {
    std::sort(channels.begin(), channels.end(),
              [](const channel& a, const channel& b) { return a.name < b.name; });

    byte_sink file;
    file.i32(20000630);
    file.i32(2);

    {   // chlist: name, pixel type, pLinear + 3 reserved bytes, x/y sampling
        byte_sink v;
        for (const auto& c : channels) {
            v.str(c.name);
            v.i32(static_cast<int32_t>(c.type));
            v.u8(0);
            v.u8(0);
            v.u8(0);
            v.u8(0);
            v.i32(1);
            v.i32(1);
        }
        v.u8(0);
        file.attribute("channels", "chlist", v.data());
    }
    file.attribute("compression", "compression", {3});  // ZIP_COMPRESSION
    {
        byte_sink v;
        v.i32(0);
        v.i32(0);
        v.i32(static_cast<int32_t>(width) - 1);
        v.i32(static_cast<int32_t>(height) - 1);
        file.attribute("dataWindow", "box2i", v.data());
        file.attribute("displayWindow", "box2i", v.data());
    }
    file.attribute("lineOrder", "lineOrder", {0});  // INCREASING_Y
    {
        byte_sink v;
        v.f32(1.0f);
        file.attribute("pixelAspectRatio", "float", v.data());
    }
    {
        byte_sink v;
        v.f32(0.0f);
        v.f32(0.0f);
        file.attribute("screenWindowCenter", "v2f", v.data());
    }
    {
        byte_sink v;
        v.f32(1.0f);
        file.attribute("screenWindowWidth", "float", v.data());
    }
    file.u8(0);  // end of header

    constexpr uint32_t lines_per_chunk = 16;
    const uint32_t chunk_count = (height + lines_per_chunk - 1) / lines_per_chunk;
    const size_t table_pos = file.data().size();
    for (uint32_t i = 0; i < chunk_count; i++) file.u64(0);

    size_t bytes_per_line = 0;
    for (const auto& c : channels) bytes_per_line += (c.type == pixel_type::half_float ? 2u : 4u) * width;

    std::vector<uint64_t> offsets(chunk_count);
    std::vector<uint8_t> raw;
    for (uint32_t chunk = 0; chunk < chunk_count; chunk++) {
        const uint32_t y0 = chunk * lines_per_chunk;
        const uint32_t y1 = std::min(height, y0 + lines_per_chunk);
        raw.clear();
        raw.reserve(bytes_per_line * (y1 - y0));
        for (uint32_t y = y0; y < y1; y++) {
            for (const auto& c : channels) {
                const float* src = c.source + static_cast<size_t>(y) * width * 4;
                if (c.type == pixel_type::half_float) {
                    for (uint32_t x = 0; x < width; x++) {
                        const uint16_t h = float_to_half(src[static_cast<size_t>(x) * 4]);
                        raw.push_back(static_cast<uint8_t>(h & 0xffu));
                        raw.push_back(static_cast<uint8_t>(h >> 8));
                    }
                } else {
                    for (uint32_t x = 0; x < width; x++) {
                        uint8_t b[4];
                        std::memcpy(b, &src[static_cast<size_t>(x) * 4], 4);
                        raw.insert(raw.end(), b, b + 4);
                    }
                }
            }
        }
        std::vector<uint8_t> packed = zip_chunk(raw);
        const std::vector<uint8_t>& payload = (!packed.empty() && packed.size() < raw.size()) ? packed : raw;
        offsets[chunk] = file.data().size();
        file.i32(static_cast<int32_t>(y0));
        file.i32(static_cast<int32_t>(payload.size()));
        file.bytes(payload.data(), payload.size());
    }
    std::memcpy(file.data().data() + table_pos, offsets.data(), offsets.size() * sizeof(uint64_t));

    std::ofstream out(path, std::ios::binary | std::ios::trunc);
    if (!out) return std::unexpected{exr_error::write_failed};
    out.write(reinterpret_cast<const char*>(file.data().data()), static_cast<std::streamsize>(file.data().size()));
    out.close();
    if (!out) return std::unexpected{exr_error::write_failed};
    return {};
}

}  // namespace

std::expected<void, exr_error> write_exr(const std::filesystem::path& path, const Q_readback_result& result) {
    if (!result.data || result.width == 0 || result.height == 0) {
        return std::unexpected{exr_error::invalid_data};
    }
    if (auto ok = ensure_parent(path); !ok) return ok;
    std::vector<channel> channels = {
        {"R", pixel_type::half_float, result.data + 0},
        {"G", pixel_type::half_float, result.data + 1},
        {"B", pixel_type::half_float, result.data + 2},
        {"A", pixel_type::half_float, result.data + 3},
    };
    return write_scanline_file(path, result.width, result.height, std::move(channels));
}

std::expected<void, exr_error> write_exr(const std::filesystem::path& path,
                                         const Q_readback_aov_result& result) {
    const Q_aov_buffer& beauty = result.buffers[Q_AOV_BEAUTY];
    if (!beauty.data || beauty.width == 0 || beauty.height == 0) {
        return std::unexpected{exr_error::invalid_data};
    }
    if (auto ok = ensure_parent(path); !ok) return ok;

    std::vector<channel> channels = {
        {"R", pixel_type::half_float, beauty.data + 0},
        {"G", pixel_type::half_float, beauty.data + 1},
        {"B", pixel_type::half_float, beauty.data + 2},
        {"A", pixel_type::half_float, beauty.data + 3},
    };
    if (const float* a = result.buffers[Q_AOV_ALBEDO].data) {
        channels.push_back({"albedo.R", pixel_type::half_float, a + 0});
        channels.push_back({"albedo.G", pixel_type::half_float, a + 1});
        channels.push_back({"albedo.B", pixel_type::half_float, a + 2});
    }
    if (const float* n = result.buffers[Q_AOV_NORMAL].data) {
        channels.push_back({"normal.X", pixel_type::half_float, n + 0});
        channels.push_back({"normal.Y", pixel_type::half_float, n + 1});
        channels.push_back({"normal.Z", pixel_type::half_float, n + 2});
    }
    if (const float* d = result.buffers[Q_AOV_DEPTH].data) {
        channels.push_back({"depth.Z", pixel_type::full_float, d});  // channel 0, full float
    }
    return write_scanline_file(path, beauty.width, beauty.height, std::move(channels));
}

std::filesystem::path make_timestamped_path(const std::filesystem::path& directory) {
    const std::time_t now = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
    std::tm local{};
    localtime_r(&now, &local);
    char name[64];
    std::snprintf(name, sizeof name, "quasi_%04d%02d%02d_%02d%02d%02d.exr", local.tm_year + 1900,
                  local.tm_mon + 1, local.tm_mday, local.tm_hour, local.tm_min, local.tm_sec);
    return directory / name;
}

}  // namespace Q::io
