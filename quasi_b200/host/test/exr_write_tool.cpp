/// @file exr_write_tool.cpp
/// @brief Test helper: raw float32 layers in, EXR out, through Q::io::write_exr.
///
/// Usage: exr_write_tool <in.bin> <out.exr> [beauty-only]
/// in.bin = uint32 width, uint32 height, uint32 layer_count (1 or 4), then
/// layer_count RGBA float32 images. With one layer the Q_readback_result
/// overload is used, otherwise the Q_readback_aov_result overload.

#include <quasi/io/exr_writer.hpp>

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    uint32_t hdr[3];
    if (!in.read(reinterpret_cast<char*>(hdr), sizeof hdr)) return 2;
    const size_t layer = static_cast<size_t>(hdr[0]) * hdr[1] * 4;
    std::vector<float> data(layer * hdr[2]);
    if (!in.read(reinterpret_cast<char*>(data.data()), static_cast<std::streamsize>(data.size() * 4))) return 2;
    std::expected<void, Q::io::exr_error> r;
    if (hdr[2] == 1) {
        Q_readback_result rb{data.data(), hdr[0], hdr[1], 4};
        r = Q::io::write_exr(argv[2], rb);
    } else {
        Q_readback_aov_result rb{};
        for (uint32_t i = 0; i < hdr[2] && i < Q_AOV_COUNT; i++) rb.buffers[i] = {data.data() + layer * i, hdr[0], hdr[1], 4};
        r = Q::io::write_exr(argv[2], rb);
    }
    if (!r) {
        std::fprintf(stderr, "%s\n", Q::io::to_string(r.error()));
        return 1;
    }
    return 0;
}
