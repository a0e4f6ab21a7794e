// Host check of quasi_b200/csrc/sincos_table.hpp (the code the CUDA kernels compile:
// device_common.cuh:sincos_round_once) against the oracle's definition of the shader's sin and cos
// (oracle/quasi_oracle.c:qo_sin: the binary64 libm value rounded once to binary32), for every
// phi = 2*pi*r1 the shader can draw: r1 = float(state) * 2^-32 over all 2^32 states takes
// 2^24 + 8 * 2^23 distinct values (plugin.mm:104-107, 113).
//   usage: sincos_table_check [stride]   -> every stride-th distinct r1 (default 1: all of them)
//   prints counts, exit code 1 if more than 2 of the values differ (2^-28 per call expected)
#include "sincos_table.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

float from_bits(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

}  // namespace

int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? static_cast<uint32_t>(std::atoi(argv[1])) : 1u;
    // distinct r1: 0, then every binary32 in [2^-32, 1.0] that is a multiple of 2^-32: all floats from
    // 2^-32 up to 1.0 with exponent >= -32 whose mantissa fits; enumerating the bit patterns of
    // [2^-32, 1.0] covers them (patterns that are not multiples of 2^-32 do not occur below 2^-8,
    // and are skipped by the test on the scaled value).
    const uint32_t first = 0x2f800000u;  // 2^-32
    const uint32_t last = 0x3f800000u;   // 1.0
    const unsigned n_threads = std::max(1u, std::thread::hardware_concurrency());
    std::vector<long> bad_sin(n_threads, 0), bad_cos(n_threads, 0), tested(n_threads, 0);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < n_threads; t++) {
        pool.emplace_back([&, t] {
            for (uint64_t u = static_cast<uint64_t>(first) + static_cast<uint64_t>(t) * stride; u <= last;
                 u += static_cast<uint64_t>(n_threads) * stride) {
                const float r1 = from_bits(static_cast<uint32_t>(u));
                const double scaled = static_cast<double>(r1) * 4294967296.0;
                if (scaled != std::floor(scaled)) continue;  // not float(state) * 2^-32 for any state
                const float phi = 6.28318548202514648f * r1;
                float s, c;
                Q::cuda::sincos::by_table(phi, Q::cuda::sincos::k_host_table, s, c);
                const float want_s = static_cast<float>(std::sin(static_cast<double>(phi)));
                const float want_c = static_cast<float>(std::cos(static_cast<double>(phi)));
                tested[t]++;
                if (std::memcmp(&s, &want_s, 4) != 0) {
                    if (bad_sin[t]++ < 3) std::printf("sin differs at phi %.9g: %.9g vs %.9g\n", phi, s, want_s);
                }
                if (std::memcmp(&c, &want_c, 4) != 0) {
                    if (bad_cos[t]++ < 3) std::printf("cos differs at phi %.9g: %.9g vs %.9g\n", phi, c, want_c);
                }
            }
        });
    }
    for (auto& th : pool) th.join();
    long n = 0, bs = 0, bc = 0;
    for (unsigned t = 0; t < n_threads; t++) n += tested[t], bs += bad_sin[t], bc += bad_cos[t];
    // phi = 0 (r1 = 0)
    float s, c;
    Q::cuda::sincos::by_table(0.0f, Q::cuda::sincos::k_host_table, s, c);
    if (s != 0.0f || std::signbit(s) || c != 1.0f) bs++;
    std::printf("tested %ld sin_mismatches %ld cos_mismatches %ld\n", n + 1, bs, bc);
    return bs + bc <= 2 ? 0 : 1;
}
