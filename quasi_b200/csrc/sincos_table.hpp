/// @file sincos_table.hpp
/// @brief sin and cos of phi in [0, 2*pi] as correctly rounded binary32, by table (host and device).
///
/// The shader draws phi = 2*pi*r1 and takes cos(phi), sin(phi) (/root/reference/backends/metal/
/// plugin.mm:113-124); the oracle defines them as the binary64 libm values rounded once to
/// binary32 (oracle/quasi_oracle.c:qo_sin). device_common.cuh:sincos_round_once evaluates two
/// degree-13/14 polynomials after a reduction by pi/2: 71 instructions. This form reduces by
/// pi/512 instead, reads {sin a, cos a} of the nearest multiple a = j*pi/512 from a table of
/// correctly rounded binary64 values (tools/make_sincos_table.py; exact 0 and +-1 at the
/// quarter turns) and corrects by the remainder b, |b| <= pi/1024:
///     sin(a + b) = sin a + (sin a * (cos b - 1) + cos a * sin b)
///     cos(a + b) = cos a + (cos a * (cos b - 1) - sin a * sin b)
/// with sin b = b + b^3 (-1/6 + b^2/120) and cos b - 1 = b^2 (-1/2 + b^2/24); the terms left
/// out are below 2^-66 relative to the result. Every operation is one IEEE binary64 operation
/// (explicit fma), so host and device produce the same bits; the binary64 result is within a
/// few 2^-53 of the true value, i.e. its rounding to binary32 differs from the correctly rounded
/// one with probability ~2^-28 per call, as for the polynomial form.
/// tests/sincos_table_check.cpp compares it with the oracle's definition for every phi the RNG
/// can produce.

#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define QUASI_SINCOS_FN __host__ __device__ __forceinline__
#define QUASI_SINCOS_TABLE_SPACE static __device__ const
#else
#define QUASI_SINCOS_FN inline
#define QUASI_SINCOS_TABLE_SPACE static const
#endif

namespace Q::cuda::sincos {

#include "sincos_table_constants.inc"

struct alignas(16) pair64 {
    double sin_a, cos_a;
};

inline constexpr int k_steps_per_turn = 1024;

/// Host copy of the table (the device copy lives in device_common.cuh).
static const pair64 k_host_table[k_steps_per_turn + 1] = {
#include "sincos_table_data.inc"
};

/// @param table k_steps_per_turn + 1 entries, entry j = {sin, cos}(j * pi / 512).
QUASI_SINCOS_FN void by_table(float phi, const pair64* table, float& sin_phi, float& cos_phi) {
    const double x = static_cast<double>(phi);
    const double jd = rint(x * QUASI_SINCOS_INV_STEP);
    const int j = static_cast<int>(jd);
    double b = fma(-jd, QUASI_SINCOS_STEP_HI, x);  // exact product: STEP_HI has 41 bits, j at most 11
    b = fma(-jd, QUASI_SINCOS_STEP_LO, b);
    const double b2 = b * b;
    const double sin_b = fma(b2 * b, fma(b2, 0x1.1111111111111p-7, -0x1.5555555555555p-3), b);  // 1/120, -1/6
    const double cos_b_m1 = b2 * fma(b2, 0x1.5555555555555p-5, -0.5);                            // 1/24
#if defined(__CUDA_ARCH__)
    const double2 e = __ldg(reinterpret_cast<const double2*>(table) + j);
    const double sa = e.x, ca = e.y;
#else
    const double sa = table[j].sin_a, ca = table[j].cos_a;
#endif
    const double sv = sa + fma(ca, sin_b, sa * cos_b_m1);
    const double cv = ca + fma(-sa, sin_b, ca * cos_b_m1);
    sin_phi = static_cast<float>(sv);
    cos_phi = static_cast<float>(cv);
}

}  // namespace Q::cuda::sincos
