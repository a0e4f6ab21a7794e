/// @file scan_decide.hpp
/// @brief The DECISION arithmetic of the quad scene's closest-hit scan, for device and host.
///
/// pathtrace.cu finds the closest hit of trace_scene + intersect_quad
/// (/root/reference/backends/metal/plugin.mm:157-212) without running the literal loop
/// for every ray: per quad (or per box of five quads) it computes a cheap form of the
/// same quantities and decides with explicit margins; a ray inside any margin goes to
/// the literal loop, and the accepted hit is recomputed with the literal operations, so
/// nothing computed here ever reaches a pixel. This header holds exactly that deciding
/// code. It compiles for the device (nvcc) and, with the two intrinsics it uses mapped to
/// their host equivalents, for the host, where tests/test_scan_decide_cpu.py runs it
/// against the literal loop on millions of rays aimed at the edges, the seams between a
/// box and the floor and the t_min band.

#pragma once

#include "scene_params.hpp"

#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define QUASI_SCAN_FN __device__ __forceinline__
#else
#define QUASI_SCAN_FN inline
#endif

namespace Q::cuda::scan {

using Q::cuda::v3;

#if defined(__CUDA_ARCH__)
/// One MUFU.RCP, relative error <= 2^-23.
QUASI_SCAN_FN float rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
QUASI_SCAN_FN float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#elif defined(QUASI_SCAN_CHECK_PERTURB)
/// Host check only: the exact quotient moved two units in the last place up or down
/// (which way depends on the argument's last bit), twice the error MUFU.RCP may have.
QUASI_SCAN_FN float rcp(float x) {
    const float r = 1.0f / x;
    std::uint32_t bits;
    std::memcpy(&bits, &x, sizeof bits);
    const float towards = (bits & 1u) != 0u ? INFINITY : -INFINITY;
    return std::nextafter(std::nextafter(r, towards), towards);
}
QUASI_SCAN_FN float fma_rn(float a, float b, float c) { return a * b + c; }  // and no contraction either
#else
QUASI_SCAN_FN float rcp(float x) { return 1.0f / x; }
QUASI_SCAN_FN float fma_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
#endif

// Margins of the fast scan. t' = num * rcp(denom) is within 2^-21 relative of the
// IEEE quotient t; the bands below are ten times wider than that.
inline constexpr float k_inf = __builtin_huge_valf();
inline constexpr float k_t_min_lo = 0.001f * (1.0f - 4e-6f);  // below: surely t < t_min
inline constexpr float k_t_min_hi = 0.001f * (1.0f + 4e-6f);  // above: surely t >= t_min
inline constexpr float k_tie = 1.0f + 4e-6f;                  // two hits closer than this: literal scan

struct scan_state {
    float best_t = k_inf;        ///< smallest t' among quads that are surely hit
    float challenger_t = k_inf;  ///< smallest t' among every other quad that may be hit
    int best_i = -1;
};

/// Files one plane hit. `m` = max(|alpha' - 1/2|, |beta' - 1/2|): the point is
/// surely outside the quad beyond half_plus and surely inside within half_minus.
/// `ahead`: the plane is not surely behind t_min. The scan evaluates every quad for every
/// lane without a branch: the inside test is 20 instructions, a divergent branch around it
/// costs 3 and nearly always has a lane that takes it. What a lane computes past a failed
/// `ahead` is discarded here.
///
/// Two of the shader's rejections are NOT taken per quad but once per ray, on the winner
/// (accept_winner below): |denom| < 1e-8 (plugin.mm:174) and t < t_min inside the band
/// (k_t_min_lo, k_t_min_hi] (plugin.mm:177). Per quad they cost two comparisons each, sixteen
/// times per ray, and they all but never fire. That is sound because a quad the shader rejects
/// for one of them can only be filed here as a candidate it is not, never the other way round:
/// if such a quad is the closest "sure" one, accept_winner sees its t' or its denominator and
/// hands the ray to the literal loop; if it is not, it is a challenger like any other (literal
/// loop when it is in front of, or ties with, the winner; irrelevant when it is behind).
/// With denom == 0 the reciprocal is infinite and t' is +-inf or NaN: the point, alpha', beta'
/// and m are then infinite or NaN, never finite, so neither `maybe` nor `sure` holds
/// (comparisons with NaN are false; fmax(|a|, |b|) is NaN only if both are).
QUASI_SCAN_FN void consider(scan_state& s, float half_plus, float half_minus, int i, float t_fast, float m, bool ahead) {
    const bool maybe = ahead && m <= half_plus;
    const bool sure = ahead && m < half_minus;
    const bool better = sure && t_fast < s.best_t;
    const float loser = better ? s.best_t : (maybe ? t_fast : k_inf);
    s.challenger_t = std::fmin(s.challenger_t, loser);
    s.best_t = better ? t_fast : s.best_t;
    s.best_i = better ? i : s.best_i;
}

QUASI_SCAN_FN float axis(v3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

/// Quad whose normal is along axis K (two normal components are exactly zero, so
/// the shader's three-term dot products collapse to one product, bit for bit)
/// and whose barycentrics do not depend on coordinate K.
/// The constants come from the quad's packed record (scene_params.hpp:scan_record), in the order
/// the scan uses them: sm_100 has no constant-bank operands for FP32 instructions, every scene
/// constant costs an LDCU issue slot, and three 16-byte loads per quad replace eight small ones.
template <int K>
QUASI_SCAN_FN void scan_flat(scan_state& s, const scan_record& q, int i, v3 ro, v3 rd) {
    constexpr int I = (K + 1) % 3, J = (K + 2) % 3;
    const float denom = q.r[0] * axis(rd, K);
    const float num = q.r[1] - q.r[0] * axis(ro, K);
    const float t_fast = num * rcp(denom);
    const bool ahead = t_fast > k_t_min_lo;
    const float pi = fma_rn(axis(rd, I), t_fast, axis(ro, I)) - q.r[2];
    const float pj = fma_rn(axis(rd, J), t_fast, axis(ro, J)) - q.r[3];
    const float alpha = fma_rn(pi, q.r[4], pj * q.r[5]);
    const float beta = fma_rn(pi, q.r[6], pj * q.r[7]);
    consider(s, q.r[8], q.r[9], i, t_fast, std::fmax(std::fabs(alpha), std::fabs(beta)), ahead);
}

/// Quad with a horizontal normal and one vertical edge (the side of a box
/// standing on the floor): alpha lives in (x, z), beta along y.
QUASI_SCAN_FN void scan_upright(scan_state& s, const scan_record& q, int i, v3 ro, v3 rd) {
    const float denom = q.r[0] * rd.x + q.r[1] * rd.z;
    const float num = q.r[2] - (q.r[0] * ro.x + q.r[1] * ro.z);
    const float t_fast = num * rcp(denom);
    const bool ahead = t_fast > k_t_min_lo;
    const float px = fma_rn(rd.x, t_fast, ro.x) - q.r[3];
    const float py = fma_rn(rd.y, t_fast, ro.y) - q.r[4];
    const float pz = fma_rn(rd.z, t_fast, ro.z) - q.r[5];
    const float alpha = fma_rn(px, q.r[6], pz * q.r[7]);
    const float beta = py * q.r[8];
    consider(s, q.r[9], q.r[10], i, t_fast, std::fmax(std::fabs(alpha), std::fabs(beta)), ahead);
}

QUASI_SCAN_FN void scan_general(scan_state& s, const quad_params& q, int i, v3 ro, v3 rd) {
    // dot(normal, rd) and plane_d - dot(normal, ro) in the shader's association: the shader's values
    const float denom = q.normal[0] * rd.x + q.normal[1] * rd.y + q.normal[2] * rd.z;
    const float num = q.plane_d - (q.normal[0] * ro.x + q.normal[1] * ro.y + q.normal[2] * ro.z);
    const float t_fast = num * rcp(denom);
    const bool ahead = t_fast > k_t_min_lo;
    {
        const v3 planar{fma_rn(rd.x, t_fast, ro.x) - q.center[0], fma_rn(rd.y, t_fast, ro.y) - q.center[1],
                        fma_rn(rd.z, t_fast, ro.z) - q.center[2]};
        const float alpha = fma_rn(planar.x, q.fast_a[0], fma_rn(planar.y, q.fast_a[1], planar.z * q.fast_a[2]));
        const float beta = fma_rn(planar.x, q.fast_b[0], fma_rn(planar.y, q.fast_b[1], planar.z * q.fast_b[2]));
        consider(s, q.half_plus, q.half_minus, i, t_fast, std::fmax(std::fabs(alpha), std::fabs(beta)), ahead);
    }
}

/// The five faces of a box in one slab test (scene_params.hpp:box_params). Per slab the
/// ray can only enter through the side it moves towards; that face's plane distance is
/// computed with the face's own constants in the form of scan_upright / scan_flat<1>, so
/// it is the t' the per-quad scan would have filed. The entry point is where the latest of
/// the three is reached; it is a hit if it lies within the other two slabs.
QUASI_SCAN_FN void scan_box(scan_state& s, const box_params& b, v3 ro, v3 rd) {
    float t_entry = -k_inf;
    int face = -1, entry_slab = 0;
    bool parallel = false;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < 2; k++) {
        const float along = b.axis[k][0] * rd.x + b.axis[k][1] * rd.z;
        const bool high = along < 0.0f;  // moving against the slab direction: enters through the high side
        const float nx = high ? b.side_normal[k][1][0] : b.side_normal[k][0][0];
        const float nz = high ? b.side_normal[k][1][1] : b.side_normal[k][0][1];
        const float d = high ? b.side_d[k][1] : b.side_d[k][0];
        const float denom = nx * rd.x + nz * rd.z;          // scan_upright's value for that face
        const float num = d - (nx * ro.x + nz * ro.z);      // scan_upright's value for that face
        const float t = num * rcp(denom);
        parallel = parallel || std::fabs(along) < 1e-6f;
        const bool later = t > t_entry;
        t_entry = later ? t : t_entry;
        face = later ? static_cast<int>(high ? b.side_index[k][1] : b.side_index[k][0]) : face;
        entry_slab = later ? k : entry_slab;
    }
    {
        // vertical slab: the top face when moving down, the open bottom (no face) when moving up
        const bool down = rd.y < 0.0f;
        const float t_top = (b.top_d - b.top_normal_y * ro.y) * rcp(b.top_normal_y * rd.y);  // scan_flat<1>
        const float t_bottom = (b.y_bottom - ro.y) * rcp(rd.y);
        const float t = down ? t_top : t_bottom;
        parallel = parallel || std::fabs(rd.y) < 1e-6f;
        const bool later = t > t_entry;
        t_entry = later ? t : t_entry;
        face = later ? (down ? static_cast<int>(b.top_index) : -1) : face;
        entry_slab = later ? 2 : entry_slab;
    }
    // Box coordinates along the ray are affine in t: c_k(t) = c_k(ro) + t * (axis_k . rd).
    const float s0 = fma_rn(b.axis[0][0], ro.x, b.axis[0][1] * ro.z), s1 = fma_rn(b.axis[1][0], ro.x, b.axis[1][1] * ro.z);
    const float a0 = fma_rn(b.axis[0][0], rd.x, b.axis[0][1] * rd.z), a1 = fma_rn(b.axis[1][0], rd.x, b.axis[1][1] * rd.z);
    // where the entry point sits in the other two slabs (its own slab is met by construction)
    const float e0 = std::fabs(fma_rn(a0, t_entry, s0) - b.mid[0]) - b.half[0];
    const float e1 = std::fabs(fma_rn(a1, t_entry, s1) - b.mid[1]) - b.half[1];
    const float e2 = std::fabs(fma_rn(rd.y, t_entry, ro.y) - b.mid[2]) - b.half[2];
    const float outside = std::fmax(entry_slab == 0 ? -k_inf : e0, std::fmax(entry_slab == 1 ? -k_inf : e1,
                                                                       entry_slab == 2 ? -k_inf : e2));
    // A ray that starts within t_min of a face it is not leaving (the floor next to a block)
    // passes THROUGH that face (plugin.mm:177 rejects t < t_min) and meets the far side from
    // inside. The slab logic has no such case: if the point at t_min is inside the box the
    // ray goes to the literal scan.
    const float m0 = std::fabs(fma_rn(a0, 0.001f, s0) - b.mid[0]) - b.half[0];
    const float m1 = std::fabs(fma_rn(a1, 0.001f, s1) - b.mid[1]) - b.half[1];
    const float m2 = std::fabs(fma_rn(rd.y, 0.001f, ro.y) - b.mid[2]) - b.half[2];
    const bool tunnels = std::fmax(m0, std::fmax(m1, m2)) <= b.delta;
    // "Surely inside slab k" is e_k < -delta, measured at the entry point. Seen from the face
    // of slab k the same edge is approached at the speed the ray crosses the ENTRY slab: when
    // that is the slower one (a ray almost in the plane of the entry face) the ray is still
    // within delta of the box while it crosses the plane of slab k, well before the entry, and
    // the literal test may count THAT face as hit. So the margin of slab k is delta times the
    // ratio of the two speeds where that exceeds one: e_k * v_entry + delta * max(v_entry, v_k) < 0.
    const float v0 = std::fabs(a0), v1 = std::fabs(a1), v2 = std::fabs(rd.y);
    const float v_entry = entry_slab == 0 ? v0 : (entry_slab == 1 ? v1 : v2);
    const float q0 = fma_rn(e0, v_entry, b.delta * std::fmax(v_entry, v0));
    const float q1 = fma_rn(e1, v_entry, b.delta * std::fmax(v_entry, v1));
    const float q2 = fma_rn(e2, v_entry, b.delta * std::fmax(v_entry, v2));
    const float edge = std::fmax(entry_slab == 0 ? -k_inf : q0, std::fmax(entry_slab == 1 ? -k_inf : q1,
                                                                    entry_slab == 2 ? -k_inf : q2));
    const bool ahead = t_entry > k_t_min_lo;
    const bool maybe = ahead && outside <= b.delta;
    // a ray parallel to a slab, or entering where there is no face, is never decided here
    const bool sure = ahead && edge < 0.0f && t_entry > k_t_min_hi && !parallel && face >= 0;
    const bool better = sure && t_entry < s.best_t;
    // a box that may be hit but is not surely hit at the entry may be hit earlier, on another face
    float loser = better ? s.best_t : (maybe ? 0.0f : k_inf);
    loser = (parallel || tunnels) ? 0.0f : loser;  // forces the literal scan
    s.challenger_t = std::fmin(s.challenger_t, loser);
    s.best_t = better ? t_entry : s.best_t;
    s.best_i = better ? face : s.best_i;
}

/// What scan_flat and scan_upright read a quad from: its packed record.
#define QUASI_SCAN_SOURCE(scene, i) (scene).records[i]

/// The whole fast scan for any scene (one loop per shape with launch-uniform bounds, then
/// the boxes). The kernel unrolls the same calls for the demo scene's layout.
QUASI_SCAN_FN void scan_scene(scan_state& s, const scene_params& scene, v3 ro, v3 rd) {
    for (uint32_t i = scene.shape_begin[k_shape_flat_y]; i < scene.shape_begin[k_shape_flat_y + 1]; i++) {
        scan_flat<1>(s, QUASI_SCAN_SOURCE(scene, i), static_cast<int>(i), ro, rd);
    }
    for (uint32_t i = scene.shape_begin[k_shape_upright]; i < scene.shape_begin[k_shape_upright + 1]; i++) {
        scan_upright(s, QUASI_SCAN_SOURCE(scene, i), static_cast<int>(i), ro, rd);
    }
    for (uint32_t i = scene.shape_begin[k_shape_flat_x]; i < scene.shape_begin[k_shape_flat_x + 1]; i++) {
        scan_flat<0>(s, QUASI_SCAN_SOURCE(scene, i), static_cast<int>(i), ro, rd);
    }
    for (uint32_t i = scene.shape_begin[k_shape_flat_z]; i < scene.shape_begin[k_shape_flat_z + 1]; i++) {
        scan_flat<2>(s, QUASI_SCAN_SOURCE(scene, i), static_cast<int>(i), ro, rd);
    }
    for (uint32_t i = scene.shape_begin[k_shape_general]; i < scene.shape_begin[k_shape_general + 1]; i++) {
        scan_general(s, scene.quads[i], static_cast<int>(i), ro, rd);
    }
    for (uint32_t n = 0; n < scene.box_count; n++) scan_box(s, scene.boxes[n], ro, rd);
}

/// True when the scan could not decide and the literal loop has to.
QUASI_SCAN_FN bool needs_literal_scan(const scan_state& s) {
    // A challenger matters only if it could be in front of, or tie with, the sure
    // winner. With no sure winner (best_t == inf) any challenger matters.
    // A winner inside the t_min band may be one the shader rejects (plugin.mm:177).
    return (s.challenger_t < k_inf && s.challenger_t <= s.best_t * k_tie) || s.best_t <= k_t_min_hi;
}

/// Second half of the per-ray acceptance, for a scan that named a winner: `denom` is the
/// shader's dot(normal, rd) of that quad (the caller computes it anyway, for the literal t).
/// plugin.mm:174 rejects the quad below 1e-8; the literal loop then decides the ray.
QUASI_SCAN_FN bool accept_winner(float denom) { return std::fabs(denom) >= 1e-8f; }

}  // namespace Q::cuda::scan
