/// @file slab_test.hpp
/// @brief The pruning arithmetic of the mesh traversal (mesh_trace.cu), for device and host.
///
/// The slab test never decides a hit, it only decides which subtrees are not visited, so it
/// may be approximate as long as it is conservative: a box that holds a triangle the ray
/// hits (by the literal Moeller-Trumbore test) must never be skipped. The builder pads every
/// box by 1e-4 of its extent + 1e-5 (mesh_scene.hpp); what is evaluated here, one MUFU.RCP
/// per axis (relative error 2^-23, clamped so that origin * 1/dir stays finite) and one FFMA
/// per plane against the precomputed origin * 1/dir, is off by less than 2e-6 in distance.
/// tests/mesh_prune_check.cpp runs exactly this code on the host, with the reciprocal pushed
/// two ulps either way, against a brute-force closest hit over all triangles.

#pragma once

#include "scan_decide.hpp"  // v3, rcp(), fma_rn(), QUASI_SCAN_FN

namespace Q::cuda::slab {

using scan::fma_rn;
using scan::rcp;

/// 1/x for the slab test: one MUFU.RCP, clamped so that origin * result stays finite.
QUASI_SCAN_FN float reciprocal(float x) { return std::fmin(std::fmax(rcp(x), -0x1p+100f), 0x1p+100f); }

/// Slab test against a padded box {lo.xyz, hi.xyz}: `inv` = reciprocal(dir), `oi` = origin * inv.
/// Returns the entry distance in t_near. NaNs cannot occur because inv is finite.
template <class F4>
QUASI_SCAN_FN bool box_hit(F4 lo, F4 hi, v3 inv, v3 oi, float closest, float& t_near) {
    const float t0x = fma_rn(lo.x, inv.x, -oi.x), t1x = fma_rn(hi.x, inv.x, -oi.x);
    const float t0y = fma_rn(lo.y, inv.y, -oi.y), t1y = fma_rn(hi.y, inv.y, -oi.y);
    const float t0z = fma_rn(lo.z, inv.z, -oi.z), t1z = fma_rn(hi.z, inv.z, -oi.z);
    const float tn = std::fmax(std::fmax(std::fmin(t0x, t1x), std::fmin(t0y, t1y)), std::fmax(std::fmin(t0z, t1z), 0.0f));
    const float tf = std::fmin(std::fmin(std::fmax(t0x, t1x), std::fmax(t0y, t1y)), std::fmin(std::fmax(t0z, t1z), closest));
    t_near = tn;
    return tn <= tf;
}

}  // namespace Q::cuda::slab
