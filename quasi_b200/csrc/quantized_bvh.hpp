/// @file quantized_bvh.hpp
/// @brief The binary BVH of the mesh scene with both children of a node in ONE 32-byte sector: boxes as 16-bit
/// positions on a grid over the scene, for the wavefront trace kernel (mesh_wavefront.cu). Host and device.
///
/// Why: once the warp's lanes follow different rays, a visit of the float tree is four 16-byte loads per lane from
/// a line of its own, and ncu shows what that saturates: the L1's data pipe (l1tex__data_pipe_lsu_wavefronts 68 %
/// of peak in the one-kernel path tracer, with the issue slots at 60 %); the waves of scattered rays run three
/// times slower per ray than the wave of camera rays with the same instructions. Halving the bytes of a visit
/// halves the wavefronts, the sectors fetched from L2 and the memory the tree takes there (37 -> 19 MB).
///
/// The tree is the builder's tree: same nodes, same references, pair p holds the children 2p + 1 and 2p + 2 of the
/// node that refers to them (both builders number children that way). Only the boxes change, and only outwards: a
/// coordinate becomes a grid index, rounded down (lower bounds) or up (upper bounds) and moved one more step out.
/// The slab test only decides which subtrees are NOT visited (slab_test.hpp), so a larger box cannot change a hit:
/// the image stays the oracle's bit for bit.
///
/// A pair is 8 words: {Llo.x | Lhi.x << 16, Llo.y | Lhi.y << 16, Llo.z | Lhi.z << 16, Rlo.x | Rhi.x << 16,
/// Rlo.y | Rhi.y << 16, Rlo.z | Rhi.z << 16, reference of L, reference of R} (references as in the float tree:
/// left_first << 3 | count). No integer-to-float conversion on the way to the slab test: a 16-bit index q is put
/// into the mantissa of 2^23 (one PRMT), which is the float 8388608 + q, and the plane's distance along the ray is
/// one FFMA with per-ray constants, t = (8388608 + q) * a + b, a = scale / dir, b = (grid origin - ray origin) / dir
/// - 8388608 * a. The rounding of b (half a grid step at worst, 8388608 * a * 2^-24) is what the extra step of
/// padding pays for; tests/mesh_prune_check.cpp ... quant runs this arithmetic on the host against a brute-force
/// closest hit, with the reciprocal two ulps off either way.

#pragma once

#include "slab_test.hpp"

#include "mesh_scene.hpp"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace Q::cuda::slab {

/// a * b + c rounded ONCE, on the device and on the host (std::fmaf), also in the host check's perturbed build,
/// which otherwise rounds twice: the half grid step above is the error of a fused multiply-add, and the device's
/// FFMA is one.
QUASI_SCAN_FN float fused(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return std::fmaf(a, b, c);
#endif
}

/// Per-ray constants of the quantized slab test. `near`: for each axis the PRMT selector that takes, from a word
/// {lower bound | upper bound << 16}, the bound the ray meets first (the lower one where the direction is positive)
/// and puts it into the mantissa of 2^23; `near ^ k_other_half` takes the other bound. The order of the two plane
/// distances of an axis is known from the sign of `a` (a fused multiply-add is monotonic in the index), so the slab
/// test needs no minimum and maximum per axis: seven instructions less per box than box_hit, same values.
struct quant_ray {
    v3 a;  ///< grid step / direction
    v3 b;  ///< (grid origin - ray origin) / direction - 8388608 * a
    uint32_t near_x, near_y, near_z;
};
inline constexpr uint32_t k_low_half = 0x7610u, k_high_half = 0x7632u, k_other_half = 0x0022u;

/// `inv` = reciprocal(dir), `oi` = origin * inv, as for box_hit.
QUASI_SCAN_FN quant_ray make_quant_ray(const float grid_origin[3], const float grid_step[3], v3 inv, v3 oi) {
    quant_ray r;
    r.a = {grid_step[0] * inv.x, grid_step[1] * inv.y, grid_step[2] * inv.z};
    r.b = {fused(-8388608.0f, r.a.x, fused(grid_origin[0], inv.x, -oi.x)),
           fused(-8388608.0f, r.a.y, fused(grid_origin[1], inv.y, -oi.y)),
           fused(-8388608.0f, r.a.z, fused(grid_origin[2], inv.z, -oi.z))};
    r.near_x = r.a.x < 0.0f ? k_high_half : k_low_half;
    r.near_y = r.a.y < 0.0f ? k_high_half : k_low_half;
    r.near_z = r.a.z < 0.0f ? k_high_half : k_low_half;
    return r;
}

/// The float 8388608 + (the half of w that `selector` names).
QUASI_SCAN_FN float grid_float(uint32_t w, uint32_t selector) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(w, 0x4b000000u, selector));
#else
    const uint32_t bits = 0x4b000000u | (selector == k_low_half ? (w & 0xffffu) : (w >> 16));
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
#endif
}

/// box_hit (slab_test.hpp) for a box given as three words {lo.x | hi.x << 16, lo.y | hi.y << 16, lo.z | hi.z << 16}.
QUASI_SCAN_FN bool box_hit_quantized(uint32_t wx, uint32_t wy, uint32_t wz, const quant_ray& r, float closest, float& t_near) {
    const float nx = fused(grid_float(wx, r.near_x), r.a.x, r.b.x), fx = fused(grid_float(wx, r.near_x ^ k_other_half), r.a.x, r.b.x);
    const float ny = fused(grid_float(wy, r.near_y), r.a.y, r.b.y), fy = fused(grid_float(wy, r.near_y ^ k_other_half), r.a.y, r.b.y);
    const float nz = fused(grid_float(wz, r.near_z), r.a.z, r.b.z), fz = fused(grid_float(wz, r.near_z ^ k_other_half), r.a.z, r.b.z);
    const float tn = std::fmax(std::fmax(nx, ny), std::fmax(nz, 0.0f));
    const float tf = std::fmin(std::fmin(fx, fy), std::fmin(fz, closest));
    t_near = tn;
    return tn <= tf;
}

}  // namespace Q::cuda::slab

namespace Q::cuda::mesh {

inline constexpr uint32_t k_pair_words = 8;

struct quantized_tree {
    std::vector<uint32_t> pairs;  ///< k_pair_words per pair; empty: the tree cannot be put in this form
    float origin[3] = {0, 0, 0};
    float step[3] = {0, 0, 0};
    uint32_t pair_count = 0;
};

/// @brief The children of every interior node of s.nodes on the grid. Returns an empty tree if the nodes are not
/// numbered in pairs (2p + 1, 2p + 2) or the scene has no extent.
inline quantized_tree quantize_pairs(const scene& s) {
    quantized_tree q;
    const size_t n = s.nodes.size();
    if (n < 3 || n % 2 == 0) return q;
    {  // every node that can be reached (the GPU builder leaves slots of collapsed subtrees unused)
        std::vector<uint32_t> todo{0u};
        size_t seen = 0;
        while (!todo.empty()) {
            const node& nd = s.nodes[todo.back()];
            todo.pop_back();
            if (++seen > n) return q;  // not a tree
            if (nd.count != 0) continue;
            if (nd.left_first % 2 == 0 || static_cast<size_t>(nd.left_first) + 1 >= n) return q;
            todo.push_back(nd.left_first);
            todo.push_back(nd.left_first + 1);
        }
    }
    // the grid: 65 520 steps across the root's box, eight steps of margin on either side
    const node& root = s.nodes[0];
    for (int a = 0; a < 3; a++) {
        const double extent = static_cast<double>(root.bmax[a]) - static_cast<double>(root.bmin[a]);
        if (!(extent > 0.0) || !std::isfinite(extent)) return q;
        q.step[a] = static_cast<float>(extent / 65520.0);
        q.origin[a] = static_cast<float>(static_cast<double>(root.bmin[a]) - 8.0 * static_cast<double>(q.step[a]));
    }
    q.pair_count = static_cast<uint32_t>((n - 1) / 2);
    q.pairs.assign(static_cast<size_t>(q.pair_count) * k_pair_words, 0u);
    auto index = [&q](float x, int a, bool upper) {
        // a box outside the root's (an unused slot of the GPU builder) ends up clamped: it is never referred to
        const double at = (static_cast<double>(x) - static_cast<double>(q.origin[a])) / static_cast<double>(q.step[a]);
        double i = upper ? std::ceil(at) + 1.0 : std::floor(at) - 1.0;
        if (!(i >= 0.0)) i = 0.0;
        if (i > 65535.0) i = 65535.0;
        return static_cast<uint32_t>(i);
    };
    for (uint32_t p = 0; p < q.pair_count; p++) {
        uint32_t* w = &q.pairs[static_cast<size_t>(p) * k_pair_words];
        for (int c = 0; c < 2; c++) {
            const node& nd = s.nodes[2 * static_cast<size_t>(p) + 1 + c];
            const uint32_t lo[3] = {index(nd.bmin[0], 0, false), index(nd.bmin[1], 1, false), index(nd.bmin[2], 2, false)};
            const uint32_t hi[3] = {index(nd.bmax[0], 0, true), index(nd.bmax[1], 1, true), index(nd.bmax[2], 2, true)};
            for (int a = 0; a < 3; a++) w[3 * c + a] = lo[a] | (hi[a] << 16);
            w[6 + c] = (nd.left_first << 3) | nd.count;
        }
    }
    return q;
}

}  // namespace Q::cuda::mesh
