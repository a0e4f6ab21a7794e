/// @file wide_bvh.hpp
/// @brief A 4-wide BVH collapsed from the binary one (host side), and the traversal step on
/// one of its nodes (host and device).
///
/// Nothing in the reference corresponds to this file (its scene intersection is a linear
/// scan, /root/reference/backends/metal/plugin.mm:199-212). The binary tree of
/// mesh_scene.hpp stays the contract with the oracle (node order, leaf order, padded boxes:
/// oracle/quasi_oracle_mesh.h); the wide tree is a second VIEW of it for the kernel: every
/// wide node adopts up to four nodes of the binary tree as its children, with their boxes
/// as they are, so it prunes exactly what the binary tree prunes and the closest hit, which
/// never depended on the order of the visits, is the same (mesh_trace.cu, header).
///
/// Why: a ray visits half as many wide nodes as binary ones, so the chain of dependent
/// loads per ray is half as long and the per-visit overhead of the kernel's scheduler loop
/// (votes, pops) is paid half as often; the kernel is bound by issue slots at low SIMD
/// efficiency and by L1/L2 latency, not by bytes.
///
/// Node layout (128 bytes, one cache line, 7 x LDG.128): the boxes of the four children
/// transposed, {lo.x[4]} {lo.y[4]} {lo.z[4]} {hi.x[4]} {hi.y[4]} {hi.z[4]}, then the four
/// child references in the encoding of mesh_trace.cu (index << 3 | count; count 0 = wide
/// node, 1..4 = leaf starting at that slot of the leaf order), then 16 unused bytes.
/// A node with fewer than four children holds boxes no ray can hit in the free places.

#pragma once

#include "mesh_scene.hpp"
#include "slab_test.hpp"

#include <vector>

namespace Q::cuda::mesh {

inline constexpr uint32_t k_wide = 4;
inline constexpr uint32_t k_wide_node_float4 = 8;
/// Coordinate of a box no ray can hit: with |1/dir| >= 1 its plane distances are at least
/// 3e38 in magnitude on every axis, beyond any closest-hit distance (at most 1e30).
inline constexpr float k_nowhere = 3e38f;

struct wide_tree {
    std::vector<float4_t> nodes;  ///< k_wide_node_float4 per wide node, node 0 is the root
    uint32_t node_count = 0;
    uint32_t depth = 0;       ///< longest root-to-leaf path in wide nodes
    uint32_t stack_need = 0;  ///< most entries a traversal can have deferred at any time
    uint32_t children = 0;    ///< occupied child places over all nodes
};

/// @brief Collapses the binary tree of `s` into 4-wide nodes.
///
/// A wide node starts from the two children of a binary interior node and, while it has a
/// free place, replaces the interior child with the largest box by that child's own two
/// children (in place: the order of the children stays the depth-first order of the binary
/// tree). Wide nodes are numbered depth first, children before siblings.
/// A binary root that is a leaf becomes a wide root with that one child.
///
/// A visit defers up to three children where the binary tree defers one, on fewer levels; on
/// the longest paths of a deep tree that can add up to more than the binary tree needs (39
/// entries against 24 for the million-triangle scene). `stack_limit` bounds it: a child is
/// opened only if, counting one deferred entry per remaining level of the binary tree below,
/// the stack still fits. Any binary tree whose own traversal fits (depth <= stack_limit + 1)
/// can be collapsed under this rule; nodes on its longest paths stay narrower.
inline wide_tree collapse_wide(const scene& s, uint32_t stack_limit = 0xffffffffu) {
    wide_tree out;
    if (s.nodes.empty()) return out;
    // levels of interior nodes below and including each binary node (children have higher numbers)
    std::vector<uint8_t> levels(s.nodes.size(), 0);
    for (size_t i = s.nodes.size(); i-- > 0;) {
        const node& n = s.nodes[i];
        if (n.count == 0) {
            const uint8_t below = levels[n.left_first] > levels[n.left_first + 1] ? levels[n.left_first] : levels[n.left_first + 1];
            levels[i] = static_cast<uint8_t>(below + 1 > 255 ? 255 : below + 1);
        }
    }
    auto area = [](const node& n) {
        const float dx = n.bmax[0] - n.bmin[0], dy = n.bmax[1] - n.bmin[1], dz = n.bmax[2] - n.bmin[2];
        return (dx * dy + dy * dz) + dz * dx;
    };
    constexpr uint32_t k_no_parent = 0xffffffffu;
    struct pending {
        uint32_t binary;    ///< the binary node whose subtree this wide node covers
        uint32_t depth;     ///< in wide nodes, the root is 1
        uint32_t deferred;  ///< entries already on the stack when a traversal descends into this node
        uint32_t parent, place;
    };
    std::vector<pending> todo{{0u, 1u, 0u, k_no_parent, 0u}};
    while (!todo.empty()) {
        const pending p = todo.back();
        todo.pop_back();
        const uint32_t me = out.node_count++;
        out.nodes.resize(out.nodes.size() + k_wide_node_float4, float4_t{0.0f, 0.0f, 0.0f, 0.0f});
        if (p.parent != k_no_parent) {
            (&out.nodes[static_cast<size_t>(p.parent) * k_wide_node_float4 + 6].x)[p.place] = bits_to_float(me << 3);
        }

        uint32_t child[k_wide];
        uint32_t n_children = 0;
        const node& b = s.nodes[p.binary];
        if (b.count != 0) {
            child[n_children++] = p.binary;  // a tree that is one leaf
        } else {
            child[n_children++] = b.left_first;
            child[n_children++] = b.left_first + 1u;
            while (n_children < k_wide) {
                int widest = -1;
                float widest_area = -1.0f;
                for (uint32_t c = 0; c < n_children; c++) {
                    const node& x = s.nodes[child[c]];
                    if (x.count != 0 || !(area(x) > widest_area)) continue;
                    // entries deferred here once c is opened, plus one per level of the deepest child left
                    uint32_t deepest = static_cast<uint32_t>(levels[child[c]]) - 1u;
                    for (uint32_t o = 0; o < n_children; o++) {
                        if (o != c && levels[child[o]] > deepest) deepest = levels[child[o]];
                    }
                    if (static_cast<uint64_t>(p.deferred) + n_children + deepest > stack_limit) continue;
                    widest_area = area(x);
                    widest = static_cast<int>(c);
                }
                if (widest < 0) break;
                const uint32_t opened = child[widest];
                for (uint32_t c = n_children; c > static_cast<uint32_t>(widest) + 1u; c--) child[c] = child[c - 1];
                child[widest] = s.nodes[opened].left_first;
                child[widest + 1] = s.nodes[opened].left_first + 1u;
                n_children++;
            }
        }

        float4_t* f = &out.nodes[static_cast<size_t>(me) * k_wide_node_float4];
        float* lo[3] = {&f[0].x, &f[1].x, &f[2].x};
        float* hi[3] = {&f[3].x, &f[4].x, &f[5].x};
        float* ref = &f[6].x;
        for (uint32_t c = 0; c < k_wide; c++) {
            if (c >= n_children) {
                for (int a = 0; a < 3; a++) lo[a][c] = hi[a][c] = k_nowhere;
                ref[c] = bits_to_float(0xffffffffu);
                continue;
            }
            const node& x = s.nodes[child[c]];
            for (int a = 0; a < 3; a++) {
                lo[a][c] = x.bmin[a];
                hi[a][c] = x.bmax[a];
            }
            if (x.count != 0) ref[c] = bits_to_float((x.left_first << 3) | x.count);  // interior: set when numbered
        }

        const uint32_t deferred_below = p.deferred + n_children - 1u;
        if (p.depth > out.depth) out.depth = p.depth;
        if (deferred_below > out.stack_need) out.stack_need = deferred_below;
        out.children += n_children;
        // last child first, so that the first child's subtree gets the next numbers
        for (uint32_t c = n_children; c-- > 0;) {
            if (s.nodes[child[c]].count == 0) todo.push_back({child[c], p.depth + 1u, deferred_below, me, c});
        }
    }
    return out;
}

}  // namespace Q::cuda::mesh

namespace Q::cuda::slab {

#if defined(__CUDA_ARCH__)
QUASI_SCAN_FN uint32_t bits_of(float x) { return __float_as_uint(x); }
#else
QUASI_SCAN_FN uint32_t bits_of(float x) {
    uint32_t u;
    std::memcpy(&u, &x, sizeof u);
    return u;
}
#endif
using scan::k_inf;

/// Slab test against child `c` of a wide node (lo/hi are rows of the transposed boxes).
template <class F4>
QUASI_SCAN_FN bool wide_box_hit(const F4& lx, const F4& ly, const F4& lz, const F4& hx, const F4& hy, const F4& hz, int c,
                                v3 inv, v3 oi, float closest, float& t_near) {
    struct box3 {
        float x, y, z;
    };
    auto pick = [c](const F4& v) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); };
    return box_hit(box3{pick(lx), pick(ly), pick(lz)}, box3{pick(hx), pick(hy), pick(hz)}, inv, oi, closest, t_near);
}

/// One visit of a wide node: tests the four children, returns the reference of the nearest
/// one that is hit and defers the others (with their entry distances) on `stack`, the
/// nearer ones on top; with no child hit, returns stack.after_node().
/// Branch-free: (distance, reference) pairs, a miss at distance infinity, go through a
/// five-exchange sorting network; each of the three that are not taken is stored at the top
/// of the stack whether it is a hit or not and the stack pointer moves only for a hit, so
/// the stack needs room for one entry more than it can logically hold.
/// `Stack` has push_if(bool, ref, distance bits) and after_node(); `refs` holds the four
/// references as the bits of its floats.
template <class F4, class Stack>
QUASI_SCAN_FN uint32_t wide_visit(const F4& lx, const F4& ly, const F4& lz, const F4& hx, const F4& hy, const F4& hz,
                                  const F4& refs, v3 inv, v3 oi, float closest, Stack& stack) {
    float t0, t1, t2, t3;
    const bool h0 = wide_box_hit(lx, ly, lz, hx, hy, hz, 0, inv, oi, closest, t0);
    const bool h1 = wide_box_hit(lx, ly, lz, hx, hy, hz, 1, inv, oi, closest, t1);
    const bool h2 = wide_box_hit(lx, ly, lz, hx, hy, hz, 2, inv, oi, closest, t2);
    const bool h3 = wide_box_hit(lx, ly, lz, hx, hy, hz, 3, inv, oi, closest, t3);
    t0 = h0 ? t0 : k_inf;
    t1 = h1 ? t1 : k_inf;
    t2 = h2 ? t2 : k_inf;
    t3 = h3 ? t3 : k_inf;
    uint32_t r0 = bits_of(refs.x), r1 = bits_of(refs.y), r2 = bits_of(refs.z), r3 = bits_of(refs.w);
    auto exchange = [](float& ta, float& tb, uint32_t& ra, uint32_t& rb) {
        const bool swap = tb < ta;
        const float t_lo = swap ? tb : ta, t_hi = swap ? ta : tb;
        const uint32_t r_lo = swap ? rb : ra, r_hi = swap ? ra : rb;
        ta = t_lo, tb = t_hi, ra = r_lo, rb = r_hi;
    };
    exchange(t0, t1, r0, r1);
    exchange(t2, t3, r2, r3);
    exchange(t0, t2, r0, r2);
    exchange(t1, t3, r1, r3);
    exchange(t1, t2, r1, r2);
    stack.push_if(t3 < k_inf, r3, bits_of(t3));
    stack.push_if(t2 < k_inf, r2, bits_of(t2));
    stack.push_if(t1 < k_inf, r1, bits_of(t1));
    return t0 < k_inf ? r0 : stack.after_node();
}

}  // namespace Q::cuda::slab
