/// @file mesh_scene.hpp
/// @brief The procedural triangle scene of BASELINE config 3 and its SAH BVH (host side).
///
/// Nothing in the reference corresponds to this file: its scene intersection is a
/// linear scan over at most 32 quads (/root/reference/backends/metal/plugin.mm:
/// 199-212). BASELINE.json asks for a 1 M-triangle procedural scene behind a SAH
/// BVH, so the scene, the builder and the node layout are defined by this
/// repository; oracle/quasi_oracle_mesh.h states the contract (mesh arithmetic,
/// builder rules, padding, tie rule) and holds an independent C implementation
/// that tests compare with this one bit for bit.
///
/// Scene ("terrain box"): the Cornell room's six quads as 12 triangles, then a
/// height-field terrain of grid x grid cells (two triangles each) whose heights
/// are three octaves of value noise from the shader's pcg_hash. Only + - * / on
/// binary32 are used, so the mesh is the same bits on every machine.
///
/// Device layout (uploaded by plugin.cu): 32-byte nodes as two float4
/// {min.xyz, left_first} {max.xyz, count}; triangles in LEAF ORDER as three
/// float4 arrays {v0.xyz, material} {e1.xyz, original index} {e2.xyz, 0}, so a
/// node is two 16-byte loads, a pair of sibling nodes one 64-byte line, and a
/// triangle three 16-byte loads; plus, per triangle, a 48-byte shading record
/// (normal and cosine-sampling frame) read once per hit.

#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <future>
#include <vector>

namespace Q::cuda::mesh {

struct float4_t {
    float x, y, z, w;
};

inline float bits_to_float(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

struct triangle {
    float v0[3];
    uint32_t material;
    float e1[3];
    uint32_t pad0;
    float e2[3];
    uint32_t pad1;
};
static_assert(sizeof(triangle) == 48);

struct node {
    float bmin[3];
    uint32_t left_first;  ///< interior: left child (right = +1); leaf: first slot in leaf order
    float bmax[3];
    uint32_t count;       ///< 0 = interior
};
static_assert(sizeof(node) == 32);

struct material {
    float albedo[3];
    float emission[3];
};

inline constexpr uint32_t k_material_count = 5;
inline constexpr uint32_t k_slot_bits = 28;  ///< a recorded hit: slot in leaf order | material << k_slot_bits (fewer than 2^28 triangles, 8 materials)
inline constexpr uint32_t k_sah_bins = 16;
inline constexpr uint32_t k_leaf_max = 2;
/// Triangles whose box area exceeds this fraction of the scene's are split off at the root.
inline constexpr float k_outlier_fraction = 1.0f / 64.0f;

struct scene {
    std::vector<triangle> triangles;   ///< original order
    std::vector<node> nodes;
    std::vector<uint32_t> leaf_order;  ///< leaf slot -> original triangle index
    material materials[k_material_count];
    uint32_t depth = 0;                ///< longest root-to-leaf path, in nodes
};

namespace detail {

/// plugin.mm:98-102
inline uint32_t pcg_hash(uint32_t input) {
    uint32_t state = input * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}

inline float lattice_value(uint32_t a, uint32_t b, uint32_t octave, uint32_t seed) {
    return static_cast<float>(pcg_hash(a + b * 64u + octave * 4096u + seed * 65536u)) *
           2.3283064365386963e-10f;
}

inline float height_at(uint32_t i, uint32_t j, uint32_t grid, uint32_t seed) {
    constexpr float amplitude[3] = {0.35f, 0.15f, 0.07f};
    const float u = static_cast<float>(i) / static_cast<float>(grid);
    const float v = static_cast<float>(j) / static_cast<float>(grid);
    float h = 0.03f;
    for (uint32_t octave = 0; octave < 3; octave++) {
        const uint32_t freq = 4u << octave;
        const float fx = u * static_cast<float>(freq);
        const float fz = v * static_cast<float>(freq);
        uint32_t ix = static_cast<uint32_t>(fx);
        uint32_t iz = static_cast<uint32_t>(fz);
        if (ix >= freq) ix = freq - 1u;
        if (iz >= freq) iz = freq - 1u;
        const float tx = fx - static_cast<float>(ix);
        const float tz = fz - static_cast<float>(iz);
        const float sx = tx * tx * (3.0f - 2.0f * tx);
        const float sz = tz * tz * (3.0f - 2.0f * tz);
        const float l00 = lattice_value(ix, iz, octave, seed);
        const float l10 = lattice_value(ix + 1u, iz, octave, seed);
        const float l01 = lattice_value(ix, iz + 1u, octave, seed);
        const float l11 = lattice_value(ix + 1u, iz + 1u, octave, seed);
        const float n0 = l00 + (l10 - l00) * sx;
        const float n1 = l01 + (l11 - l01) * sx;
        const float n = n0 + (n1 - n0) * sz;
        h = h + amplitude[octave] * n;
    }
    return h;
}

struct point {
    float c[3];
};

inline point grid_vertex(uint32_t i, uint32_t j, uint32_t grid, uint32_t seed) {
    return {{-1.0f + 2.0f * (static_cast<float>(i) / static_cast<float>(grid)), height_at(i, j, grid, seed),
             -1.0f + 2.0f * (static_cast<float>(j) / static_cast<float>(grid))}};
}

inline triangle make_triangle(const point& v0, const point& a, const point& b, bool edges_given,
                              uint32_t material) {
    triangle t{};
    for (int k = 0; k < 3; k++) {
        t.v0[k] = v0.c[k];
        t.e1[k] = edges_given ? a.c[k] : a.c[k] - v0.c[k];
        t.e2[k] = edges_given ? b.c[k] : b.c[k] - v0.c[k];
    }
    t.material = material;
    return t;
}

struct aabb {
    float lo[3] = {1e30f, 1e30f, 1e30f};
    float hi[3] = {-1e30f, -1e30f, -1e30f};
    void grow(const float p[3]) {
        for (int a = 0; a < 3; a++) {
            if (p[a] < lo[a]) lo[a] = p[a];
            if (p[a] > hi[a]) hi[a] = p[a];
        }
    }
    void grow(const aabb& o) {
        grow(o.lo);
        grow(o.hi);
    }
    [[nodiscard]] float area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.0f * ((dx * dy + dy * dz) + dz * dx);
    }
};

}  // namespace detail

/// @brief Generates the triangle list: 12 room triangles, then 2 per terrain cell, row by row.
inline void generate_triangles(uint32_t grid, uint32_t seed, scene& out) {
    using namespace detail;
    const material mats[k_material_count] = {
        {{0.73f, 0.73f, 0.73f}, {0, 0, 0}},  // white
        {{0.65f, 0.05f, 0.05f}, {0, 0, 0}},  // red
        {{0.12f, 0.45f, 0.15f}, {0, 0, 0}},  // green
        {{0, 0, 0}, {15.0f, 15.0f, 15.0f}},  // light
        {{0.6f, 0.55f, 0.4f}, {0, 0, 0}},    // terrain
    };
    std::memcpy(out.materials, mats, sizeof mats);

    // floor, ceiling, back, left, right, light (cornell_box.hpp:103-121): origin, u, v, material
    struct quad {
        point o, u, v;
        uint32_t m;
    };
    const quad room[6] = {
        {{{-1, 0, -1}}, {{2, 0, 0}}, {{0, 0, 2}}, 0},  {{{-1, 2, 1}}, {{2, 0, 0}}, {{0, 0, -2}}, 0},
        {{{-1, 0, -1}}, {{2, 0, 0}}, {{0, 2, 0}}, 0},  {{{-1, 0, 1}}, {{0, 0, -2}}, {{0, 2, 0}}, 1},
        {{{1, 0, -1}}, {{0, 0, 2}}, {{0, 2, 0}}, 2},   {{{-0.25f, 1.99f, -0.25f}}, {{0.5f, 0, 0}}, {{0, 0, 0.5f}}, 3},
    };
    out.triangles.clear();
    out.triangles.reserve(12u + 2u * static_cast<size_t>(grid) * grid);
    for (const quad& q : room) {
        point diag{{q.u.c[0] + q.v.c[0], q.u.c[1] + q.v.c[1], q.u.c[2] + q.v.c[2]}};
        out.triangles.push_back(make_triangle(q.o, q.u, diag, true, q.m));
        out.triangles.push_back(make_triangle(q.o, diag, q.v, true, q.m));
    }
    for (uint32_t j = 0; j < grid; j++) {
        for (uint32_t i = 0; i < grid; i++) {
            const point p00 = grid_vertex(i, j, grid, seed), p10 = grid_vertex(i + 1u, j, grid, seed);
            const point p01 = grid_vertex(i, j + 1u, grid, seed), p11 = grid_vertex(i + 1u, j + 1u, grid, seed);
            out.triangles.push_back(make_triangle(p00, p01, p11, false, 4u));
            out.triangles.push_back(make_triangle(p00, p11, p10, false, 4u));
        }
    }
}

namespace detail {

/// Read-only inputs of the build plus the permutation it sorts in place.
struct build_context {
    const std::vector<aabb>& prim_box;
    const std::vector<float>& centroid;  ///< 3 per triangle
    std::vector<uint32_t>& leaf_order;
};

/// A finished subtree, numbered as if it were the whole tree: `root` is its node 0 and
/// below[i - 1] its node i, the children of its j-th interior node (depth first, left
/// before right) at 1 + 2 j and 2 + 2 j. That numbering is what makes the build parallel:
/// placing the subtree where its root is the k-th interior node of a bigger tree adds 2 k
/// to every index (embed()), so subtrees can be built anywhere, in any order, and the
/// result is the array a sequential depth-first build writes.
struct subtree {
    node root{};
    std::vector<node> below;
    uint32_t depth = 0;  ///< longest root-to-leaf path in nodes
    [[nodiscard]] uint32_t interior_count() const { return static_cast<uint32_t>(below.size() / 2); }
};

inline node embed(node n, uint32_t rank_offset) {
    if (n.count == 0) n.left_first += 2u * rank_offset;
    return n;
}

inline int bin_of(const build_context& c, uint32_t t, int axis, float lo, float extent) {
    const int bin = static_cast<int>(((c.centroid[static_cast<size_t>(t) * 3 + axis] - lo) / extent) *
                                     static_cast<float>(k_sah_bins));
    return bin > static_cast<int>(k_sah_bins) - 1 ? static_cast<int>(k_sah_bins) - 1 : bin;
}

/// Bounds of a range, its padded node box, and (unless it becomes a leaf or is split by the
/// outlier rule) the SAH partition of leaf_order[first, first + count). Returns the number of
/// triangles that went left, 0 for a leaf. `scratch` holds at least `count` entries.
inline uint32_t split_range(const build_context& c, uint32_t first, uint32_t count, bool is_root, uint32_t leaf_max,
                            std::vector<uint32_t>& scratch, node& out, bool& right_is_outliers) {
    right_is_outliers = false;
    aabb bounds, cbounds;
    for (uint32_t k = first; k < first + count; k++) {
        const uint32_t t = c.leaf_order[k];
        bounds.grow(c.prim_box[t]);
        cbounds.grow(&c.centroid[static_cast<size_t>(t) * 3]);
    }
    for (int a = 0; a < 3; a++) {
        const float pad = (bounds.hi[a] - bounds.lo[a]) * 1e-4f + 1e-5f;
        out.bmin[a] = bounds.lo[a] - pad;
        out.bmax[a] = bounds.hi[a] + pad;
    }
    if (count <= leaf_max) {
        out.left_first = first;
        out.count = count;
        return 0;
    }
    out.count = 0;
    out.left_first = 1;  // local numbering: the children of a subtree's root are its nodes 1 and 2

    // Root only: triangles whose box is larger than 1/64 of the scene's box (walls, floor)
    // go to the right subtree as they are, in leaves of one; everything else goes left. A
    // greedy SAH never isolates them by itself and each would inflate one box per level
    // (measured: 69 box tests per ray without this rule, 29 with it).
    if (is_root) {
        const float limit = bounds.area() * k_outlier_fraction;
        uint32_t nl = 0, nr = 0;
        for (uint32_t k = first; k < first + count; k++) {  // stable partition
            const uint32_t t = c.leaf_order[k];
            if (c.prim_box[t].area() > limit) {
                scratch[nr++] = t;
            } else {
                c.leaf_order[first + nl++] = t;
            }
        }
        std::memcpy(&c.leaf_order[first + nl], scratch.data(), static_cast<size_t>(nr) * sizeof(uint32_t));
        if (nl > 0 && nr > 0) {
            right_is_outliers = true;
            return nl;
        }
    }

    int best_axis = -1;
    uint32_t best_split = 0;
    float best_cost = 1e30f;
    for (int axis = 0; axis < 3; axis++) {
        const float lo = cbounds.lo[axis];
        const float extent = cbounds.hi[axis] - lo;
        if (!(extent > 0.0f)) continue;
        aabb bin_box[k_sah_bins];
        uint32_t bin_count[k_sah_bins] = {};
        for (uint32_t k = first; k < first + count; k++) {
            const uint32_t t = c.leaf_order[k];
            const int bin = bin_of(c, t, axis, lo, extent);
            bin_count[bin]++;
            bin_box[bin].grow(c.prim_box[t]);
        }
        aabb suffix_box[k_sah_bins];
        uint32_t suffix_count[k_sah_bins] = {};
        aabb acc;
        uint32_t seen = 0;
        for (int i = static_cast<int>(k_sah_bins) - 1; i >= 1; i--) {
            acc.grow(bin_box[i]);
            seen += bin_count[i];
            suffix_box[i] = acc;
            suffix_count[i] = seen;
        }
        acc = aabb{};
        seen = 0;
        for (uint32_t split = 1; split < k_sah_bins; split++) {
            acc.grow(bin_box[split - 1]);
            seen += bin_count[split - 1];
            if (seen == 0 || suffix_count[split] == 0) continue;
            const float cost = acc.area() * static_cast<float>(seen) +
                               suffix_box[split].area() * static_cast<float>(suffix_count[split]);
            if (cost < best_cost) {
                best_cost = cost;
                best_axis = axis;
                best_split = split;
            }
        }
    }

    if (best_axis < 0) return count / 2u;  // all centroids coincide: split by position
    const float lo = cbounds.lo[best_axis];
    const float extent = cbounds.hi[best_axis] - lo;
    uint32_t nl = 0, nr = 0;
    for (uint32_t k = first; k < first + count; k++) {  // stable partition
        const uint32_t t = c.leaf_order[k];
        if (static_cast<uint32_t>(bin_of(c, t, best_axis, lo, extent)) < best_split) {
            c.leaf_order[first + nl++] = t;
        } else {
            scratch[nr++] = t;
        }
    }
    std::memcpy(&c.leaf_order[first + nl], scratch.data(), static_cast<size_t>(nr) * sizeof(uint32_t));
    return nl;
}

/// One range, one thread: iterative, a node is finished and its two children numbered when
/// it is taken off the work stack, the left child first (depth-first numbering).
inline subtree build_sequential(const build_context& c, uint32_t first, uint32_t count, uint32_t leaf_max) {
    struct work {
        uint32_t node, first, count, depth;
    };
    std::vector<node> nodes(1);
    nodes.reserve(2 * static_cast<size_t>(count) / (leaf_max > 1 ? leaf_max / 2 + 1 : 1) + 8);
    std::vector<uint32_t> scratch(count);
    std::vector<work> stack{{0, first, count, 1}};
    subtree out;
    while (!stack.empty()) {
        const work w = stack.back();
        stack.pop_back();
        if (w.depth > out.depth) out.depth = w.depth;
        node n{};
        bool unused = false;
        const uint32_t left_count = split_range(c, w.first, w.count, false, leaf_max, scratch, n, unused);
        if (left_count != 0) {
            n.left_first = static_cast<uint32_t>(nodes.size());
            nodes.push_back(node{});
            nodes.push_back(node{});
            stack.push_back({n.left_first + 1u, w.first + left_count, w.count - left_count, w.depth + 1u});
            stack.push_back({n.left_first, w.first, left_count, w.depth + 1u});  // taken first
        }
        nodes[w.node] = n;
    }
    out.root = nodes[0];
    out.below.assign(nodes.begin() + 1, nodes.end());
    return out;
}

/// Large ranges: split here, build the two halves concurrently, join them by embed().
inline subtree build_parallel(const build_context& c, uint32_t first, uint32_t count, bool is_root, uint32_t leaf_max,
                              uint32_t task_limit);

}  // namespace detail

/// @brief Top-down 16-bin SAH build (rules: oracle/quasi_oracle_mesh.h).
///
/// The result is the node array and leaf order of a sequential depth-first build (the
/// contract tests compare with the oracle's, byte for byte), produced on every host core:
/// ranges above 32 K triangles are split on the calling thread and their halves built
/// concurrently, smaller ones are built by one thread each, and subtrees are joined by the
/// index arithmetic of detail::subtree. 2.6 s on one core, a few hundred ms on sixteen.
/// @param task_limit Ranges of at most this many triangles are built by one thread (tests lower it
///        to push small meshes through the concurrent path; the result does not depend on it).
inline void build_bvh(scene& s, uint32_t task_limit = 32768) {
    using detail::aabb;
    const uint32_t n = static_cast<uint32_t>(s.triangles.size());
    std::vector<aabb> prim_box(n);
    std::vector<float> centroid(static_cast<size_t>(n) * 3);
    s.leaf_order.resize(n);
    for (uint32_t t = 0; t < n; t++) {
        const triangle& tri = s.triangles[t];
        float p1[3], p2[3];
        for (int a = 0; a < 3; a++) {
            p1[a] = tri.v0[a] + tri.e1[a];
            p2[a] = tri.v0[a] + tri.e2[a];
            centroid[static_cast<size_t>(t) * 3 + a] = ((tri.v0[a] + p1[a]) + p2[a]) * (1.0f / 3.0f);
        }
        prim_box[t].grow(tri.v0);
        prim_box[t].grow(p1);
        prim_box[t].grow(p2);
        s.leaf_order[t] = t;
    }
    const detail::build_context context{prim_box, centroid, s.leaf_order};
    detail::subtree tree = detail::build_parallel(context, 0, n, true, k_leaf_max, task_limit);
    s.nodes.clear();
    s.nodes.reserve(tree.below.size() + 1);
    s.nodes.push_back(tree.root);
    s.nodes.insert(s.nodes.end(), tree.below.begin(), tree.below.end());
    s.depth = tree.depth;
}

namespace detail {

inline subtree build_parallel(const build_context& c, uint32_t first, uint32_t count, bool is_root, uint32_t leaf_max,
                              uint32_t task_limit) {
    if (!is_root && count <= task_limit) return build_sequential(c, first, count, leaf_max);
    subtree out;
    std::vector<uint32_t> scratch(count);
    bool right_is_outliers = false;
    const uint32_t left_count = split_range(c, first, count, is_root, leaf_max, scratch, out.root, right_is_outliers);
    scratch = std::vector<uint32_t>();
    out.depth = 1;
    if (left_count == 0) return out;  // a leaf
    auto left_future = std::async(std::launch::async, [&c, first, left_count, leaf_max, task_limit] {
        return build_parallel(c, first, left_count, false, leaf_max, task_limit);
    });
    // the outliers split off at the root are built in leaves of one (oracle/quasi_oracle_mesh.h)
    const subtree right = build_parallel(c, first + left_count, count - left_count, false,
                                         right_is_outliers ? 1u : leaf_max, task_limit);
    const subtree left = left_future.get();
    // this subtree's interior nodes in depth-first order: its root (rank 0), the left
    // subtree's (from rank 1), the right subtree's (from rank 1 + left.interior_count())
    const uint32_t left_rank = 1, right_rank = 1 + left.interior_count();
    out.below.reserve(2 + left.below.size() + right.below.size());
    out.below.push_back(embed(left.root, left_rank));
    out.below.push_back(embed(right.root, right_rank));
    for (const node& x : left.below) out.below.push_back(embed(x, left_rank));
    for (const node& x : right.below) out.below.push_back(embed(x, right_rank));
    out.depth = 1 + (left.depth > right.depth ? left.depth : right.depth);
    return out;
}

}  // namespace detail

namespace detail {

struct vec3 {
    float x, y, z;
};
inline vec3 cross3(vec3 a, vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
/// v / sqrt((v.x*v.x + v.y*v.y) + v.z*v.z): the reference's normalize (math/vec.hpp:92-114).
inline vec3 normalize3(vec3 v) {
    const float len = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return {v.x / len, v.y / len, v.z / len};
}

}  // namespace detail

/// @brief What shading a hit on a triangle needs, none of which depends on the ray
/// except for one sign: the geometric normal ng = normalize(cross(e1, e2)) and the
/// cosine-sampling frame of plugin.mm:115-121 for n = +ng, i.e. w = normalize(ng),
/// a = |w.x| > 0.9 ? (0,1,0) : (1,0,0), v = normalize(cross(w, a)), u = cross(w, v).
/// Each operation is one correctly rounded binary32 operation in the shader's
/// order (this file is compiled with -ffp-contract=off), so the values are the
/// ones the per-hit evaluation of oracle/quasi_oracle_mesh.c produces. For
/// n = -ng the frame is (-w, -v, u), exactly (negation commutes with every step).
/// Packed as three float4: {ng.xyz, w.x} {w.y, w.z, v.x, v.y} {v.z, u.xyz}.
inline void shading_record(const triangle& t, float4_t out[3]) {
    using namespace detail;
    const vec3 ng = normalize3(cross3({t.e1[0], t.e1[1], t.e1[2]}, {t.e2[0], t.e2[1], t.e2[2]}));
    const vec3 w = normalize3(ng);
    const vec3 a = std::fabs(w.x) > 0.9f ? vec3{0.0f, 1.0f, 0.0f} : vec3{1.0f, 0.0f, 0.0f};
    const vec3 v = normalize3(cross3(w, a));
    const vec3 u = cross3(w, v);
    out[0] = {ng.x, ng.y, ng.z, w.x};
    out[1] = {w.y, w.z, v.x, v.y};
    out[2] = {v.z, u.x, u.y, u.z};
}

/// @brief Flattens nodes and leaf-ordered triangles into the float4 arrays the kernel reads.
/// `tri_shade` gets three float4 per leaf slot (shading_record).
inline void pack_for_device(const scene& s, std::vector<float4_t>& nodes, std::vector<float4_t>& tri_v0,
                            std::vector<float4_t>& tri_e1, std::vector<float4_t>& tri_e2,
                            std::vector<float4_t>& tri_shade) {
    nodes.resize(s.nodes.size() * 2);
    for (size_t i = 0; i < s.nodes.size(); i++) {
        const node& n = s.nodes[i];
        nodes[2 * i] = {n.bmin[0], n.bmin[1], n.bmin[2], bits_to_float(n.left_first)};
        nodes[2 * i + 1] = {n.bmax[0], n.bmax[1], n.bmax[2], bits_to_float(n.count)};
    }
    const size_t n = s.triangles.size();
    tri_v0.resize(n);
    tri_e1.resize(n);
    tri_e2.resize(n);
    tri_shade.resize(n * 3);
    for (size_t slot = 0; slot < n; slot++) {
        const uint32_t original = s.leaf_order[slot];
        const triangle& t = s.triangles[original];
        tri_v0[slot] = {t.v0[0], t.v0[1], t.v0[2], bits_to_float(t.material)};
        tri_e1[slot] = {t.e1[0], t.e1[1], t.e1[2], bits_to_float(original)};
        // what a hit is recorded as: the slot, with the material in the top four bits (k_slot_bits), so that shading
        // a hit needs the shading record only
        tri_e2[slot] = {t.e2[0], t.e2[1], t.e2[2], bits_to_float(static_cast<uint32_t>(slot) | (t.material << k_slot_bits))};
        shading_record(t, &tri_shade[slot * 3]);
    }
}

/// @brief The triangles in ORIGINAL order as the float4 arrays the GPU builder reads.
inline void pack_original_order(const scene& s, std::vector<float4_t>& tri_v0, std::vector<float4_t>& tri_e1,
                                std::vector<float4_t>& tri_e2, std::vector<float4_t>& tri_shade) {
    const size_t n = s.triangles.size();
    tri_v0.resize(n);
    tri_e1.resize(n);
    tri_e2.resize(n);
    tri_shade.resize(n * 3);
    for (size_t i = 0; i < n; i++) {
        const triangle& t = s.triangles[i];
        tri_v0[i] = {t.v0[0], t.v0[1], t.v0[2], bits_to_float(t.material)};
        tri_e1[i] = {t.e1[0], t.e1[1], t.e1[2], 0.0f};
        tri_e2[i] = {t.e2[0], t.e2[1], t.e2[2], 0.0f};
        shading_record(t, &tri_shade[i * 3]);
    }
}

}  // namespace Q::cuda::mesh
