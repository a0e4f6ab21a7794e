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

/// @brief Top-down 16-bin SAH build (rules: oracle/quasi_oracle_mesh.h).
///
/// Iterative: a node is finished, and its two children numbered, when it is taken
/// off the work stack; the left child is taken first. That numbers nodes in the
/// same depth-first order as a recursive build.
inline void build_bvh(scene& s) {
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

    struct work {
        uint32_t node, first, count, depth;
    };
    std::vector<work> stack;
    std::vector<uint32_t> right_side(n);
    s.nodes.assign(1, node{});
    s.nodes.reserve(static_cast<size_t>(n));
    s.depth = 0;
    uint32_t outlier_first = 0xffffffffu;  // leaf slots from here on hold the outliers
    stack.push_back({0, 0, n, 1});

    auto bin_of = [&](uint32_t t, int axis, float lo, float extent) {
        int bin = static_cast<int>(((centroid[static_cast<size_t>(t) * 3 + axis] - lo) / extent) *
                                   static_cast<float>(k_sah_bins));
        return bin > static_cast<int>(k_sah_bins) - 1 ? static_cast<int>(k_sah_bins) - 1 : bin;
    };

    while (!stack.empty()) {
        const work w = stack.back();
        stack.pop_back();
        if (w.depth > s.depth) s.depth = w.depth;

        aabb bounds, cbounds;
        for (uint32_t k = w.first; k < w.first + w.count; k++) {
            const uint32_t t = s.leaf_order[k];
            bounds.grow(prim_box[t]);
            cbounds.grow(&centroid[static_cast<size_t>(t) * 3]);
        }
        {
            node& out = s.nodes[w.node];
            for (int a = 0; a < 3; a++) {
                const float pad = (bounds.hi[a] - bounds.lo[a]) * 1e-4f + 1e-5f;
                out.bmin[a] = bounds.lo[a] - pad;
                out.bmax[a] = bounds.hi[a] + pad;
            }
            if (w.count <= (w.first >= outlier_first ? 1u : k_leaf_max)) {
                out.left_first = w.first;
                out.count = w.count;
                continue;
            }
        }

        // Root only: triangles whose box is larger than 1/64 of the scene's box (walls, floor)
        // go to the right subtree as they are, in leaves of one; everything else goes left. A
        // greedy SAH never isolates them by itself and each would inflate one box per level
        // (measured: 69 box tests per ray without this rule, 29 with it).
        if (w.node == 0) {
            const float limit = bounds.area() * k_outlier_fraction;
            uint32_t nl = 0, nr = 0;
            for (uint32_t k = w.first; k < w.first + w.count; k++) {  // stable partition
                const uint32_t t = s.leaf_order[k];
                if (prim_box[t].area() > limit) {
                    right_side[nr++] = t;
                } else {
                    s.leaf_order[w.first + nl++] = t;
                }
            }
            std::memcpy(&s.leaf_order[w.first + nl], right_side.data(), static_cast<size_t>(nr) * sizeof(uint32_t));
            if (nl > 0 && nr > 0) {
                outlier_first = w.first + nl;
                const uint32_t left = static_cast<uint32_t>(s.nodes.size());
                s.nodes.push_back(node{});
                s.nodes.push_back(node{});
                s.nodes[w.node].left_first = left;
                s.nodes[w.node].count = 0;
                stack.push_back({left + 1u, w.first + nl, nr, w.depth + 1u});
                stack.push_back({left, w.first, nl, w.depth + 1u});  // taken first
                continue;
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
            for (uint32_t k = w.first; k < w.first + w.count; k++) {
                const uint32_t t = s.leaf_order[k];
                const int bin = bin_of(t, axis, lo, extent);
                bin_count[bin]++;
                bin_box[bin].grow(prim_box[t]);
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

        uint32_t left_count = w.count / 2u;  // all centroids coincide: split by position
        if (best_axis >= 0) {
            const float lo = cbounds.lo[best_axis];
            const float extent = cbounds.hi[best_axis] - lo;
            uint32_t nl = 0, nr = 0;
            for (uint32_t k = w.first; k < w.first + w.count; k++) {  // stable partition
                const uint32_t t = s.leaf_order[k];
                if (static_cast<uint32_t>(bin_of(t, best_axis, lo, extent)) < best_split) {
                    s.leaf_order[w.first + nl++] = t;
                } else {
                    right_side[nr++] = t;
                }
            }
            std::memcpy(&s.leaf_order[w.first + nl], right_side.data(), static_cast<size_t>(nr) * sizeof(uint32_t));
            left_count = nl;
        }

        const uint32_t left = static_cast<uint32_t>(s.nodes.size());
        s.nodes.push_back(node{});
        s.nodes.push_back(node{});
        s.nodes[w.node].left_first = left;
        s.nodes[w.node].count = 0;
        stack.push_back({left + 1u, w.first + left_count, w.count - left_count, w.depth + 1u});
        stack.push_back({left, w.first, left_count, w.depth + 1u});  // taken first
    }
}

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
        tri_e2[slot] = {t.e2[0], t.e2[1], t.e2[2], 0.0f};
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
