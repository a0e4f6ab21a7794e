// Host check of the mesh traversal's pruning (quasi_b200/csrc/slab_test.hpp, the code the
// CUDA kernel compiles) on the backend's own scene and BVH (mesh_scene.hpp): for every ray
// the closest hit found THROUGH the tree, skipping what the slab test says can be skipped,
// must be the closest hit of a brute-force loop over all triangles, with the same
// Moeller-Trumbore test (operation order of oracle/quasi_oracle_mesh.c:qm_intersect) and
// the same tie rule. Compiled with -DQUASI_SCAN_CHECK_PERTURB the reciprocal is two ulps
// off either way, twice what MUFU.RCP may be, and a * b + c is rounded twice.
// Rays: from points above the terrain in random directions, bounce rays from points on
// random triangles, directions almost or exactly parallel to an axis (where 1/dir is huge
// or clamped), origins exactly on box planes of the tree.
// With a third argument "wide" the same rays go through the 4-wide tree collapsed from the
// binary one (wide_bvh.hpp: collapse_wide and wide_visit, the code the kernel compiles), after
// a structural check of that tree: every leaf slot reachable exactly once, every child box
// the box of a node of the binary tree, the stack bound it reports never exceeded.
// With "quant" the rays go through the binary tree in its quantized form (quantized_bvh.hpp), after a check that every
// quantized box holds its float box with half a grid step to spare and carries the tree's references.
//   usage: mesh_prune_check [grid] [rays] [wide [stack limit] | quant]   -> prints counts, exit code 1 on any mismatch
#include "mesh_scene.hpp"
#include "quantized_bvh.hpp"
#include "slab_test.hpp"
#include "wide_bvh.hpp"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <utility>
#include <vector>

using namespace Q::cuda;

namespace {

struct hit {
    float t = 1e30f;
    uint32_t prim = 0xffffffffu;
};

struct f4 {
    float x, y, z, w;
};

void intersect(const mesh::triangle& tri, uint32_t prim, v3 o, v3 d, hit& best) {
    const v3 e1{tri.e1[0], tri.e1[1], tri.e1[2]}, e2{tri.e2[0], tri.e2[1], tri.e2[2]};
    const v3 pvec{d.y * e2.z - d.z * e2.y, d.z * e2.x - d.x * e2.z, d.x * e2.y - d.y * e2.x};
    const float det = e1.x * pvec.x + e1.y * pvec.y + e1.z * pvec.z;
    if (std::fabs(det) < 1e-12f) return;
    const float inv_det = 1.0f / det;
    const v3 tvec{o.x - tri.v0[0], o.y - tri.v0[1], o.z - tri.v0[2]};
    const float u = (tvec.x * pvec.x + tvec.y * pvec.y + tvec.z * pvec.z) * inv_det;
    if (u < 0.0f || u > 1.0f) return;
    const v3 qvec{tvec.y * e1.z - tvec.z * e1.y, tvec.z * e1.x - tvec.x * e1.z, tvec.x * e1.y - tvec.y * e1.x};
    const float v = (d.x * qvec.x + d.y * qvec.y + d.z * qvec.z) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return;
    const float t = (e2.x * qvec.x + e2.y * qvec.y + e2.z * qvec.z) * inv_det;
    if (t < 0.001f || t > best.t) return;
    if (t == best.t && best.prim != 0xffffffffu && prim < best.prim) return;
    best.t = t;
    best.prim = prim;
}

hit brute_force(const mesh::scene& s, v3 o, v3 d) {
    hit best;
    for (uint32_t t = 0; t < s.triangles.size(); t++) intersect(s.triangles[t], t, o, d, best);
    return best;
}

/// The traversal of mesh_trace.cu in one thread: near child first, far child deferred with its
/// entry distance (kept at the precision of the kernel's stack: upper 16 bits), culled on pop.
/// With `grid` (mode "quant") the children are tested in their quantized form (quantized_bvh.hpp: the pairs, the
/// per-ray constants and box_hit_quantized as the wavefront trace kernel compiles them); the root stays a float box.
hit through_the_tree(const mesh::scene& s, v3 o, v3 d, uint64_t& boxes, const mesh::quantized_tree* grid = nullptr) {
    hit best;
    const v3 inv{slab::reciprocal(d.x), slab::reciprocal(d.y), slab::reciprocal(d.z)};
    const v3 oi{o.x * inv.x, o.y * inv.y, o.z * inv.z};
    slab::quant_ray qr{};
    if (grid) qr = slab::make_quant_ray(grid->origin, grid->step, inv, oi);
    auto box = [&](const mesh::node& n, float& tn) {
        boxes++;
        const size_t index = static_cast<size_t>(&n - s.nodes.data());
        if (grid && index != 0) {
            const uint32_t* w = &grid->pairs[(index - 1) / 2 * mesh::k_pair_words + 3 * ((index - 1) % 2)];
            return slab::box_hit_quantized(w[0], w[1], w[2], qr, best.t, tn);
        }
        return slab::box_hit(f4{n.bmin[0], n.bmin[1], n.bmin[2], 0.0f}, f4{n.bmax[0], n.bmax[1], n.bmax[2], 0.0f}, inv, oi,
                             best.t, tn);
    };
    struct entry {
        uint32_t node;
        float t_near;
    };
    entry stack[64];
    int sp = 0;
    float tn;
    if (!box(s.nodes[0], tn)) return best;
    uint32_t cur = 0;
    for (;;) {
        const mesh::node& n = s.nodes[cur];
        bool descended = false;
        if (n.count == 0) {
            float tl, tr;
            const bool hl = box(s.nodes[n.left_first], tl), hr = box(s.nodes[n.left_first + 1], tr);
            if (hl && hr) {
                const bool go_left = tl <= tr;
                uint32_t bits;
                const float far_t = go_left ? tr : tl;
                std::memcpy(&bits, &far_t, 4);
                bits &= 0xffff0000u;  // the kernel keeps the upper 16 bits of the entry distance
                float kept;
                std::memcpy(&kept, &bits, 4);
                stack[sp++] = {go_left ? n.left_first + 1 : n.left_first, kept};
                cur = go_left ? n.left_first : n.left_first + 1;
                descended = true;
            } else if (hl || hr) {
                cur = hl ? n.left_first : n.left_first + 1;
                descended = true;
            }
        } else {
            for (uint32_t k = 0; k < n.count; k++) {
                const uint32_t prim = s.leaf_order[n.left_first + k];
                intersect(s.triangles[prim], prim, o, d, best);
            }
        }
        if (descended) continue;
        bool found = false;
        while (sp > 0) {
            sp--;
            if (stack[sp].t_near <= best.t) {
                cur = stack[sp].node;
                found = true;
                break;
            }
        }
        if (!found) break;
    }
    return best;
}

/// The stack of mesh_trace.cu as the wide visit sees it: references, and entry distances cut
/// to their upper 16 bits.
struct host_stack {
    struct entry {
        uint32_t ref;
        float t_near;
    };
    entry entries[96];
    int sp = 0, high_water = 0;
    void push_if(bool keep, uint32_t ref, uint32_t distance_bits) {
        distance_bits &= 0xffff0000u;
        float kept;
        std::memcpy(&kept, &distance_bits, 4);
        entries[sp] = {ref, kept};  // written either way, as in the kernel
        sp += keep ? 1 : 0;
        if (sp > high_water) high_water = sp;
    }
    static constexpr uint32_t k_pop = 0xfffffffeu, k_done = 0xffffffffu;
    [[nodiscard]] uint32_t after_node() const { return sp > 0 ? k_pop : k_done; }
};

/// The wide traversal of mesh_trace.cu in one thread.
hit through_the_wide_tree(const mesh::scene& s, const mesh::wide_tree& w, v3 o, v3 d, uint64_t& boxes, uint64_t& visits,
                          int& high_water) {
    hit best;
    const v3 inv{slab::reciprocal(d.x), slab::reciprocal(d.y), slab::reciprocal(d.z)};
    const v3 oi{o.x * inv.x, o.y * inv.y, o.z * inv.z};
    host_stack stack;
    uint32_t cur = 0;  // the root, a wide node
    while (cur != host_stack::k_done) {
        if (cur == host_stack::k_pop) {
            stack.sp--;
            cur = stack.entries[stack.sp].t_near <= best.t ? stack.entries[stack.sp].ref : stack.after_node();
            continue;
        }
        if ((cur & 7u) == 0u) {
            const mesh::float4_t* n = &w.nodes[static_cast<size_t>(cur >> 3) * mesh::k_wide_node_float4];
            boxes += 4;
            visits++;
            cur = slab::wide_visit(n[0], n[1], n[2], n[3], n[4], n[5], n[6], inv, oi, best.t, stack);
        } else {
            const uint32_t first = cur >> 3, count = cur & 7u;
            for (uint32_t k = 0; k < count; k++) {
                const uint32_t prim = s.leaf_order[first + k];
                intersect(s.triangles[prim], prim, o, d, best);
            }
            cur = stack.after_node();
        }
    }
    if (stack.high_water > high_water) high_water = stack.high_water;
    return best;
}

/// Every leaf slot in exactly one leaf; every child of a wide node covers the slot range of
/// one node of the binary tree and carries that node's box, bit for bit; every wide node is
/// referenced once; counts as reported. Returns the number of violations.
long check_wide_structure(const mesh::scene& s, const mesh::wide_tree& w) {
    long bad = 0;
    if (w.nodes.size() != static_cast<size_t>(w.node_count) * mesh::k_wide_node_float4) return 1;
    // slot range of every binary node -> that node (ranges are unique: both children of an interior node are non-empty)
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> by_range;
    {
        std::vector<std::pair<uint32_t, uint32_t>> range(s.nodes.size());
        for (size_t i = s.nodes.size(); i-- > 0;) {  // children have higher numbers than their parent
            const mesh::node& n = s.nodes[i];
            range[i] = n.count != 0 ? std::make_pair(n.left_first, n.count)
                                    : std::make_pair(range[n.left_first].first, range[n.left_first].second + range[n.left_first + 1].second);
            if (n.count == 0 && range[n.left_first].first + range[n.left_first].second != range[n.left_first + 1].first) bad++;
            by_range[range[i]] = static_cast<uint32_t>(i);
        }
    }
    std::vector<uint8_t> slot_seen(s.triangles.size(), 0);
    std::vector<uint8_t> node_seen(w.node_count, 0);
    std::vector<std::pair<uint32_t, uint32_t>> wide_range(w.node_count, {0u, 0u});
    uint32_t children = 0;
    // wide nodes are numbered parents first, so walking the numbers downwards sees children before parents
    for (uint32_t i = w.node_count; i-- > 0;) {
        const mesh::float4_t* n = &w.nodes[static_cast<size_t>(i) * mesh::k_wide_node_float4];
        uint32_t first = 0, count = 0;
        bool gap = false;
        for (uint32_t c = 0; c < mesh::k_wide; c++) {
            uint32_t ref;
            std::memcpy(&ref, &(&n[6].x)[c], 4);
            const bool empty = (&n[0].x)[c] == mesh::k_nowhere;
            if (empty) {
                gap = true;
                continue;
            }
            if (gap) bad++;  // occupied places come first
            children++;
            std::pair<uint32_t, uint32_t> r;
            if ((ref & 7u) != 0u) {
                r = {ref >> 3, ref & 7u};
                for (uint32_t k = 0; k < r.second; k++) {
                    if (r.first + k >= slot_seen.size() || slot_seen[r.first + k]++ != 0) bad++;
                }
            } else {
                const uint32_t child = ref >> 3;
                if (child <= i || child >= w.node_count || node_seen[child]++ != 0) {
                    bad++;
                    continue;
                }
                r = wide_range[child];
            }
            if (count != 0 && first + count != r.first) bad++;  // children in slot order, contiguous
            if (count == 0) first = r.first;
            count += r.second;
            const auto found = by_range.find(r);
            if (found == by_range.end()) {
                bad++;
                continue;
            }
            const mesh::node& b = s.nodes[found->second];
            for (int a = 0; a < 3; a++) {
                if ((&n[a].x)[c] != b.bmin[a] || (&n[3 + a].x)[c] != b.bmax[a]) bad++;
            }
            if (((ref & 7u) != 0u) != (b.count != 0)) bad++;
        }
        wide_range[i] = {first, count};
    }
    if (w.node_count != 0 && wide_range[0] != std::make_pair(0u, static_cast<uint32_t>(s.triangles.size()))) bad++;
    for (uint8_t seen : slot_seen) bad += seen != 1;
    for (uint32_t i = 1; i < w.node_count; i++) bad += node_seen[i] != 1;
    if (children != w.children) bad++;
    return bad;
}

}  // namespace

int main(int argc, char** argv) {
    const uint32_t grid = argc > 1 ? static_cast<uint32_t>(std::atoi(argv[1])) : 40u;
    const long rays = argc > 2 ? std::atol(argv[2]) : 200000;
    const bool wide = argc > 3 && std::strcmp(argv[3], "wide") == 0;
    const bool quant = argc > 3 && std::strcmp(argv[3], "quant") == 0;
    mesh::scene s;
    mesh::generate_triangles(grid, 1, s);
    mesh::build_bvh(s);
    mesh::wide_tree w;
    long structure = 0;
    uint64_t visits = 0;
    int high_water = 0;
    if (wide) {
        const uint32_t stack_limit = argc > 4 ? static_cast<uint32_t>(std::atoi(argv[4])) : 31u;
        w = mesh::collapse_wide(s, stack_limit);
        structure = check_wide_structure(s, w) + (w.stack_need > stack_limit);
        std::printf("wide tree: %u nodes (binary: %zu, %zu interior), %.2f children per node, depth %u (binary %u), stack bound %u, "
                    "structure violations %ld\n",
                    w.node_count, s.nodes.size(), s.nodes.size() / 2, static_cast<double>(w.children) / w.node_count, w.depth,
                    s.depth, w.stack_need, structure);
    }

    mesh::quantized_tree lattice;
    if (quant) {
        lattice = mesh::quantize_pairs(s);
        // every quantized box holds the float box it stands for, with room, and the references are the tree's
        long outside = lattice.pairs.empty() ? 1 : 0;
        for (size_t i = 1; i < s.nodes.size() && !lattice.pairs.empty(); i++) {
            const uint32_t* w = &lattice.pairs[(i - 1) / 2 * mesh::k_pair_words];
            const uint32_t* b = w + 3 * ((i - 1) % 2);
            for (int a = 0; a < 3; a++) {
                const double lo = lattice.origin[a] + static_cast<double>(b[a] & 0xffffu) * lattice.step[a];
                const double hi = lattice.origin[a] + static_cast<double>(b[a] >> 16) * lattice.step[a];
                if (!(lo <= s.nodes[i].bmin[a] - 0.5 * lattice.step[a]) || !(hi >= s.nodes[i].bmax[a] + 0.5 * lattice.step[a])) outside++;
            }
            if (w[6 + (i - 1) % 2] != ((s.nodes[i].left_first << 3) | s.nodes[i].count)) outside++;
        }
        structure += outside;
        std::printf("quantized tree: %u pairs, grid step %.3g %.3g %.3g, structure violations %ld\n", lattice.pair_count,
                    lattice.step[0], lattice.step[1], lattice.step[2], outside);
    }

    std::mt19937 rng(11);
    std::uniform_real_distribution<float> uniform(0.0f, 1.0f);
    auto coin = [&] { return (rng() & 1u) != 0u; };
    long mismatches = 0, hits = 0;
    uint64_t boxes = 0;
    for (long it = 0; it < rays; it++) {
        v3 o{}, d{};
        const int kind = static_cast<int>(it % 4);
        float len2;
        do {
            d = {2.0f * uniform(rng) - 1.0f, 2.0f * uniform(rng) - 1.0f, 2.0f * uniform(rng) - 1.0f};
            len2 = d.x * d.x + d.y * d.y + d.z * d.z;
        } while (len2 > 1.0f || len2 < 1e-4f);
        if (kind == 0) {  // anywhere in the room above the terrain, or the camera's place
            o = coin() ? v3{0.0f, 1.0f, 3.5f} : v3{1.9f * uniform(rng) - 0.95f, 0.7f + 1.25f * uniform(rng), 1.9f * uniform(rng) - 0.95f};
            if (o.z > 2.0f) d = {0.6f * (d.x), 0.5f * d.y, -1.0f};
        } else {  // a bounce: a point of a random triangle, lifted 0.001 along its normal
            const mesh::triangle& tri = s.triangles[rng() % s.triangles.size()];
            float a = uniform(rng), b = uniform(rng);
            if (a + b > 1.0f) a = 1.0f - a, b = 1.0f - b;
            const v3 n{tri.e1[1] * tri.e2[2] - tri.e1[2] * tri.e2[1], tri.e1[2] * tri.e2[0] - tri.e1[0] * tri.e2[2],
                       tri.e1[0] * tri.e2[1] - tri.e1[1] * tri.e2[0]};
            const float nl = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
            float side = (d.x * n.x + d.y * n.y + d.z * n.z) < 0.0f ? -1.0f : 1.0f;
            o = {tri.v0[0] + tri.e1[0] * a + tri.e2[0] * b + n.x / nl * 0.001f * side,
                 tri.v0[1] + tri.e1[1] * a + tri.e2[1] * b + n.y / nl * 0.001f * side,
                 tri.v0[2] + tri.e1[2] * a + tri.e2[2] * b + n.z / nl * 0.001f * side};
        }
        if (kind == 2) {  // almost or exactly parallel to an axis plane: 1/dir huge, infinite (clamped) or denormal
            float* comp = (rng() % 3u == 0) ? &d.x : (coin() ? &d.y : &d.z);
            const uint32_t how = rng() % 4u;
            *comp = how == 0 ? 0.0f : (how == 1 ? -0.0f : (how == 2 ? (uniform(rng) - 0.5f) * 1e-7f : (uniform(rng) - 0.5f) * 1e-38f));
        }
        if (kind == 3) {  // an origin exactly in a box plane of the tree
            const mesh::node& n = s.nodes[rng() % s.nodes.size()];
            const uint32_t axis = rng() % 3u;
            (axis == 0 ? o.x : (axis == 1 ? o.y : o.z)) = coin() ? n.bmin[axis] : n.bmax[axis];
        }
        const float len = std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
        if (!(len > 0.0f)) continue;
        d = {d.x / len, d.y / len, d.z / len};
        const hit want = brute_force(s, o, d);
        const hit got = wide ? through_the_wide_tree(s, w, o, d, boxes, visits, high_water)
                             : through_the_tree(s, o, d, boxes, quant ? &lattice : nullptr);
        hits += want.prim != 0xffffffffu;
        if (got.prim != want.prim || got.t != want.t) {
            if (mismatches < 10) {
                std::printf("MISMATCH tree %u (t %.9g) brute force %u (t %.9g) o %.9g %.9g %.9g d %.9g %.9g %.9g\n", got.prim,
                            got.t, want.prim, want.t, o.x, o.y, o.z, d.x, d.y, d.z);
            }
            mismatches++;
        }
    }
    std::printf("rays %ld hits %ld boxes_per_ray %.1f mismatches %ld\n", rays, hits, static_cast<double>(boxes) / rays, mismatches);
    if (wide) {
        std::printf("wide visits_per_ray %.1f deepest stack %d (bound %u)\n", static_cast<double>(visits) / rays, high_water,
                    w.stack_need);
        if (high_water > static_cast<int>(w.stack_need)) structure++;
    }
    return mismatches == 0 && structure == 0 ? 0 : 1;
}
