/// @file mesh_traverse.cuh
/// @brief What one lane does to walk the BVH of the mesh scene: the per-lane stack in shared memory, the
/// triangle test in the oracle's operation order, one step down the binary or the 4-wide tree. Shared by the
/// one-kernel path tracer (mesh_trace.cu) and the wavefront form (mesh_wavefront.cu); DESIGN.md section 4.3.
#pragma once

#include "device_common.cuh"
#include "kernels.hpp"
#include "quantized_bvh.hpp"
#include "slab_test.hpp"
#include "wide_bvh.hpp"

namespace Q::cuda {
namespace {

using namespace dev;

constexpr int k_block_threads = 256;  ///< of every kernel that walks the tree: the stride of the stack arrays
constexpr size_t k_stack_bytes = static_cast<size_t>(k_mesh_stack_entries) * k_block_threads * 6;
constexpr uint32_t k_miss = 0xffffffffu;
constexpr uint32_t k_done = 0xffffffffu;  ///< node reference: the query has finished
constexpr uint32_t k_pop = 0xfffffffeu;   ///< node reference: take the next entry off the stack

struct mesh_hit {
    float t;
    uint32_t prim;  ///< original triangle index, k_miss = none
    uint32_t slot;  ///< position in leaf order (where its float4s are), material in the top four bits
};

/// Per-lane traversal stack, all of it in shared memory: node references as
/// uint32 [entry][thread], entry distances as the upper 16 bits of the float
/// (truncation rounds a non-negative float down, which only makes the cull on
/// pop a little less eager) as uint16 [entry][thread]. 6 bytes per entry.
struct lane_stack {
    uint32_t* refs;
    uint16_t* dist;
    int sp = 0;

    __device__ __forceinline__ void push(uint2 e) { push(e.x, e.y); }
    /// `distance_bits`: the entry distance (>= 0) as the bits of its float.
    __device__ __forceinline__ void push(uint32_t ref, uint32_t distance_bits) {
        refs[sp * k_block_threads] = ref;
        dist[sp * k_block_threads] = static_cast<uint16_t>(distance_bits >> 16);
        sp++;
    }
    /// Stores the entry at the top and keeps it only if `keep`: needs one entry of room beyond
    /// what the stack logically holds (wide_bvh.hpp:collapse_wide is given that limit).
    __device__ __forceinline__ void push_if(bool keep, uint32_t ref, uint32_t distance_bits) {
        refs[sp * k_block_threads] = ref;
        dist[sp * k_block_threads] = static_cast<uint16_t>(distance_bits >> 16);
        sp += keep ? 1 : 0;
    }
    __device__ __forceinline__ uint32_t pop_one(float closest) {
        sp--;
        const uint32_t ref = refs[sp * k_block_threads];
        const float t_near = __uint_as_float(static_cast<uint32_t>(dist[sp * k_block_threads]) << 16);
        const uint32_t rest = sp > 0 ? k_pop : k_done;
        return t_near <= closest ? ref : rest;
    }
    __device__ __forceinline__ uint32_t after_node() const { return sp > 0 ? k_pop : k_done; }
};

using slab::box_hit;  // slab_test.hpp: the pruning arithmetic, shared with the host check

/// 1/x for the slab test (slab_test.hpp).
__device__ __forceinline__ float slab_reciprocal(float x) { return slab::reciprocal(x); }

/// Moeller-Trumbore with the operations, their order and the decisions of
/// oracle/quasi_oracle_mesh.c:qm_intersect, evaluated without branches: every
/// lane computes u, v and t and the early returns become one predicate. Values
/// computed past a failed test are never used (NaN and infinity compare false).
__device__ __forceinline__ void intersect(float4 v0m, float4 e1p, float4 e2_, v3 o, v3 d, mesh_hit& best) {
    const v3 e1 = make_v3(e1p.x, e1p.y, e1p.z), e2 = make_v3(e2_.x, e2_.y, e2_.z);
    const v3 pvec = cross(d, e2);
    const float det = dot(e1, pvec);
    const float inv_det = 1.0f / det;
    const v3 tvec = o - make_v3(v0m.x, v0m.y, v0m.z);
    const float u = dot(tvec, pvec) * inv_det;
    const v3 qvec = cross(tvec, e1);
    const float v = dot(d, qvec) * inv_det;
    const float t = dot(e2, qvec) * inv_det;
    const uint32_t prim = __float_as_uint(e1p.w);
    // the oracle's rejections, literally (so that an unordered compare decides the same way)
    bool reject = fabsf(det) < 1e-12f;
    reject = reject || u < 0.0f || u > 1.0f;
    reject = reject || v < 0.0f || u + v > 1.0f;
    reject = reject || t < 0.001f || t > best.t;
    reject = reject || (t == best.t && best.prim != k_miss && prim < best.prim);
    const bool ok = !reject;
    best.t = ok ? t : best.t;
    best.prim = ok ? prim : best.prim;
    best.slot = ok ? __float_as_uint(e2_.w) : best.slot;  // slot | material << k_slot_bits (mesh_scene.hpp)
}

/// Node references are {left_first << 3 | count}: count 0 = interior, 1..4 = leaf.
__device__ __forceinline__ uint32_t node_ref(float4 lo, float4 hi) {
    return (__float_as_uint(lo.w) << 3) | __float_as_uint(hi.w);
}

/// One interior step: tests both children (one 64-byte line), descends into the
/// nearer one that is hit and defers the other, or asks for the next stack entry.
__device__ __forceinline__ void interior_step(const float4* __restrict__ nodes, lane_stack& stack, v3 inv, v3 oi,
                                              float closest, uint32_t& cur) {
    const float4* c = nodes + 2u * static_cast<size_t>(cur >> 3);
    const float4 l_lo = __ldg(c), l_hi = __ldg(c + 1), r_lo = __ldg(c + 2), r_hi = __ldg(c + 3);
    float tl, tr;
    const bool hit_l = box_hit(l_lo, l_hi, inv, oi, closest, tl);
    const bool hit_r = box_hit(r_lo, r_hi, inv, oi, closest, tr);
    const uint32_t ref_l = node_ref(l_lo, l_hi), ref_r = node_ref(r_lo, r_hi);
    if (hit_l && hit_r) {
        const bool go_left = tl <= tr;
        stack.push(make_uint2(go_left ? ref_r : ref_l, __float_as_uint(go_left ? tr : tl)));
        cur = go_left ? ref_l : ref_r;
    } else if (hit_l || hit_r) {
        cur = hit_l ? ref_l : ref_r;
    } else {
        cur = stack.after_node();
    }
}

/// interior_step on the quantized pairs (quantized_bvh.hpp): both children are one 32-byte sector, two loads.
__device__ __forceinline__ void interior_step_quantized(const uint4* __restrict__ pairs, lane_stack& stack,
                                                        const slab::quant_ray& ray, float closest, uint32_t& cur) {
    const uint4* c = pairs + 2u * static_cast<size_t>(cur >> 4);  // cur >> 3 is the left child, 2p + 1
    const uint4 a = __ldg(c), b = __ldg(c + 1);
    float tl, tr;
    const bool hit_l = slab::box_hit_quantized(a.x, a.y, a.z, ray, closest, tl);
    const bool hit_r = slab::box_hit_quantized(a.w, b.x, b.y, ray, closest, tr);
    const uint32_t ref_l = b.z, ref_r = b.w;
    if (hit_l && hit_r) {
        const bool go_left = tl <= tr;
        stack.push(make_uint2(go_left ? ref_r : ref_l, __float_as_uint(go_left ? tr : tl)));
        cur = go_left ? ref_l : ref_r;
    } else if (hit_l || hit_r) {
        cur = hit_l ? ref_l : ref_r;
    } else {
        cur = stack.after_node();
    }
}

/// One visit of a 4-wide node (wide_bvh.hpp): one 128-byte line, four slab tests.
__device__ __forceinline__ void wide_step(const float4* __restrict__ nodes, lane_stack& stack, v3 inv, v3 oi,
                                          float closest, uint32_t& cur) {
    const float4* n = nodes + mesh::k_wide_node_float4 * static_cast<size_t>(cur >> 3);
    const float4 lx = __ldg(n), ly = __ldg(n + 1), lz = __ldg(n + 2), hx = __ldg(n + 3), hy = __ldg(n + 4),
                 hz = __ldg(n + 5), refs = __ldg(n + 6);
    cur = slab::wide_visit(lx, ly, lz, hx, hy, hz, refs, inv, oi, closest, stack);
}

}  // namespace
}  // namespace Q::cuda
