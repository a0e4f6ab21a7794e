/// @file mesh_trace.cu
/// @brief Path tracer over a triangle mesh behind a BVH (BASELINE config 3), sm_100a.
///
/// Same frame as pathtrace.cu (seed, camera ray, path loop, cosine sampling,
/// Russian roulette, AOVs, in-place accumulation: the reference shader,
/// /root/reference/backends/metal/plugin.mm:98-330), with trace_scene replaced by
/// a closest-hit query against one million triangles. The reference has no such
/// path; the arithmetic contract is oracle/quasi_oracle_mesh.h and the result is
/// bit-identical to it.
///
/// Bound: L2 / latency, not HBM and not FP32 issue. The
/// working set (20.6 MB of nodes + 48 MB of triangles) sits in the 126 MB L2, so
/// the kernel is built around memory-level parallelism and coherence:
///  - 32-byte nodes read as two float4; siblings are adjacent, so visiting an
///    interior node is one 64-byte line (4 x LDG.128, read-only path).
///  - Triangles in leaf order as three float4 arrays (SoA): 3 x LDG.128 each,
///    consecutive slots of a leaf are consecutive addresses.
///  - Per-lane traversal stack in shared memory, laid out [entry][thread] so a
///    push or pop by the whole warp touches 32 different banks. Entries carry
///    the entry distance, and a popped subtree that has fallen behind the
///    current closest hit is skipped without touching memory.
///  - Near child first, far child pushed (while-while traversal).
///  - Persistent warps with path regeneration (ballot + popc compaction), as in
///    pathtrace.cu: all lanes of a warp are always inside the same bounce.
/// The order in which nodes are visited cannot change the result: boxes are
/// padded (they only prune what the ray provably misses) and an equal t is won
/// by the higher triangle index.

#include "device_common.cuh"
#include "kernels.hpp"

#include <cuda_runtime.h>

namespace Q::cuda {

namespace {

using namespace dev;

constexpr int k_block_threads = 256;
constexpr int k_blocks_per_sm = 3;
constexpr int k_warps_per_block = k_block_threads / 32;
constexpr uint32_t k_miss = 0xffffffffu;
#ifndef QUASI_MESH_MIN_WAITING
#define QUASI_MESH_MIN_WAITING 8
#endif
constexpr int k_min_waiting = QUASI_MESH_MIN_WAITING;  ///< waiting lanes that trigger a service phase

struct mesh_hit {
    float t;
    uint32_t prim;  ///< original triangle index, k_miss = none
    uint32_t slot;  ///< position in leaf order (where its float4s are)
};

/// Slab test of oracle/quasi_oracle_mesh.c:box_hit; also returns the entry distance.
__device__ __forceinline__ bool box_hit(float4 lo, float4 hi, v3 o, v3 idir, float closest, float& t_near) {
    const float t0x = (lo.x - o.x) * idir.x, t1x = (hi.x - o.x) * idir.x;
    const float t0y = (lo.y - o.y) * idir.y, t1y = (hi.y - o.y) * idir.y;
    const float t0z = (lo.z - o.z) * idir.z, t1z = (hi.z - o.z) * idir.z;
    const float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
    const float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), closest));
    t_near = tn;
    return tn <= tf;
}

/// Moeller-Trumbore in the operation order of oracle/quasi_oracle_mesh.c:qm_intersect.
__device__ __forceinline__ void intersect(float4 v0m, float4 e1p, float4 e2_, uint32_t slot, v3 o, v3 d,
                                          mesh_hit& best) {
    const v3 e1 = make_v3(e1p.x, e1p.y, e1p.z), e2 = make_v3(e2_.x, e2_.y, e2_.z);
    const v3 pvec = cross(d, e2);
    const float det = dot(e1, pvec);
    if (fabsf(det) < 1e-12f) return;
    const float inv_det = 1.0f / det;
    const v3 tvec = o - make_v3(v0m.x, v0m.y, v0m.z);
    const float u = dot(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return;
    const v3 qvec = cross(tvec, e1);
    const float v = dot(d, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return;
    const float t = dot(e2, qvec) * inv_det;
    if (t < 0.001f || t > best.t) return;
    const uint32_t prim = __float_as_uint(e1p.w);
    if (t == best.t && best.prim != k_miss && prim < best.prim) return;
    best.t = t;
    best.prim = prim;
    best.slot = slot;
}

constexpr uint32_t k_done = 0xffffffffu;  ///< node reference: traversal finished

/// Node references are {left_first << 3 | count}: count 0 = interior, 1..4 = leaf.
__device__ __forceinline__ uint32_t node_ref(float4 lo, float4 hi) {
    return (__float_as_uint(lo.w) << 3) | __float_as_uint(hi.w);
}

/// Next deferred subtree that can still hold a closer (or tying) hit, or k_done.
__device__ __forceinline__ uint32_t pop(const uint2* stack, int& sp, float closest) {
    while (sp > 0) {
        sp--;
        const uint2 e = stack[sp * k_block_threads];
        if (__uint_as_float(e.y) <= closest) return e.x;
    }
    return k_done;
}

/// One interior step: tests both children (one 64-byte line), descends into the
/// nearer one that is hit and defers the other, or pops.
__device__ __forceinline__ void interior_step(const float4* __restrict__ nodes, uint2* stack, v3 o, v3 idir,
                                              float closest, uint32_t& cur, int& sp) {
    const float4* c = nodes + 2u * static_cast<size_t>(cur >> 3);
    const float4 l_lo = __ldg(c), l_hi = __ldg(c + 1), r_lo = __ldg(c + 2), r_hi = __ldg(c + 3);
    float tl, tr;
    const bool hit_l = box_hit(l_lo, l_hi, o, idir, closest, tl);
    const bool hit_r = box_hit(r_lo, r_hi, o, idir, closest, tr);
    const uint32_t ref_l = node_ref(l_lo, l_hi), ref_r = node_ref(r_lo, r_hi);
    if (hit_l && hit_r) {
        const bool go_left = tl <= tr;
        stack[sp * k_block_threads] = make_uint2(go_left ? ref_r : ref_l, __float_as_uint(go_left ? tr : tl));
        sp++;
        cur = go_left ? ref_l : ref_r;
    } else if (hit_l || hit_r) {
        cur = hit_l ? ref_l : ref_r;
    } else {
        cur = pop(stack, sp, closest);
    }
}

/// Persistent warps; every lane is a small state machine:
///   idle      no path (start, or the warp's pixels have run out)
///   tracing   `cur` != k_done: a closest-hit query is in flight (cursor, stack
///             pointer, best hit and ray live in registers, the stack in shared memory)
///   waiting   the query has finished; the hit has to be shaded
/// The warp alternates between two phases and reconverges between them:
///   service   waiting lanes shade their hit (next bounce, or write the pixel);
///             lanes without a path take the warp's next pixels (ballot + popc
///             compaction); every lane with a new ray tests the root and starts tracing.
///   traverse  warp-synchronous while-while (interior descent until every tracing
///             lane holds a leaf; leaf triangles; repeat) until k_min_waiting lanes
///             are waiting or nobody is tracing.
/// So a lane whose ray left the scene after 3 nodes does not sit out the 150
/// nodes of its neighbour: it is given the next ray as soon as enough lanes
/// are waiting to make the service phase worth issuing (Aila & Laine's
/// persistent while-while with dynamic fetch, with paths instead of rays).
__global__ void __launch_bounds__(k_block_threads, k_blocks_per_sm)
mesh_pathtrace_kernel(const __grid_constant__ mesh_params scene, const __grid_constant__ frame_params frame,
                      float4* __restrict__ beauty, float4* __restrict__ albedo, float4* __restrict__ normal,
                      float4* __restrict__ depth, unsigned long long* __restrict__ counters) {
    extern __shared__ uint2 s_stack[];  // [k_mesh_stack_entries][k_block_threads]
    __shared__ float s_material[k_mesh_materials * 8];
    for (int i = threadIdx.x; i < static_cast<int>(k_mesh_materials) * 8; i += k_block_threads) {
        s_material[i] = scene.materials[i];
    }
    __syncthreads();
    uint2* my_stack = s_stack + threadIdx.x;
    const float4* __restrict__ nodes = scene.nodes;

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lanes_below = (1u << lane) - 1u;
    const uint32_t warp_global = blockIdx.x * k_warps_per_block + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * k_warps_per_block;
    const uint32_t tiles_x = (frame.width + 7u) >> 3;
    const uint32_t n_tiles = tiles_x * ((frame.height + 3u) >> 2);
    const bool first_sample = frame.sample_index == 0;
    const float weight = frame.weight;
    const uint32_t max_bounces = scene.max_bounces;
    const v3 cam_origin = load_v3(frame.position);

    // path state
    bool active = false;
    v3 ro = cam_origin, rd = make_v3(0.0f, 0.0f, -1.0f);
    v3 throughput = make_v3(1.0f, 1.0f, 1.0f), color = make_v3(0.0f, 0.0f, 0.0f);
    uint32_t rng = 0, bounce = 0, pixel = 0;
    // query state
    uint32_t cur = k_done;
    int sp = 0;
    mesh_hit best{1e30f, k_miss, 0u};
    v3 idir = make_v3(0.0f, 0.0f, 0.0f);

    uint32_t rays = 0, nodes_visited = 0, tris_tested = 0;
    uint32_t next_item = 0;  // warp-uniform cursor into this warp's pixel sequence

    while (true) {
        // ---------------------------------------------------------- service ----
        bool new_ray = false;
        if (active && cur == k_done) {
            // One bounce of path_trace (plugin.mm:232-273) around the finished query.
            const bool is_hit = best.prim != k_miss;
            v3 n = make_v3(0.0f, 0.0f, 0.0f), q_emission = n, q_albedo = n;
            bool is_light = false;
            if (is_hit) {
                const float4 v0m = __ldg(scene.tri_v0 + best.slot);
                const float4 e1 = __ldg(scene.tri_e1 + best.slot), e2 = __ldg(scene.tri_e2 + best.slot);
                const float* m = s_material + __float_as_uint(v0m.w) * 8u;
                q_albedo = load_v3(m);
                q_emission = load_v3(m + 4);
                is_light = fmaxf(q_emission.x, fmaxf(q_emission.y, q_emission.z)) > 0.1f;
                const v3 ng = normalize(cross(make_v3(e1.x, e1.y, e1.z), make_v3(e2.x, e2.y, e2.z)));
                n = dot(ng, rd) < 0.0f ? ng : make_v3(-ng.x, -ng.y, -ng.z);
            }

            if (bounce == 0) {
                v3 a = make_v3(0.0f, 0.0f, 0.0f), nn = a;
                float first_t = 0.0f;
                if (is_hit) {
                    a = is_light ? q_emission : q_albedo;
                    nn = make_v3(n.x * 0.5f + 0.5f, n.y * 0.5f + 0.5f, n.z * 0.5f + 0.5f);
                    first_t = best.t;
                }
                const float nd = fminf(fmaxf((first_t - 1.0f) / 4.0f, 0.0f), 1.0f);
                accumulate_pixel(albedo, pixel, make_float4(a.x, a.y, a.z, 1.0f), first_sample, weight);
                accumulate_pixel(normal, pixel, make_float4(nn.x, nn.y, nn.z, 1.0f), first_sample, weight);
                accumulate_pixel(depth, pixel, make_float4(nd, nd, nd, 1.0f), first_sample, weight);
            }

            bool done = !is_hit;
            if (is_hit) {
                color = color + throughput * q_emission;
                done = is_light;
                if (!done) {
                    throughput = throughput * q_albedo;
                    if (bounce > 2) {
                        const float p = fmaxf(0.05f, fmaxf(throughput.x, fmaxf(throughput.y, throughput.z)));
                        if (random_float(rng) > p) {
                            done = true;
                        } else {
                            throughput = div3(throughput, p);
                        }
                    }
                }
            }
            bounce++;
            done = done || bounce >= max_bounces;

            if (!done) {
                const v3 point = ro + rd * best.t;
                ro = point + n * 0.001f;
                // cosine_sample_hemisphere, plugin.mm:111-125, literally
                const float r1 = random_float(rng);
                const float r2 = random_float(rng);
                const float phi = 6.28318548202514648f * r1;
                const float cos_theta = sqrtf(1.0f - r2);
                const float sin_theta = sqrtf(r2);
                float sin_phi, cos_phi;
                sincos_round_once(phi, sin_phi, cos_phi);
                const v3 w = normalize(n);
                const v3 a = fabsf(w.x) > 0.9f ? make_v3(0.0f, 1.0f, 0.0f) : make_v3(1.0f, 0.0f, 0.0f);
                const v3 v = normalize(cross(w, a));
                const v3 u = cross(w, v);
                const v3 dir = (u * cos_phi) * sin_theta + (v * sin_phi) * sin_theta + w * cos_theta;
                rd = normalize(dir);
                new_ray = true;
            } else {
                accumulate_pixel(beauty, pixel, make_float4(color.x, color.y, color.z, 1.0f), first_sample,
                                 weight);
                active = false;
            }
        }
        __syncwarp();

        // Regeneration: lanes without a path take the warp's next pixels (see pathtrace.cu).
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u && warp_global + (next_item >> 5) * n_warps < n_tiles) {
            if (!active) {
                const uint32_t item = next_item + __popc(idle & lanes_below);
                const uint32_t tile = warp_global + (item >> 5) * n_warps;
                if (tile < n_tiles) {
                    uint32_t tile_y = __umulhi(tile, frame.tiles_x_magic);
                    uint32_t tile_x = tile - tile_y * tiles_x;
                    if (tile_x >= tiles_x) {
                        tile_x -= tiles_x;
                        tile_y++;
                    }
                    const uint32_t px = tile_x * 8u + (item & 7u);
                    const uint32_t py = tile_y * 4u + ((item >> 3) & 3u);
                    if (px < frame.width && py < frame.height) {
                        float uv_x = (static_cast<float>(px) + 0.5f) / static_cast<float>(frame.width);
                        float uv_y = 1.0f - (static_cast<float>(py) + 0.5f) / static_cast<float>(frame.height);
                        rng = pcg_hash(px + py * 1920u + frame.frame_index * 1920u * 1080u);
                        const float jx = random_float(rng);
                        const float jy = random_float(rng);
                        uv_x += (jx - 0.5f) * 0.001f;
                        uv_y += (jy - 0.5f) * 0.001f;
                        const v3 target = load_v3(frame.lower_left) + load_v3(frame.horizontal) * uv_x +
                                          load_v3(frame.vertical) * uv_y;
                        ro = cam_origin;
                        rd = normalize(target - cam_origin);
                        throughput = make_v3(1.0f, 1.0f, 1.0f);
                        color = make_v3(0.0f, 0.0f, 0.0f);
                        bounce = 0;
                        pixel = py * frame.width + px;
                        active = true;
                        new_ray = true;
                        if (!first_sample) {
                            prefetch_l2(beauty + pixel);
                            prefetch_l2(albedo + pixel);
                            prefetch_l2(normal + pixel);
                            prefetch_l2(depth + pixel);
                        }
                    }
                }
            }
            next_item += __popc(idle);
        }

        // Every lane with a new ray: root test, start of the query.
        if (new_ray) {
            rays++;
            best = mesh_hit{1e30f, k_miss, 0u};
            idir = make_v3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
            sp = 0;
            float tn;
            const float4 lo = __ldg(nodes), hi = __ldg(nodes + 1);
            nodes_visited++;
            cur = box_hit(lo, hi, ro, idir, best.t, tn) ? node_ref(lo, hi) : k_done;
        }
        if (!__any_sync(0xffffffffu, active)) {
            if (warp_global + (next_item >> 5) * n_warps >= n_tiles) break;  // no paths, no pixels left
            continue;  // only off-image pixels of an edge tile were handed out
        }

        // --------------------------------------------------------- traverse ----
        while (__any_sync(0xffffffffu, cur != k_done)) {
            while ((cur & 7u) == 0u) {
                nodes_visited += 2u;
                interior_step(nodes, my_stack, ro, idir, best.t, cur, sp);
            }
            __syncwarp();
            if (cur != k_done) {
                const uint32_t first = cur >> 3, count = cur & 7u;
                for (uint32_t k = 0; k < count; k++) {
                    const uint32_t slot = first + k;
                    intersect(__ldg(scene.tri_v0 + slot), __ldg(scene.tri_e1 + slot), __ldg(scene.tri_e2 + slot),
                              slot, ro, rd, best);
                }
                tris_tested += count;
                cur = pop(my_stack, sp, best.t);
            }
            const uint32_t waiting = __ballot_sync(0xffffffffu, active && cur == k_done);
            if (__popc(waiting) >= k_min_waiting) break;
        }
    }

    rays = __reduce_add_sync(0xffffffffu, rays);
    nodes_visited = __reduce_add_sync(0xffffffffu, nodes_visited);
    tris_tested = __reduce_add_sync(0xffffffffu, tris_tested);
    if (lane == 0 && rays != 0) {
        atomicAdd(counters + 0, static_cast<unsigned long long>(rays));
        atomicAdd(counters + 1, static_cast<unsigned long long>(nodes_visited));
        atomicAdd(counters + 2, static_cast<unsigned long long>(tris_tested));
    }
}

}  // namespace

cudaError_t launch_mesh_pathtrace(const mesh_params& scene, const frame_params& frame, float* const layers[4],
                                  unsigned long long* counters, int sm_count, cudaStream_t stream) {
    static const cudaError_t attr = cudaFuncSetAttribute(
        mesh_pathtrace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
        static_cast<int>(k_mesh_stack_entries * k_block_threads * sizeof(uint2)));
    if (attr != cudaSuccess) return attr;
    const uint32_t n_tiles = ((frame.width + 7u) / 8u) * ((frame.height + 3u) / 4u);
    uint32_t blocks = static_cast<uint32_t>(sm_count) * k_blocks_per_sm;
    const uint32_t needed = (n_tiles + k_warps_per_block - 1u) / k_warps_per_block;
    if (needed < blocks) blocks = needed;
    if (blocks == 0) return cudaSuccess;
    const size_t smem = static_cast<size_t>(k_mesh_stack_entries) * k_block_threads * sizeof(uint2);
    mesh_pathtrace_kernel<<<blocks, k_block_threads, smem, stream>>>(
        scene, frame, reinterpret_cast<float4*>(layers[0]), reinterpret_cast<float4*>(layers[1]),
        reinterpret_cast<float4*>(layers[2]), reinterpret_cast<float4*>(layers[3]), counters);
    return cudaGetLastError();
}

}  // namespace Q::cuda
