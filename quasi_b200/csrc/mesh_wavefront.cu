/// @file mesh_wavefront.cu
/// @brief The mesh scene (BASELINE config 3) as a wavefront: rays in queues in HBM, one kind of work per kernel.
///
/// mesh_trace.cu keeps a path on its lane from the camera to its end, so a warp has to stop walking the tree
/// whenever enough of its lanes wait to be shaded: shading (sin/cos, three branches, the AOV writes) is only worth
/// issuing for 16 lanes at once, and until then those lanes idle. ncu and the emulator's pass counts agree on
/// what that costs: descent passes at 19 of 32 lanes, shading at 7 to 17. Here the same
/// arithmetic is cut at the closest-hit query:
///   mesh_wave_raygen_kernel   every (pixel, sample) of the batch: seed, jitter, camera ray (plugin.mm:129-153,
///                             288-298) into queue 0, tile by tile (a warp's 32 rays are one 8x4 tile);
///   mesh_wave_trace_kernel    persistent; a lane whose query has finished stores {t, triangle, slot} where its ray
///                             was and takes the next unfetched ray of the queue, a handful of instructions, so it
///                             is done as soon as four lanes want it and the warp goes back to the tree with
///                             (nearly) every lane. Same stack, steps, triangle test and vote-driven schedule as
///                             mesh_trace.cu (mesh_traverse.cuh);
///   mesh_wave_shade_kernel    one bounce of path_trace (plugin.mm:232-273) around the hits of a wave, 32 lanes of
///                             the same work, coalesced: first-hit AOVs, emission, roulette, the next direction
///                             (cosine_sample_hemisphere, plugin.mm:111-125); continuing paths are compacted into
///                             the other queue (one atomic per warp), finished ones store their colour;
///   mesh_wave_resolve_kernel  accumulate (plugin.mm:313-330): the k samples of a pixel folded into the four layers
///                             in sample order, layers read and written once per batch instead of once per sample.
/// All rays of wave b are at bounce b, so a frame of k samples is 2 * max_bounces + 3 launches however large k is:
/// the plugin collects the headless render calls of a frame (QUASI_CUDA_MESH_BATCH, default 16) as it does for the
/// demo scene. Every draw of a path's RNG stream happens in the order of the shader, every sample is folded with
/// the weight of its own call: the image is the oracle's bit for bit, as with the one-kernel form (which stays for
/// calls with a drawable, one sample at a time).
/// Memory: 184 bytes per path of the batch (two queues of 64, the hit, three result float4), 6.1 GB for 16 samples
/// of 1080p; 0.19 KB of streamed traffic per ray next to the 1.3 KB of nodes and triangles it reads from L1/L2.

#include "mesh_traverse.cuh"

#include <cuda_runtime.h>

#if defined(QUASI_HOST_EMULATION) && defined(QUASI_EMU_STATS)
#define QUASI_WAVE_PASS(phase, lane_works) emu::phase_pass(phase, __ballot_sync(0xffffffffu, lane_works));
#else
#define QUASI_WAVE_PASS(phase, lane_works)
#endif

namespace Q::cuda {

namespace {

#ifndef QUASI_WAVE_REFILL
#define QUASI_WAVE_REFILL 8
#endif
#ifndef QUASI_WAVE_LEAF_BATCH
#define QUASI_WAVE_LEAF_BATCH 8
#endif
#ifndef QUASI_WAVE_RUN_EXIT
#define QUASI_WAVE_RUN_EXIT 12
#endif
#ifndef QUASI_WAVE_CHUNK
#define QUASI_WAVE_CHUNK 128
#endif

constexpr int k_wave_blocks_per_sm = 4;
constexpr int k_wave_warps_per_block = k_block_threads / 32;
constexpr int k_refill = QUASI_WAVE_REFILL;          ///< finished lanes that send the warp to the queue
constexpr int k_wave_leaf_batch = QUASI_WAVE_LEAF_BATCH;
constexpr int k_wave_run_exit = QUASI_WAVE_RUN_EXIT;
constexpr uint32_t k_chunk = QUASI_WAVE_CHUNK;       ///< rays a warp reserves per atomic on the queue head

__device__ __forceinline__ float4 load_once(const float4* p) {
#if defined(QUASI_HOST_EMULATION)
    return *p;
#else
    return __ldcs(p);
#endif
}
__device__ __forceinline__ void store_hit(float2* p, float t, uint32_t slot) {
#if defined(QUASI_HOST_EMULATION)
    *p = make_float2(t, __uint_as_float(slot));
#else
    __stcs(p, make_float2(t, __uint_as_float(slot)));
#endif
}
__device__ __forceinline__ float2 load_hit(const float2* p) {
#if defined(QUASI_HOST_EMULATION)
    return *p;
#else
    return __ldcs(p);
#endif
}
__device__ __forceinline__ void store_once(float4* p, float4 v) {
#if defined(QUASI_HOST_EMULATION)
    *p = v;
#else
    __stcs(p, v);
#endif
}

constexpr uint32_t k_dead_path = 0xffffffffu;   ///< a queue entry that is no ray: an off-image lane of an edge tile
constexpr uint32_t k_has_colour = 0x80000000u;  ///< in a path word: the path's colour so far is not zero (it is stored)

/// Queue 0: the camera rays of `samples` samples of every pixel, sample-major, a warp's 32 rays one 8x4 tile, at
/// position (sample * tiles + tile) * 32 + lane: no counter to contend for. Lanes of an edge tile that fall outside
/// the image leave an entry marked dead, which is neither traced nor counted.
__global__ void __launch_bounds__(k_block_threads)
mesh_wave_raygen_kernel(const __grid_constant__ frame_params frame, const __grid_constant__ mesh_wave_buffers q,
                        uint32_t samples) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tiles_x = (frame.width + 7u) >> 3;
    const uint32_t n_tiles = tiles_x * ((frame.height + 3u) >> 2);
    const uint32_t n_px = frame.width * frame.height;
    const uint32_t warps = gridDim.x * k_wave_warps_per_block;
    const uint64_t items = static_cast<uint64_t>(n_tiles) * samples;
    const v3 cam_origin = load_v3(frame.position);
    if (blockIdx.x == 0 && threadIdx.x == 0) q.counts[0] = static_cast<uint32_t>(items * 32u);
    for (uint64_t item = static_cast<uint64_t>(blockIdx.x) * k_wave_warps_per_block + (threadIdx.x >> 5); item < items;
         item += warps) {
        const uint32_t sample = static_cast<uint32_t>(item / n_tiles);
        const uint32_t tile = static_cast<uint32_t>(item - static_cast<uint64_t>(sample) * n_tiles);
        const uint32_t tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
        const uint32_t px = tile_x * 8u + (lane & 7u), py = tile_y * 4u + (lane >> 3);
        const size_t at = static_cast<size_t>(item) * 32u + lane;
        if (px < frame.width && py < frame.height) {
            // pathtrace_frag, plugin.mm:288-298, and get_camera_ray, plugin.mm:129-153 (jitter x before y)
            float uv_x = (static_cast<float>(px) + 0.5f) / static_cast<float>(frame.width);
            float uv_y = 1.0f - (static_cast<float>(py) + 0.5f) / static_cast<float>(frame.height);
            uint32_t rng = pcg_hash(px + py * 1920u + (frame.frame_index + sample * frame.frame_stride) * 1920u * 1080u);
            const float draw_a = random_float(rng);
            const float draw_b = random_float(rng);
            uv_x += (draw_a - 0.5f) * 0.001f;
            uv_y += (draw_b - 0.5f) * 0.001f;
            const v3 target = load_v3(frame.lower_left) + load_v3(frame.horizontal) * uv_x + load_v3(frame.vertical) * uv_y;
            const v3 rd = normalize(target - cam_origin);
            const uint32_t path = sample * n_px + py * frame.width + px;
            store_once(q.origin[0] + at, make_float4(cam_origin.x, cam_origin.y, cam_origin.z, __uint_as_float(path)));
            store_once(q.direction[0] + at, make_float4(rd.x, rd.y, rd.z, __uint_as_float(rng)));
        } else {
            store_once(q.origin[0] + at, make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(k_dead_path)));
        }
    }
}

/// Closest hits of the rays [0, *count) of one queue: hit[i] = {t, slot in leaf order as bits (all ones: no hit)}.
/// `Quant`: the binary tree is walked in its quantized form (scene.pairs, quantized_bvh.hpp); `inv` and `oi` then hold
/// the ray's constants of that slab test.
template <bool Wide, bool Quant>
__global__ void __launch_bounds__(k_block_threads, k_wave_blocks_per_sm)
mesh_wave_trace_kernel(const __grid_constant__ mesh_params scene, const float4* __restrict__ origin,
                       const float4* __restrict__ direction, float2* __restrict__ hit,
                       const uint32_t* __restrict__ count_ptr, uint32_t* __restrict__ head,
                       unsigned long long* __restrict__ counters) {
#if defined(QUASI_HOST_EMULATION)  // tests/cuda_emulation: the dynamic shared memory of this CTA
    uint32_t* s_refs = static_cast<uint32_t*>(emu::dynamic_shared());
#else
    extern __shared__ uint32_t s_refs[];  // [k_mesh_stack_entries][k_block_threads], then the uint16 distances
#endif
    lane_stack stack;
    stack.refs = s_refs + threadIdx.x;
    stack.dist = reinterpret_cast<uint16_t*>(s_refs + k_mesh_stack_entries * k_block_threads) + threadIdx.x;
    const float4* __restrict__ nodes = scene.nodes;
    const uint32_t count = *count_ptr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lanes_below = (1u << lane) - 1u;

    bool has_ray = false;
    uint32_t index = 0;        // of this lane's ray in the queue
    uint32_t cur = k_done;
    mesh_hit best{1e30f, k_miss, 0u};
    v3 ro = make_v3(0.0f, 0.0f, 0.0f), rd = make_v3(0.0f, 0.0f, -1.0f), inv = ro, oi = ro;
    [[maybe_unused]] uint32_t near_x = 0, near_y = 0, near_z = 0;  // Quant: quant_ray's selectors
    uint32_t chunk_next = 0, chunk_end = 0;  // warp-uniform: the rays this warp has reserved and not yet handed out
    bool dry = false;                        // warp-uniform: the queue head has passed the last ray
    uint32_t rays = 0, nodes_visited = 0, tris_tested = 0;

    while (true) {
        // ------------------------------------------------------------ refill ----
        const bool wants = cur == k_done;
        const uint32_t m_wants = __ballot_sync(0xffffffffu, wants);
        QUASI_WAVE_PASS(0, wants && (has_ray || !dry || chunk_next != chunk_end))
        if (m_wants != 0u) {
            if (has_ray && wants) {
                store_hit(hit + index, best.t, best.prim != k_miss ? best.slot : k_miss);
                has_ray = false;
            }
            const uint32_t need = static_cast<uint32_t>(__popc(m_wants));
            const uint32_t left = chunk_end - chunk_next;  // of the chunk in hand
            uint32_t fresh = 0, fresh_end = 0;              // a newly reserved chunk, if the one in hand is too short
            if (left < need && !dry) {
                if (lane == 0u) fresh = atomicAdd(head, k_chunk);
                fresh = __shfl_sync(0xffffffffu, fresh, 0);
                fresh_end = fresh + k_chunk < count ? fresh + k_chunk : count;
                if (fresh >= count) {
                    dry = true;
                    fresh = fresh_end = 0;
                }
            }
            const uint32_t rank = static_cast<uint32_t>(__popc(m_wants & lanes_below));
            uint32_t take = 0xffffffffu;
            if (wants) {
                if (rank < left) {
                    take = chunk_next + rank;
                } else if (rank - left < fresh_end - fresh) {
                    take = fresh + (rank - left);
                }
            }
            if (need <= left) {
                chunk_next += need;
            } else {
                const uint32_t from_fresh = need - left < fresh_end - fresh ? need - left : fresh_end - fresh;
                chunk_next = fresh + from_fresh;
                chunk_end = fresh_end;
            }
            float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (take != 0xffffffffu) {
                o = load_once(origin + take);
                if (__float_as_uint(o.w) == k_dead_path) take = 0xffffffffu;  // no ray here: the lane stays without one
            }
            if (take != 0xffffffffu) {
                const float4 d = load_once(direction + take);
                index = take;
                has_ray = true;
                ro = make_v3(o.x, o.y, o.z);
                rd = make_v3(d.x, d.y, d.z);
                rays++;
                best = mesh_hit{1e30f, k_miss, 0u};
                inv = make_v3(slab::reciprocal(rd.x), slab::reciprocal(rd.y), slab::reciprocal(rd.z));
                oi = ro * inv;
                stack.sp = 0;
                if constexpr (Wide) {
                    cur = 0u;  // the root of the wide tree; its children's boxes are the first test
                } else {
                    float tn;
                    const float4 lo = __ldg(nodes), hi = __ldg(nodes + 1);
                    nodes_visited++;
                    cur = box_hit(lo, hi, inv, oi, best.t, tn) ? node_ref(lo, hi) : k_done;
                    if constexpr (Quant) {
                        const slab::quant_ray r = slab::make_quant_ray(scene.grid_origin, scene.grid_step, inv, oi);
                        inv = r.a;
                        oi = r.b;
                        near_x = r.near_x, near_y = r.near_y, near_z = r.near_z;
                    }
                }
            }
        }
        const bool source_empty = dry && chunk_next == chunk_end;  // warp-uniform
        if (__ballot_sync(0xffffffffu, has_ray) == 0u) {
            if (source_empty) break;
            continue;
        }

        // ---------------------------------------------------------- traverse ----
        while (true) {
            // low 3 bits of cur: 0 interior, 1..4 leaf with that many triangles, 6 k_pop, 7 k_done
            const bool descending = (cur & 7u) == 0u || cur == k_pop;
            const uint32_t m_descending = __ballot_sync(0xffffffffu, descending);
            const uint32_t m_leaf = __ballot_sync(0xffffffffu, (cur & 7u) - 1u < 4u);
            const uint32_t m_tracing = m_descending | m_leaf;
            // nothing to walk, or enough lanes without a query while the queue still has rays for them
            if (m_tracing == 0u || (!source_empty && 32 - __popc(m_tracing) >= k_refill)) break;
            QUASI_WAVE_PASS(2, m_tracing != 0u)
            if (m_descending != 0u && __popc(m_leaf) < k_wave_leaf_batch) {
                const int keep = max(__popc(m_descending) - k_wave_run_exit, 0);
                do {
                    QUASI_WAVE_PASS(3, (cur & 7u) == 0u || cur == k_pop)
                    if (cur == k_pop) cur = stack.pop_one(best.t);
                    if ((cur & 7u) == 0u) {
                        if constexpr (Wide) {
                            nodes_visited += mesh::k_wide;
                            wide_step(nodes, stack, inv, oi, best.t, cur);
                        } else if constexpr (Quant) {
                            nodes_visited += 2u;
                            interior_step_quantized(scene.pairs, stack, slab::quant_ray{inv, oi, near_x, near_y, near_z}, best.t, cur);
                        } else {
                            nodes_visited += 2u;
                            interior_step(nodes, stack, inv, oi, best.t, cur);
                        }
                    }
                } while (__popc(__ballot_sync(0xffffffffu, (cur & 7u) == 0u || cur == k_pop)) > keep);
            } else {
#if defined(QUASI_HOST_EMULATION) && defined(QUASI_EMU_STATS)
                for (uint32_t k = 0; k < 4u; k++) QUASI_WAVE_PASS(4, (cur & 7u) - 1u < 4u && (cur & 7u) > k)
#endif
                if ((cur & 7u) - 1u < 4u) {
                    const uint32_t first = cur >> 3, n = cur & 7u;
                    for (uint32_t k = 0; k < n; k++) {
                        const uint32_t slot = first + k;
                        intersect(__ldg(scene.tri_v0 + slot), __ldg(scene.tri_e1 + slot), __ldg(scene.tri_e2 + slot), ro, rd, best);
                    }
                    tris_tested += n;
                    cur = stack.after_node();
                }
            }
        }
    }

    rays = __reduce_add_sync(0xffffffffu, rays);
    nodes_visited = __reduce_add_sync(0xffffffffu, nodes_visited);
    tris_tested = __reduce_add_sync(0xffffffffu, tris_tested);
    if (lane == 0u && rays != 0u) {
        atomicAdd(counters + 0, static_cast<unsigned long long>(rays));
        atomicAdd(counters + 1, static_cast<unsigned long long>(nodes_visited));
        atomicAdd(counters + 2, static_cast<unsigned long long>(tris_tested));
    }
}

/// One bounce of path_trace (plugin.mm:232-273) around the hits of wave `wave` (all of its rays are at that
/// bounce): queue wave & 1 and its hits in, continuing paths into the other queue, results by path.
/// The queue positions of the continuing paths come from ONE atomic per CTA and pass (the lanes' votes, the warps'
/// counts through shared memory): a warp-level atomic per 32 rays, a million of them on one address in wave 0,
/// is what the kernel would otherwise wait for.
__global__ void __launch_bounds__(k_block_threads)
mesh_wave_shade_kernel(const __grid_constant__ mesh_params scene, const __grid_constant__ mesh_wave_buffers q,
                       uint32_t wave) {
    __shared__ float s_material[k_mesh_materials * 8];
    __shared__ uint32_t s_warp_count[k_wave_warps_per_block];
    __shared__ uint32_t s_base;
    for (int i = threadIdx.x; i < static_cast<int>(k_mesh_materials) * 8; i += k_block_threads) {
        s_material[i] = scene.materials[i];
    }
    __syncthreads();
    const uint32_t in = wave & 1u, out = in ^ 1u;
    const uint32_t count = q.counts[wave];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lanes_below = (1u << lane) - 1u;
    const uint32_t stride = gridDim.x * k_block_threads;
    const uint32_t max_bounces = scene.max_bounces;
    // whole CTAs stay in the loop: the compaction below is a CTA collective
    for (uint32_t first = blockIdx.x * k_block_threads; first < count; first += stride) {
        const uint32_t i = first + threadIdx.x;
        bool goes_on = false;
        v3 ro = make_v3(0.0f, 0.0f, 0.0f), dir = make_v3(0.0f, 0.0f, -1.0f), throughput = make_v3(1.0f, 1.0f, 1.0f),
           color = make_v3(0.0f, 0.0f, 0.0f);
        uint32_t rng = 0, path = k_dead_path;
        float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (i < count) {
            o = load_once(q.origin[in] + i);
            path = __float_as_uint(o.w);
        }
        if (path != k_dead_path) {
            const float4 d = load_once(q.direction[in] + i);
            const float2 h = load_hit(q.hit + i);
            const v3 rd = make_v3(d.x, d.y, d.z);
            ro = make_v3(o.x, o.y, o.z);
            rng = __float_as_uint(d.w);
            if (wave != 0u) {
                const float4 t = load_once(q.throughput[in] + i);
                throughput = make_v3(t.x, t.y, t.z);
                if ((path & k_has_colour) != 0u) {
                    const float4 c = load_once(q.colour[in] + i);
                    color = make_v3(c.x, c.y, c.z);
                }
                path &= ~k_has_colour;
            }
            const float best_t = h.x;
            const uint32_t recorded = __float_as_uint(h.y);  // slot | material << k_slot_bits, or all ones
            const uint32_t slot = recorded & ((1u << mesh::k_slot_bits) - 1u);

            const bool is_hit = recorded != k_miss;
            v3 n = make_v3(0.0f, 0.0f, 0.0f), q_emission = n, q_albedo = n, fu = n, fv = n, fw = n;
            bool is_light = false;
            if (is_hit) {
                const float4* rec = scene.tri_shade + 3u * static_cast<size_t>(slot);
                const float4 s0 = __ldg(rec), s1 = __ldg(rec + 1), s2 = __ldg(rec + 2);
                const float* m = s_material + (recorded >> mesh::k_slot_bits) * 8u;
                q_albedo = load_v3(m);
                q_emission = load_v3(m + 4);
                is_light = fmaxf(q_emission.x, fmaxf(q_emission.y, q_emission.z)) > 0.1f;
                // normal faces the ray; the frame of -ng is (u, -v, -w), exactly
                const v3 ng = make_v3(s0.x, s0.y, s0.z);
                const float side = dot(ng, rd) < 0.0f ? 1.0f : -1.0f;
                n = ng * side;
                fw = make_v3(s0.w, s1.x, s1.y) * side;
                fv = make_v3(s1.z, s1.w, s2.x) * side;
                fu = make_v3(s2.y, s2.z, s2.w);
            }

            if (wave == 0u) {  // first-hit AOVs (plugin.mm:241-249, 300-307)
                v3 a = make_v3(0.0f, 0.0f, 0.0f), nn = a;
                float first_t = 0.0f;
                if (is_hit) {
                    a = is_light ? q_emission : q_albedo;
                    nn = make_v3(n.x * 0.5f + 0.5f, n.y * 0.5f + 0.5f, n.z * 0.5f + 0.5f);
                    first_t = best_t;
                }
                const float nd = fminf(fmaxf((first_t - 1.0f) / 4.0f, 0.0f), 1.0f);
                store_once(q.out_albedo_depth + path, make_float4(a.x, a.y, a.z, nd));
                store_once(q.out_normal + path, make_float4(nn.x, nn.y, nn.z, 1.0f));
            }

            bool done = !is_hit;
            if (is_hit) {
                color = color + throughput * q_emission;
                done = is_light;
                if (!done) {
                    throughput = throughput * q_albedo;
                    if (wave > 2u) {
                        const float p = fmaxf(0.05f, fmaxf(throughput.x, fmaxf(throughput.y, throughput.z)));
                        if (random_float(rng) > p) {
                            done = true;
                        } else {
                            throughput = div3(throughput, p);
                        }
                    }
                }
            }
            done = done || wave + 1u >= max_bounces;
            if (!done) {
                const v3 point = ro + rd * best_t;
                ro = point + n * 0.001f;
                // cosine_sample_hemisphere, plugin.mm:111-125 (r1 before r2), with the triangle's frame
                const float draw_a = random_float(rng);
                const float draw_b = random_float(rng);
                const float phi = 6.28318548202514648f * draw_a;
                const float cos_theta = sqrtf(1.0f - draw_b);
                const float sin_theta = sqrtf(draw_b);
                float sin_phi, cos_phi;
                sincos_round_once(phi, sin_phi, cos_phi);
                dir = normalize((fu * cos_phi) * sin_theta + (fv * sin_phi) * sin_theta + fw * cos_theta);
                goes_on = true;
            } else {
                store_once(q.out_beauty + path, make_float4(color.x, color.y, color.z, 1.0f));
            }
        }
        // positions in the next queue: lane within warp, warp within CTA, CTA by one atomic
        const uint32_t m_goes_on = __ballot_sync(0xffffffffu, goes_on);
        if (lane == 0u) s_warp_count[warp] = static_cast<uint32_t>(__popc(m_goes_on));
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (uint32_t w = 0; w < static_cast<uint32_t>(k_wave_warps_per_block); w++) {
            const uint32_t c = s_warp_count[w];
            before += w < warp ? c : 0u;
            total += c;
        }
        if (threadIdx.x == 0 && total != 0u) s_base = atomicAdd(q.counts + wave + 1u, total);
        __syncthreads();
        if (goes_on) {
            const uint32_t at = s_base + before + static_cast<uint32_t>(__popc(m_goes_on & lanes_below));
            const bool has_colour = (__float_as_uint(color.x) | __float_as_uint(color.y) | __float_as_uint(color.z)) != 0u;
            store_once(q.origin[out] + at, make_float4(ro.x, ro.y, ro.z, __uint_as_float(path | (has_colour ? k_has_colour : 0u))));
            store_once(q.direction[out] + at, make_float4(dir.x, dir.y, dir.z, __uint_as_float(rng)));
            store_once(q.throughput[out] + at, make_float4(throughput.x, throughput.y, throughput.z, 0.0f));
            if (has_colour) store_once(q.colour[out] + at, make_float4(color.x, color.y, color.z, 0.0f));
        }
    }
}

/// accumulate (plugin.mm:313-330) for the `samples` samples of the batch, in sample order, in place. `Beauty`: the
/// beauty layer, when every path has ended; otherwise the three first-hit layers (albedo, normal, depth), which are
/// complete after the first shade pass, so that they can leave for the host while the later waves are traced.
template <bool Beauty>
__global__ void __launch_bounds__(k_block_threads)
mesh_wave_resolve_kernel(const __grid_constant__ frame_params frame, const __grid_constant__ mesh_wave_buffers q,
                         uint32_t samples, float4* __restrict__ beauty, float4* __restrict__ albedo,
                         float4* __restrict__ normal, float4* __restrict__ depth) {
    const uint32_t n_px = frame.width * frame.height;
    const uint32_t stride = gridDim.x * k_block_threads;
    for (uint32_t pixel = blockIdx.x * k_block_threads + threadIdx.x; pixel < n_px; pixel += stride) {
        const bool fresh = frame.sample_index == 0u;  // history before the first sample is defined as zero
        float4 b = make_float4(0.0f, 0.0f, 0.0f, 0.0f), a = b, n = b, d = b;
        if (!fresh) {
            if constexpr (Beauty) {
                b = load_once(beauty + pixel);
            } else {
                a = load_once(albedo + pixel);
                n = load_once(normal + pixel);
                d = load_once(depth + pixel);
            }
        }
        for (uint32_t s = 0; s < samples; s++) {
            const size_t path = static_cast<size_t>(s) * n_px + pixel;
            const uint32_t sample_index = frame.sample_index + s;
            const float weight = 1.0f / static_cast<float>(sample_index + 1u);
            if constexpr (Beauty) {
                const float4 sb = load_once(q.out_beauty + path);
                b = sample_index == 0u ? sb : mix4(b, sb, weight);
            } else {
                const float4 sa = load_once(q.out_albedo_depth + path), sn = load_once(q.out_normal + path);
                const float4 va = make_float4(sa.x, sa.y, sa.z, 1.0f), vd = make_float4(sa.w, sa.w, sa.w, 1.0f);
                if (sample_index == 0u) {
                    a = va, n = sn, d = vd;
                } else {
                    a = mix4(a, va, weight);
                    n = mix4(n, sn, weight);
                    d = mix4(d, vd, weight);
                }
            }
        }
        if constexpr (Beauty) {
            store_once(beauty + pixel, b);
        } else {
            store_once(albedo + pixel, a);
            store_once(normal + pixel, n);
            store_once(depth + pixel, d);
        }
    }
}

}  // namespace

#if !defined(QUASI_HOST_EMULATION)  // tests/cuda_emulation has its own launcher

size_t mesh_wave_bytes_per_path() { return 2 * 4 * sizeof(float4) + sizeof(float2) + 3 * sizeof(float4); }

cudaError_t launch_mesh_wavefront(const mesh_params& scene, const frame_params& frame, const mesh_wave_buffers& q,
                                  float* const layers[4], unsigned long long* counters, int sm_count, cudaStream_t stream,
                                  uint32_t* launches, const std::function<void()>& first_hit_layers_enqueued) {
    const uint32_t samples = frame.batch_samples > 1u ? frame.batch_samples : 1u;
    const uint64_t n_px = static_cast<uint64_t>(frame.width) * frame.height;
    if (n_px == 0) return cudaSuccess;
    const uint32_t n_tiles = ((frame.width + 7u) / 8u) * ((frame.height + 3u) / 4u);
    // queue 0 has an entry for every lane of every tile, inside the image or not
    if (static_cast<uint64_t>(n_tiles) * 32u * samples > q.capacity || scene.max_bounces + 1u > k_mesh_wave_counters) {
        return cudaErrorInvalidValue;
    }
    static const cudaError_t attr[3] = {
        cudaFuncSetAttribute(mesh_wave_trace_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(k_stack_bytes)),
        cudaFuncSetAttribute(mesh_wave_trace_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(k_stack_bytes)),
        cudaFuncSetAttribute(mesh_wave_trace_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(k_stack_bytes))};
    for (cudaError_t e : attr) {
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaMemsetAsync(q.counts, 0, 2 * k_mesh_wave_counters * sizeof(uint32_t), stream);  // counts, then heads
    if (e != cudaSuccess) return e;
    const uint32_t streaming_blocks = static_cast<uint32_t>(sm_count) * 8u;
    const uint64_t raygen_needed = (static_cast<uint64_t>(n_tiles) * samples + k_wave_warps_per_block - 1u) / k_wave_warps_per_block;
    mesh_wave_raygen_kernel<<<static_cast<uint32_t>(raygen_needed < streaming_blocks ? raygen_needed : streaming_blocks),
                              k_block_threads, 0, stream>>>(frame, q, samples);
    uint32_t made = 1;
    const uint64_t resolve_needed = (n_px + k_block_threads - 1u) / k_block_threads;
    const uint32_t resolve_blocks = static_cast<uint32_t>(resolve_needed < streaming_blocks ? resolve_needed : streaming_blocks);
    float4* const l[4] = {reinterpret_cast<float4*>(layers[0]), reinterpret_cast<float4*>(layers[1]),
                          reinterpret_cast<float4*>(layers[2]), reinterpret_cast<float4*>(layers[3])};
    const uint32_t trace_blocks = static_cast<uint32_t>(sm_count) * k_wave_blocks_per_sm;
    for (uint32_t wave = 0; wave < scene.max_bounces; wave++) {
        const uint32_t in = wave & 1u;
        auto* trace = scene.wide != 0u          ? mesh_wave_trace_kernel<true, false>
                      : scene.pairs != nullptr ? mesh_wave_trace_kernel<false, true>
                                               : mesh_wave_trace_kernel<false, false>;
        trace<<<trace_blocks, k_block_threads, k_stack_bytes, stream>>>(scene, q.origin[in], q.direction[in], q.hit,
                                                                        q.counts + wave,
                                                                        q.counts + k_mesh_wave_counters + wave, counters);
        mesh_wave_shade_kernel<<<streaming_blocks, k_block_threads, 0, stream>>>(scene, q, wave);
        made += 2;
        if (wave == 0u) {  // albedo, normal and depth are first-hit values: final from here on
            mesh_wave_resolve_kernel<false><<<resolve_blocks, k_block_threads, 0, stream>>>(frame, q, samples, l[0], l[1], l[2], l[3]);
            made += 1;
            if (first_hit_layers_enqueued) first_hit_layers_enqueued();
        }
    }
    mesh_wave_resolve_kernel<true><<<resolve_blocks, k_block_threads, 0, stream>>>(frame, q, samples, l[0], l[1], l[2], l[3]);
    made += 1;
    if (launches) *launches = made;
    return cudaGetLastError();
}

#endif  // !QUASI_HOST_EMULATION

}  // namespace Q::cuda
