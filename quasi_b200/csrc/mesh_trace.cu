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
/// Bound: instruction issue at low SIMD efficiency, then L1/L2 latency; not HBM
/// (the 20.6 MB of nodes + 48 MB of triangles sit in the
/// 126 MB L2). What the kernel is built around, in the order it mattered (ncu,
/// profiles/r1_mesh_*):
///  - Lanes of a warp that follow different rays must be kept together by hand.
///    Every lane is a small state machine (idle / at an interior node / holding a
///    leaf / waiting to be shaded) and the warp picks ONE kind of work per
///    iteration by ballot: descend one level for the lanes at interior nodes;
///    test triangles once k_leaf_batch lanes hold a leaf; shade and hand out new
///    rays once k_min_waiting lanes have finished their query (persistent
///    while-while with dynamic fetch, with paths instead of rays). The first
///    version, a per-lane loop, ran at 5.5 of 32 lanes per instruction.
///  - 32-byte nodes read as two float4; siblings are adjacent, so visiting an
///    interior node is one 64-byte line (4 x LDG.128, read-only path).
///  - The slab test is 6 FFMA against precomputed origin * 1/dir. It only
///    prunes, so it may be approximate as long as it is conservative: boxes are
///    padded by 1e-4 of their extent + 1e-5, the test's error is below 2e-6.
///  - Triangles in leaf order as three float4 arrays (SoA): 3 x LDG.128 each,
///    consecutive slots of a leaf are consecutive addresses. Moeller-Trumbore
///    in the oracle's operation order: the accepted t reaches the pixel.
///  - Normal and cosine-sampling frame per triangle are precomputed on the host
///    with the shader's operations (mesh_scene.hpp:shading_record) and read
///    once per hit: three normalisations less per bounce.
///  - Per-lane traversal stack in shared memory, laid out [entry][thread] so a
///    push or pop by the whole warp touches 32 different banks. Entries carry
///    the entry distance (16 bits, rounded down), and a popped subtree that has
///    fallen behind the current closest hit is skipped without touching memory.
///    6 bytes per entry: 48 KB per CTA, four CTAs per SM.
/// The order in which nodes are visited cannot change the result: boxes only
/// prune what the ray provably misses, and an equal t is won by the higher
/// triangle index.
///
/// Second form, selected per scene (mesh_params::wide, QUASI_CUDA_MESH_WIDE=1): the same
/// kernel over the 4-wide tree collapsed from the binary one (wide_bvh.hpp). A visit is
/// one 128-byte line (7 x LDG.128), four slab tests, the nearest child hit taken and the
/// others deferred nearest on top; half as many visits, dependent loads and scheduler
/// iterations per ray.

#include "mesh_traverse.cuh"

#include <cuda_runtime.h>

// tests/cuda_emulation/run_mesh.cpp with -DQUASI_EMU_STATS: how many warp-level passes each phase
// of the kernel makes and how many lanes work in them (emu::phase_pass). Nothing on the device.
#if defined(QUASI_HOST_EMULATION) && defined(QUASI_EMU_STATS)
#define QUASI_EMU_PASS(phase, lane_works) emu::phase_pass(phase, __ballot_sync(0xffffffffu, lane_works));
#else
#define QUASI_EMU_PASS(phase, lane_works)  // nothing at all: the device build must see the tokens it saw without this
#endif

namespace Q::cuda {

namespace {

using namespace dev;

#ifndef QUASI_MESH_BLOCKS_PER_SM
#define QUASI_MESH_BLOCKS_PER_SM 4
#endif
#ifndef QUASI_MESH_MIN_WAITING
#define QUASI_MESH_MIN_WAITING 16
#endif
#ifndef QUASI_MESH_LEAF_BATCH
#define QUASI_MESH_LEAF_BATCH 8
#endif
#ifndef QUASI_MESH_RUN_EXIT
#define QUASI_MESH_RUN_EXIT 4
#endif

constexpr int k_blocks_per_sm = QUASI_MESH_BLOCKS_PER_SM;
constexpr int k_warps_per_block = k_block_threads / 32;
constexpr int k_min_waiting = QUASI_MESH_MIN_WAITING;  ///< finished queries that trigger a service phase
constexpr int k_leaf_batch = QUASI_MESH_LEAF_BATCH;    ///< lanes holding a leaf that trigger the triangle phase
constexpr int k_run_exit = QUASI_MESH_RUN_EXIT;        ///< lanes leaving a descent run that end it
#ifndef QUASI_MESH_STREAM_LAYERS
#define QUASI_MESH_STREAM_LAYERS 0
#endif
/// accumulate_pixel (device_common.cuh) on a layer this launch touches exactly once per pixel: with
/// QUASI_MESH_STREAM_LAYERS the history is read and the result written with the evict-first hints (ld/st.global.cs),
/// so that the 266 MB of layer traffic per launch do not push the nodes and triangles out of the L2.
__device__ __forceinline__ void accumulate_once(float4* __restrict__ layer, size_t idx, float4 sample, bool first,
                                                float weight) {
#if QUASI_MESH_STREAM_LAYERS && !defined(QUASI_HOST_EMULATION)
    if (!first) sample = mix4(__ldcs(layer + idx), sample, weight);
    __stcs(layer + idx, sample);
#else
    accumulate_pixel(layer, idx, sample, first, weight);
#endif
}

template <bool Wide>
__global__ void __launch_bounds__(k_block_threads, k_blocks_per_sm)
mesh_pathtrace_kernel(const __grid_constant__ mesh_params scene, const __grid_constant__ frame_params frame,
                      float4* __restrict__ beauty, float4* __restrict__ albedo, float4* __restrict__ normal,
                      float4* __restrict__ depth, unsigned long long* __restrict__ counters,
                      unsigned int* __restrict__ tile_counters) {
#if defined(QUASI_HOST_EMULATION)  // tests/cuda_emulation: the dynamic shared memory of this CTA
    uint32_t* s_refs = static_cast<uint32_t*>(emu::dynamic_shared());
#else
    extern __shared__ uint32_t s_refs[];  // [k_mesh_stack_entries][k_block_threads], then the uint16 distances
#endif
    __shared__ float s_material[k_mesh_materials * 8];
    for (int i = threadIdx.x; i < static_cast<int>(k_mesh_materials) * 8; i += k_block_threads) {
        s_material[i] = scene.materials[i];
    }
    __syncthreads();
    lane_stack stack;
    stack.refs = s_refs + threadIdx.x;
    stack.dist = reinterpret_cast<uint16_t*>(s_refs + k_mesh_stack_entries * k_block_threads) + threadIdx.x;
    const float4* __restrict__ nodes = scene.nodes;

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lanes_below = (1u << lane) - 1u;
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
    mesh_hit best{1e30f, k_miss, 0u};
    v3 inv = make_v3(0.0f, 0.0f, 0.0f), oi = inv;

    uint32_t rays = 0, nodes_visited = 0, tris_tested = 0;
    uint32_t next_item = 0;  // warp-uniform cursor into this warp's pixel sequence
    tile_feed feed{tile_counters, n_tiles};

    while (true) {
        // ---------------------------------------------------------- service ----
        bool new_ray = false, fresh = false;
        v3 fu = make_v3(0.0f, 0.0f, 0.0f), fv = fu, fw = fu;  // scattering frame of a continuing path
        float uv_x = 0.0f, uv_y = 0.0f;                        // pixel centre of a new path
        QUASI_EMU_PASS(0, active && cur == k_done)  // service: shading
        if (active && cur == k_done) {
            // One bounce of path_trace (plugin.mm:232-273) around the finished query.
            const bool is_hit = best.prim != k_miss;
            v3 n = make_v3(0.0f, 0.0f, 0.0f), q_emission = n, q_albedo = n;
            bool is_light = false;
            if (is_hit) {
                const float4* rec = scene.tri_shade + 3u * static_cast<size_t>(best.slot & ((1u << mesh::k_slot_bits) - 1u));
                const float4 s0 = __ldg(rec), s1 = __ldg(rec + 1), s2 = __ldg(rec + 2);
                const float* m = s_material + (best.slot >> mesh::k_slot_bits) * 8u;
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

            if (bounce == 0) {
                v3 a = make_v3(0.0f, 0.0f, 0.0f), nn = a;
                float first_t = 0.0f;
                if (is_hit) {
                    a = is_light ? q_emission : q_albedo;
                    nn = make_v3(n.x * 0.5f + 0.5f, n.y * 0.5f + 0.5f, n.z * 0.5f + 0.5f);
                    first_t = best.t;
                }
                const float nd = fminf(fmaxf((first_t - 1.0f) / 4.0f, 0.0f), 1.0f);
                accumulate_once(albedo, pixel, make_float4(a.x, a.y, a.z, 1.0f), first_sample,
                                 weight);
                accumulate_once(normal, pixel, make_float4(nn.x, nn.y, nn.z, 1.0f), first_sample,
                                 weight);
                accumulate_once(depth, pixel, make_float4(nd, nd, nd, 1.0f), first_sample,
                                 weight);
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
                new_ray = true;  // the direction is drawn below, together with the new paths' camera rays
            } else {
                accumulate_once(beauty, pixel, make_float4(color.x, color.y, color.z, 1.0f),
                                 first_sample, weight);
                active = false;
            }
        }
        __syncwarp();

        // Regeneration: lanes without a path take the warp's next pixels (see pathtrace.cu).
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u) {
            feed.open(next_item, next_item + __popc(idle) - 1u);
            if (!active) {
                const uint32_t item = next_item + __popc(idle & lanes_below);
                const uint32_t tile = feed.tile_of(item);
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
                        uv_x = (static_cast<float>(px) + 0.5f) / static_cast<float>(frame.width);
                        uv_y = 1.0f - (static_cast<float>(py) + 0.5f) / static_cast<float>(frame.height);
                        rng = pcg_hash(px + py * 1920u + frame.frame_index * 1920u * 1080u);
                        fresh = true;
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

        // Every lane with a new ray: its direction (get_camera_ray, plugin.mm:129-153, for a new
        // path; cosine_sample_hemisphere, plugin.mm:111-125, with the per-triangle frame for a
        // continuing one: two draws and a normalisation either way, issued once for both kinds),
        // then the root test and the start of the query.
        QUASI_EMU_PASS(1, new_ray)  // new rays
        if (new_ray) {
            const float draw_a = random_float(rng);
            const float draw_b = random_float(rng);
            v3 dir;
            if (fresh) {
                uv_x += (draw_a - 0.5f) * 0.001f;
                uv_y += (draw_b - 0.5f) * 0.001f;
                const v3 target = load_v3(frame.lower_left) + load_v3(frame.horizontal) * uv_x +
                                  load_v3(frame.vertical) * uv_y;
                ro = cam_origin;
                dir = target - cam_origin;
            } else {
                const float phi = 6.28318548202514648f * draw_a;
                const float cos_theta = sqrtf(1.0f - draw_b);
                const float sin_theta = sqrtf(draw_b);
                float sin_phi, cos_phi;
                sincos_round_once(phi, sin_phi, cos_phi);
                dir = (fu * cos_phi) * sin_theta + (fv * sin_phi) * sin_theta + fw * cos_theta;
            }
            rd = normalize(dir);
            rays++;
            best = mesh_hit{1e30f, k_miss, 0u};
            inv = make_v3(slab_reciprocal(rd.x), slab_reciprocal(rd.y), slab_reciprocal(rd.z));
            oi = ro * inv;
            stack.sp = 0;
            if constexpr (Wide) {
                cur = 0u;  // the root of the wide tree; its children's boxes are the first test
            } else {
                float tn;
                const float4 lo = __ldg(nodes), hi = __ldg(nodes + 1);
                nodes_visited++;
                cur = box_hit(lo, hi, inv, oi, best.t, tn) ? node_ref(lo, hi) : k_done;
            }
        }
        const uint32_t m_active = __ballot_sync(0xffffffffu, active);
        if (m_active == 0u) {
            if (feed.exhausted(next_item)) break;  // no paths, no pixels left
            continue;  // only off-image pixels of an edge tile were handed out
        }

        // --------------------------------------------------------- traverse ----
        while (true) {
            // low 3 bits of cur: 0 interior, 1..4 leaf with that many triangles, 6 k_pop, 7 k_done
            const bool descending = (cur & 7u) == 0u || cur == k_pop;
            const uint32_t m_descending = __ballot_sync(0xffffffffu, descending);
            const uint32_t m_leaf = __ballot_sync(0xffffffffu, (cur & 7u) - 1u < 4u);
            const uint32_t m_tracing = m_descending | m_leaf;
            // enough finished queries to make shading worth issuing, or nothing to traverse
            if (m_tracing == 0u || __popc(m_active & ~m_tracing) >= k_min_waiting) break;
            QUASI_EMU_PASS(2, m_tracing != 0u)  // scheduler iterations
#if defined(QUASI_HOST_EMULATION) && defined(QUASI_EMU_STATS)
            if (!(m_descending != 0u && __popc(m_leaf) < k_leaf_batch)) {
                for (uint32_t k = 0; k < 4u; k++) QUASI_EMU_PASS(4, (cur & 7u) - 1u < 4u && (cur & 7u) > k)  // triangle passes
            }
#endif
            if (m_descending != 0u && __popc(m_leaf) < k_leaf_batch) {
                // Descent run: one level per iteration for every lane that is not holding a
                // leaf, until k_run_exit of them have reached a leaf or finished. Leaf
                // holders sit this out; they are served once k_leaf_batch have gathered.
                const int keep = max(__popc(m_descending) - k_run_exit, 0);
                do {
                    QUASI_EMU_PASS(3, (cur & 7u) == 0u || cur == k_pop)  // descent passes
                    if (cur == k_pop) cur = stack.pop_one(best.t);
                    if ((cur & 7u) == 0u) {
                        if constexpr (Wide) {
                            nodes_visited += mesh::k_wide;
                            wide_step(nodes, stack, inv, oi, best.t, cur);
                        } else {
                            nodes_visited += 2u;
                            interior_step(nodes, stack, inv, oi, best.t, cur);
                        }
                    }
                } while (__popc(__ballot_sync(0xffffffffu, (cur & 7u) == 0u || cur == k_pop)) > keep);
            } else if ((cur & 7u) - 1u < 4u) {
                const uint32_t first = cur >> 3, count = cur & 7u;
                for (uint32_t k = 0; k < count; k++) {
                    const uint32_t slot = first + k;
                    intersect(__ldg(scene.tri_v0 + slot), __ldg(scene.tri_e1 + slot), __ldg(scene.tri_e2 + slot), ro, rd, best);
                }
                tris_tested += count;
                cur = stack.after_node();
            }
        }
    }

    feed.finish(gridDim.x * k_warps_per_block);
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

#if !defined(QUASI_HOST_EMULATION)  // tests/cuda_emulation has its own launcher
cudaError_t launch_mesh_pathtrace(const mesh_params& scene, const frame_params& frame, float* const layers[4],
                                  unsigned long long* counters, unsigned int* tile_counters, int sm_count,
                                  cudaStream_t stream, const l2_window& keep) {
    const uint32_t n_tiles = ((frame.width + 7u) / 8u) * ((frame.height + 3u) / 4u);
    uint32_t blocks = static_cast<uint32_t>(sm_count) * k_blocks_per_sm;
    const uint32_t needed = (n_tiles + k_warps_per_block - 1u) / k_warps_per_block;
    if (needed < blocks) blocks = needed;
    if (blocks == 0) return cudaSuccess;
    auto* kernel = scene.wide != 0u ? mesh_pathtrace_kernel<true> : mesh_pathtrace_kernel<false>;
    static const cudaError_t attr[2] = {
        cudaFuncSetAttribute(mesh_pathtrace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(k_stack_bytes)),
        cudaFuncSetAttribute(mesh_pathtrace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(k_stack_bytes))};
    for (cudaError_t e : attr) {
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t config{};
    config.gridDim = dim3(blocks);
    config.blockDim = dim3(k_block_threads);
    config.dynamicSmemBytes = k_stack_bytes;
    config.stream = stream;
    cudaLaunchAttribute attribute{};
    if (keep.bytes != 0) {
        attribute.id = cudaLaunchAttributeAccessPolicyWindow;
        attribute.val.accessPolicyWindow.base_ptr = const_cast<void*>(keep.base);
        attribute.val.accessPolicyWindow.num_bytes = keep.bytes;
        attribute.val.accessPolicyWindow.hitRatio = keep.hit_ratio;
        attribute.val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attribute.val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        config.attrs = &attribute;
        config.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&config, kernel, scene, frame, reinterpret_cast<float4*>(layers[0]),
                              reinterpret_cast<float4*>(layers[1]), reinterpret_cast<float4*>(layers[2]),
                              reinterpret_cast<float4*>(layers[3]), counters, tile_counters);
}

#endif  // !QUASI_HOST_EMULATION

}  // namespace Q::cuda
