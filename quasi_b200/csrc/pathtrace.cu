/// @file pathtrace.cu
/// @brief Fused camera-ray / intersection / shading / accumulate kernel for sm_100a.
///
/// One launch does what Q_plugin_render's PASS 1 (pathtrace_frag into four MRT
/// targets) and PASS 2 (four accumulate dispatches) do in the Metal backend
/// (/root/reference/backends/metal/plugin.mm:715-774): every pixel gets one new
/// sample, folded straight into the four accumulated RGBA32F layers in place.
/// The 64 B/px MRT store, its re-read and the ping-pong copies never exist:
/// 128 B/px of HBM traffic per call instead of 276.
///
/// Execution model (DESIGN.md "Kernels"):
///  - Persistent warps, wavefront inside the warp. The grid is resident CTAs only (4 per SM);
///    a warp takes 8x4 pixel tiles from a device-wide work queue (one atomic per tile) into a
///    queue of primary items in shared memory, paths that go on after a hit go to a queue of
///    bounce items, and every warp iteration pops 32 items of ONE queue: all lanes of an
///    iteration do the same kind of work (pathtrace_accumulate_kernel below). In the
///    thread-per-pixel form the warp ran at 18 of 32 lanes (ncu, profiles/r1_v1_*), in the
///    lane-owns-a-path form of round 1 at 23.8, here at 28.4.
///  - The 16-quad scan is fully unrolled; plane, edge and margin constants are
///    immediate constant-bank operands. The scan DECIDES with cheap forms
///    (MUFU.RCP instead of IEEE division, one dot product per barycentric
///    instead of two cross products) plus explicit margins, then computes the
///    accepted hit with the literal IEEE sequence. Anything inside a margin
///    (an edge, a near tie between two quads, t next to t_min) goes to the
///    literal scan, so the result is bit-identical to the shader's loop.
///  - Shading constants of the hit quad are read from shared memory (the index
///    diverges across lanes; constant-bank loads would serialise).
///
/// Arithmetic contract (DESIGN.md "Numerics"): -fmad=false, IEEE division and
/// square root; every value that reaches a pixel is computed with the shader's
/// operation order. sin/cos are evaluated in binary64 and rounded once.

#include "device_common.cuh"
#include "kernels.hpp"
#include "scan_decide.hpp"

#include <cuda_runtime.h>

namespace Q::cuda {

namespace {

using namespace dev;
using namespace scan;

struct hit_record {
    float t;
    v3 point;
    int index;   ///< quad index, -1 = miss
    bool front;  ///< denom < 0: the stored normal faces the ray
};

/// The shader's loop, literally: trace_scene + intersect_quad, plugin.mm:157-212.
/// Used for the rare rays the fast scan cannot decide with its margins.
__device__ __noinline__ void trace_scene_exact(const scene_params& scene, v3 ro, v3 rd,
                                               hit_record& closest) {
    closest.t = 1e30f;
    closest.index = -1;
    closest.front = true;
    closest.point = make_v3(0.0f, 0.0f, 0.0f);
    const uint32_t count = scene.quad_count;
#pragma unroll 1
    for (uint32_t k = 0; k < count; k++) {  // original order: the later index wins a tie
        const uint32_t i = scene.scene_order[k];
        const quad_params& q = scene.quads[i];
        if (q.valid == 0.0f) continue;
        v3 normal = load_v3(q.normal);
        float denom = dot(normal, rd);
        if (fabsf(denom) < 1e-8f) continue;
        float t = (q.plane_d - dot(normal, ro)) / denom;
        if (t < 0.001f || t > closest.t) continue;

        v3 p = ro + rd * t;
        v3 planar = p - load_v3(q.origin);
        v3 w = load_v3(q.w);
        float alpha = dot(w, cross(planar, load_v3(q.v)));
        float beta = dot(w, cross(load_v3(q.u), planar));
        if (alpha < 0.0f || alpha > 1.0f || beta < 0.0f || beta > 1.0f) continue;

        closest.t = t;
        closest.point = p;
        closest.index = static_cast<int>(i);
        closest.front = denom < 0.0f;
    }
}

constexpr int k_shade_stride = 25;  // odd: lanes reading different quads hit different banks

// Offsets inside the shared-memory shading record (quad_params floats 28..51).
constexpr int k_sh_normal = 0, k_sh_is_light = 3, k_sh_frame_u = 4, k_sh_plane_d = 7,
              k_sh_frame_v = 8, k_sh_frame_w = 12, k_sh_albedo = 16, k_sh_emission = 20;

/// How many quads of each shape a scene has, known at compile time: the scan is
/// then fully unrolled with every constant an immediate constant-bank operand.
/// The sorted table order is general, flat_x, flat_y, flat_z, upright.
template <int GEN, int FX, int FY, int FZ, int UP, int BOX>
struct fixed_layout {
    static constexpr bool fixed = true;
    static constexpr int general = GEN, flat_x = FX, flat_y = FY, flat_z = FZ, upright = UP, boxes = BOX;
    static constexpr int begin_general = 0, begin_flat_x = GEN, begin_flat_y = GEN + FX,
                         begin_flat_z = GEN + FX + FY, begin_upright = GEN + FX + FY + FZ;
    static bool matches(const scene_params& s) {
        const int want[k_shape_count + 1] = {0, GEN, GEN + FX, GEN + FX + FY, GEN + FX + FY + FZ,
                                             GEN + FX + FY + FZ + UP};
        for (uint32_t k = 0; k <= k_shape_count; k++) {
            if (s.shape_begin[k] != static_cast<uint32_t>(want[k])) return false;
        }
        return s.box_count == static_cast<uint32_t>(BOX);
    }
};
/// Any scene: one loop per shape with launch-uniform bounds.
struct dynamic_layout {
    static constexpr bool fixed = false;
};
/// The demo scene (make_cornell_box): 2 side walls, floor + ceiling + light + 2 block tops,
/// the back wall, 8 block sides.
using cornell_quads_layout = fixed_layout<0, 2, 5, 1, 8, 0>;
/// The same scene with each block's five faces as one slab test (QUASI_CUDA_SCAN=boxes).
using cornell_layout = fixed_layout<0, 2, 3, 1, 0, 2>;

/// Closest hit of one ray against the scene. Same result, bit for bit,
/// as trace_scene_exact. `shade` is the shared-memory copy of the quads' shading
/// records; the winner's plane is re-read from it for the literal finalisation.
///
/// plugin.mm:177 (t < t_min) and plugin.mm:186 (alpha, beta in [0,1]) are decided with margins;
/// plugin.mm:174 (|denom| < 1e-8) exactly, on the winner (scan_decide.hpp:accept_winner).
template <class Layout>
__device__ __forceinline__ void trace_scene(const scene_params& scene, const float* shade, v3 ro,
                                            v3 rd, hit_record& hit) {
    scan_state s;
    if constexpr (Layout::fixed) {
#pragma unroll
        for (int i = Layout::begin_flat_y; i < Layout::begin_flat_y + Layout::flat_y; i++) {
            scan_flat<1>(s, QUASI_SCAN_SOURCE(scene, i), i, ro, rd);
        }
#pragma unroll
        for (int i = Layout::begin_upright; i < Layout::begin_upright + Layout::upright; i++) {
            scan_upright(s, QUASI_SCAN_SOURCE(scene, i), i, ro, rd);
        }
#pragma unroll
        for (int i = Layout::begin_flat_x; i < Layout::begin_flat_x + Layout::flat_x; i++) {
            scan_flat<0>(s, QUASI_SCAN_SOURCE(scene, i), i, ro, rd);
        }
#pragma unroll
        for (int i = Layout::begin_flat_z; i < Layout::begin_flat_z + Layout::flat_z; i++) {
            scan_flat<2>(s, QUASI_SCAN_SOURCE(scene, i), i, ro, rd);
        }
#pragma unroll
        for (int i = Layout::begin_general; i < Layout::begin_general + Layout::general; i++) {
            scan_general(s, scene.quads[i], i, ro, rd);
        }
#pragma unroll
        for (int n = 0; n < Layout::boxes; n++) scan_box(s, scene.boxes[n], ro, rd);
    } else {
        // One loop per shape over that shape's range of the sorted quad table.
        // Trip counts and table offsets are launch-uniform.
#pragma unroll 1
        for (uint32_t i = scene.shape_begin[k_shape_flat_y]; i < scene.shape_begin[k_shape_flat_y + 1]; i++) {
            scan_flat<1>(s, QUASI_SCAN_SOURCE(scene, i), static_cast<int>(i), ro, rd);
        }
#pragma unroll 1
        for (uint32_t i = scene.shape_begin[k_shape_upright]; i < scene.shape_begin[k_shape_upright + 1]; i++) {
            scan_upright(s, QUASI_SCAN_SOURCE(scene, i), static_cast<int>(i), ro, rd);
        }
#pragma unroll 1
        for (uint32_t i = scene.shape_begin[k_shape_flat_x]; i < scene.shape_begin[k_shape_flat_x + 1]; i++) {
            scan_flat<0>(s, QUASI_SCAN_SOURCE(scene, i), static_cast<int>(i), ro, rd);
        }
#pragma unroll 1
        for (uint32_t i = scene.shape_begin[k_shape_flat_z]; i < scene.shape_begin[k_shape_flat_z + 1]; i++) {
            scan_flat<2>(s, QUASI_SCAN_SOURCE(scene, i), static_cast<int>(i), ro, rd);
        }
#pragma unroll 1
        for (uint32_t i = scene.shape_begin[k_shape_general]; i < scene.shape_begin[k_shape_general + 1]; i++) {
            scan_general(s, scene.quads[i], static_cast<int>(i), ro, rd);
        }
#pragma unroll 1
        for (uint32_t n = 0; n < scene.box_count; n++) scan_box(s, scene.boxes[n], ro, rd);
    }

    if (needs_literal_scan(s)) {
        trace_scene_exact(scene, ro, rd, hit);
        return;
    }
    hit.index = s.best_i;
    if (s.best_i < 0) return;
    // The accepted hit, with the shader's operations: plugin.mm:173, 176, 179.
    const float* sh = shade + s.best_i * k_shade_stride;
    const v3 normal = load_v3(sh + k_sh_normal);
    const float denom = dot(normal, rd);
    if (!accept_winner(denom)) {  // plugin.mm:174, taken once per ray instead of once per quad
        trace_scene_exact(scene, ro, rd, hit);
        return;
    }
    hit.front = denom < 0.0f;
    hit.t = (sh[k_sh_plane_d] - dot(normal, ro)) / denom;
    hit.point = ro + rd * hit.t;
}

#ifndef QUASI_QUAD_BLOCK_THREADS
#define QUASI_QUAD_BLOCK_THREADS 256
#endif
#ifndef QUASI_QUAD_BLOCKS_PER_SM
#define QUASI_QUAD_BLOCKS_PER_SM 4
#endif
constexpr int k_block_threads = QUASI_QUAD_BLOCK_THREADS;
constexpr int k_blocks_per_sm = QUASI_QUAD_BLOCKS_PER_SM;
constexpr int k_warps_per_block = k_block_threads / 32;

/// System-scope arrival counter on memory that may belong to another GPU, with release semantics: everything this
/// thread wrote, and everything it has observed (the other lanes' stores, through the __syncwarp before), is
/// visible to whoever reads the new count. atom.release and not __threadfence_system() + atomicAdd_system: the
/// fence is MEMBAR.SC.SYS followed by CCTL.IVALL, an invalidation of the SM's whole L1 per finished tile, which
/// cost the render 4 % on two GPUs (the tiles' own accumulators live there between samples); the release atomic
/// is MEMBAR.ALL.SYS + ATOMG and leaves the L1 alone.
__device__ __forceinline__ unsigned int arrive_release_system(unsigned int* p) {
#if defined(QUASI_HOST_EMULATION)
    return std::atomic_ref<unsigned int>(*p).fetch_add(1u, std::memory_order_release);
#else
    unsigned int before;
    asm volatile("atom.add.release.sys.global.u32 %0, [%1], 1;" : "=r"(before) : "l"(p) : "memory");
    return before;
#endif
}
/// The other half, executed only by the thread whose arrival was the last: what the earlier arrivals released is
/// visible to the loads that follow.
__device__ __forceinline__ void acquire_fence_system() {
#if defined(QUASI_HOST_EMULATION)
    std::atomic_thread_fence(std::memory_order_acquire);
#else
    asm volatile("fence.acq_rel.sys;" ::: "memory");
#endif
}
/// A float4 another GPU may have written a moment ago: not from this SM's L1.
__device__ __forceinline__ float4 load_uncached(const float4* p) {
#if defined(QUASI_HOST_EMULATION)
    return *p;
#else
    return __ldcv(p);
#endif
}

// ------------------------------------------------------------------------------------------------
// Path state lives in per-warp queues in shared memory; every warp iteration is ONE kind of work
// for 32 paths.
//
// Round 1's kernel let a lane own a path from its camera ray to its end, and each iteration of the
// warp served two populations at once: about 14 lanes that had just taken a new pixel (tile
// decode, uv, seed, camera ray, and after the scan the three first-hit AOV read-modify-writes) and
// about 18 that continued a path (binary64 sin/cos, the scatter frame). SIMT executes the two one
// after the other, so ~380 of the ~1210 instructions of an iteration ran at half the lanes (ncu,
// profiles/r2_quad_lane_owned_regions.txt: 23.8 of 32 lanes overall, 40.0 warp instructions per ray).
// Here a warp keeps two queues in shared memory:
//   P  primary items  (pixel, sample of the launch): a pixel that needs a camera ray
//   B  bounce items   (origin, RNG state, the quad scattered from, throughput, colour, pixel)
// and every iteration pops up to 32 items of ONE queue, traces one ray for each and pushes the
// paths that go on to B. The queue is the compaction: all 32 lanes of an iteration run the same
// code, new-pixel work is issued only in primary iterations and scatter work only in bounce
// iterations, and only the origin, the direction and the RNG state are live across the scan
// (throughput, colour and pixel stay in the queue slot until the hit is shaded).
// Measured (B200, 1080p, 64 spp, one launch per sample): 26.5 against 22.9 Grays/s, 34.4 against
// 40.0 warp instructions per ray at 28.4 against 23.8 lanes (profiles/r2_wavefront_*).
// Policy (warp-uniform): B holds 32 -> bounce iteration; else P holds 32 -> primary iteration; else
// take a tile from the device-wide queue into P; when that has run dry, drain P, then B.
// An item yields at most one item and a tile is taken only when both queues are below 32, so B never
// exceeds 63 items and P never 95: 64 and 128 slots.
//
// `Batch` (frame.batch_samples = k > 1, plugin.cu:QUASI_CUDA_BATCH): one launch gives every pixel k
// samples. The samples of a pixel must be folded in in order (the running mean of plugin.mm:326), so
// a tile stays with the warp that took it: when a path ends before the pixel's last sample of the
// launch, the pixel goes back into P with the next sample. Every read-modify-write of a pixel is then
// separated from the previous one by a __syncwarp of the one warp that touches that pixel in the launch.
// A slot table in shared memory counts, per tile the warp holds, the pixels that have not had their last
// sample; when the count reaches zero the warp EXPORTS the tile (batch_params): the 32 pixels of all four
// layers, rows of 8 pixels (128 bytes) at a time, to pinned host memory (the plugin's readback mirror: the
// frame is on the host when the launch ends, its 64 B/px transfer spread over the launch that renders it)
// or to the GPUs that own the slices (multi-GPU: the reduce-scatter is part of the render). Tiles finish
// all through the launch, so the export overlaps the rendering of the rest of the image.
// Tried and dropped (round 2, profiles/r2_call3.log): k passes over the image inside the launch, entry
// q = tile q % n_tiles at sample q / n_tiles, ordered by a per-tile epoch in global memory. It removes the
// tail of this form (the last pixels of a warp walk through their k samples one after the other, 4.8 % of
// the SM cycles at k = 64) but every tile then finishes in the last pass, so the whole export lands at the
// end (14.5 ms per frame instead of 12.0 with the export off), and its fences and L2 history reads cost
// 4 % more instructions at 67 % issue utilisation instead of 82 %.
constexpr uint32_t k_bq = 64, k_pq = 128;  // queue slots (powers of two)
constexpr uint32_t k_bq_fields = 13;       // ro.xyz, rng, meta, throughput.xyz, colour.xyz, pixel_xy, tag
constexpr uint32_t k_tile_slots = 16;      // Batch: tiles a warp may hold at once
constexpr uint32_t k_queue_words = k_bq_fields * k_bq + 2 * k_pq + 2 * k_tile_slots;  // per warp
constexpr size_t k_wavefront_smem = static_cast<size_t>(k_warps_per_block) * k_queue_words * sizeof(uint32_t);

template <int NQ, class Layout, bool Batch = false, bool Reduce = false>
__global__ void __launch_bounds__(k_block_threads, k_blocks_per_sm)
pathtrace_accumulate_kernel(const __grid_constant__ scene_params scene, const __grid_constant__ frame_params frame,
                           const __grid_constant__ batch_params launch, float4* __restrict__ beauty,
                           float4* __restrict__ albedo, float4* __restrict__ normal, float4* __restrict__ depth,
                           unsigned long long* __restrict__ ray_counter, unsigned int* __restrict__ tile_counters) {
    __shared__ float s_shade[NQ * k_shade_stride];
    for (int idx = threadIdx.x; idx < NQ * static_cast<int>(k_shade_floats); idx += k_block_threads) {
        const int qi = idx / static_cast<int>(k_shade_floats);
        const int k = idx - qi * static_cast<int>(k_shade_floats);
        s_shade[qi * k_shade_stride + k] = reinterpret_cast<const float*>(&scene.quads[qi])[k_shade_offset + k];
    }
#if defined(QUASI_HOST_EMULATION)
    uint32_t* s_queues = static_cast<uint32_t*>(emu::dynamic_shared());
#else
    extern __shared__ uint32_t s_queues[];
#endif
    uint32_t* const q = s_queues + (threadIdx.x >> 5) * k_queue_words;
    float* const b_ro = reinterpret_cast<float*>(q);              // [3][k_bq]
    uint32_t* const b_rng = q + 3 * k_bq;
    uint32_t* const b_meta = q + 4 * k_bq;                        // bounce | quad << 8 | hit from behind << 16
    float* const b_thr = reinterpret_cast<float*>(q + 5 * k_bq);  // [3][k_bq]
    float* const b_col = reinterpret_cast<float*>(q + 8 * k_bq);  // [3][k_bq]
    uint32_t* const b_pxy = q + 11 * k_bq;
    uint32_t* const b_tag = q + 12 * k_bq;                        // Batch: tile slot | sample << 4
    uint32_t* const p_xy = q + k_bq_fields * k_bq;
    uint32_t* const p_tag = p_xy + k_pq;
    uint32_t* const t_tile = p_tag + k_pq;                        // Batch: tile of a slot
    uint32_t* const t_left = t_tile + k_tile_slots;               // Batch: its paths that have not ended yet
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lanes_below = (1u << lane) - 1u;
    const uint32_t tiles_x = (frame.width + 7u) >> 3;
    const uint32_t n_tiles = tiles_x * ((frame.height + 3u) >> 2);
    const uint32_t n_px = frame.width * frame.height;
    const uint32_t max_bounces = scene.max_bounces;
    const v3 cam_origin = load_v3(frame.position);
    const launch_divisor by_width(static_cast<float>(frame.width)), by_height(static_cast<float>(frame.height));
    [[maybe_unused]] const uint32_t batch = frame.batch_samples > 1u ? frame.batch_samples : 1u;

    uint32_t b_head = 0, b_count = 0, p_head = 0, p_count = 0;  // warp-uniform
    [[maybe_unused]] uint32_t free_slots = (1u << k_tile_slots) - 1u;  // Batch, warp-uniform
    bool dry = false;
    uint32_t rays = 0;

    while (true) {
        bool primary;
        uint32_t n;
        if (b_count >= 32u) {
            primary = false, n = 32u;
        } else if (p_count >= 32u) {
            primary = true, n = 32u;
        } else if (!dry && (!Batch || free_slots != 0u)) {
            // Take the next tile of the device-wide queue.
            uint32_t tile = 0;
            if (lane == 0u) tile = atomicAdd(tile_counters, 1u);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= n_tiles) {
                dry = true;
                continue;
            }
            // The queue hands out columns of tiles from the middle of the image outwards, every column top to
            // bottom: what is left when the queue runs dry are the outermost columns, where most rays leave the
            // scene at once, so the last tiles of a launch are cheap ones and the launch's tail is short. Row by
            // row, the last tiles were a full-price row across the floor of the box: 27.52 -> 27.79 Grays/s at
            // 1080p x 64 spp. The order in which tiles are rendered cannot change a pixel.
            const uint32_t tiles_y = (frame.height + 3u) >> 2;
            const uint32_t column = tile / tiles_y;
            uint32_t tile_y = tile - column * tiles_y;
            uint32_t tile_x = (column & 1u) != 0u ? ((tiles_x - 1u) >> 1) + ((column + 1u) >> 1) : ((tiles_x - 1u) >> 1) - (column >> 1);
            tile = tile_y * tiles_x + tile_x;  // what the tile is called everywhere else
            const uint32_t px = tile_x * 8u + (lane & 7u), py = tile_y * 4u + (lane >> 3);
            const bool inside = px < frame.width && py < frame.height;
            const uint32_t m_inside = __ballot_sync(0xffffffffu, inside);
            uint32_t tag = 0;
            if constexpr (Batch) {
                const uint32_t slot = static_cast<uint32_t>(__ffs(static_cast<int>(free_slots))) - 1u;
                free_slots &= ~(1u << slot);
                tag = slot;  // | sample << 4, sample 0
                if (lane == 0u) {
                    t_tile[slot] = tile;
                    t_left[slot] = static_cast<uint32_t>(__popc(m_inside));
                }
            }
            if (inside) {
                const uint32_t at = (p_head + p_count + __popc(m_inside & lanes_below)) & (k_pq - 1u);
                p_xy[at] = px | (py << 16);  // the ABI's sizes are far below 65536 (plugin.cu checks)
                if constexpr (Batch) p_tag[at] = tag;
            }
            p_count += __popc(m_inside);
            __syncwarp();
            continue;
        } else if (p_count > 0u) {  // the device-wide queue has run dry (or every tile slot is taken): drain
            primary = true, n = p_count;
        } else if (b_count > 0u) {
            primary = false, n = b_count;
        } else {
            break;
        }

        const bool live = lane < n;
        uint32_t slot, rng = 0, pxy = 0, tag = 0, meta = 0;
        v3 ro = cam_origin, d = make_v3(0.0f, 0.0f, -1.0f);
        float uv_x = 0.0f, uv_y = 0.0f;
        if (primary) {
            slot = (p_head + lane) & (k_pq - 1u);
            p_head = (p_head + n) & (k_pq - 1u);
            p_count -= n;
            if (live) {
                pxy = p_xy[slot];
                if constexpr (Batch) tag = p_tag[slot];
                const uint32_t px = pxy & 0xffffu, py = pxy >> 16, smp = tag >> 4;
                // pathtrace_frag, plugin.mm:287-309; uv is the closed form of the rasteriser's
                // interpolation, py = 0 on top.
                uv_x = by_width.quotient(static_cast<float>(px) + 0.5f);
                uv_y = 1.0f - by_height.quotient(static_cast<float>(py) + 0.5f);
                rng = pcg_hash(px + py * 1920u + (frame.frame_index + smp * frame.frame_stride) * 1920u * 1080u);
                if (frame.sample_index + smp != 0u) {  // the history is needed after the scan: start pulling it out of HBM
                    const size_t pixel = static_cast<size_t>(py) * frame.width + px;
                    prefetch_l2(beauty + pixel);
                    prefetch_l2(albedo + pixel);
                    prefetch_l2(normal + pixel);
                    prefetch_l2(depth + pixel);
                }
            }
        } else {
            slot = (b_head + lane) & (k_bq - 1u);
            b_head = (b_head + n) & (k_bq - 1u);
            b_count -= n;
            if (live) {
                ro = make_v3(b_ro[slot], b_ro[k_bq + slot], b_ro[2 * k_bq + slot]);
                rng = b_rng[slot];
                meta = b_meta[slot];
            }
        }

        hit_record hit;
        hit.index = -1;
        hit.front = true;
        hit.t = 0.0f;
        hit.point = make_v3(0.0f, 0.0f, 0.0f);
        v3 rd = d;
        if (live) {
            // Two draws and a normalisation either way: get_camera_ray (plugin.mm:129-153, jitter x
            // before y) in a primary iteration, cosine_sample_hemisphere (plugin.mm:111-125, r1 before
            // r2, with the per-quad frame) in a bounce iteration. The branch is warp-uniform.
            const float draw_a = random_float(rng);
            const float draw_b = random_float(rng);
            if (primary) {
                uv_x += (draw_a - 0.5f) * 0.001f;
                uv_y += (draw_b - 0.5f) * 0.001f;
                const v3 target = load_v3(frame.lower_left) + load_v3(frame.horizontal) * uv_x +
                                  load_v3(frame.vertical) * uv_y;
                d = target - cam_origin;
            } else {
                const bool behind = (meta & 0x10000u) != 0u;
                const float* from = s_shade + ((meta >> 8) & 0xffu) * k_shade_stride;
                const float side = behind ? -1.0f : 1.0f;
                const float phi = 6.28318548202514648f * draw_a;  // 2.0f * M_PI_F, exact product
                const float cos_theta = sqrtf(1.0f - draw_b);
                const float sin_theta = sqrtf(draw_b);
                float sin_phi, cos_phi;
                sincos_round_once(phi, sin_phi, cos_phi);
                const v3 fu = load_v3(from + k_sh_frame_u);
                const v3 fv = load_v3(from + k_sh_frame_v) * side;
                const v3 fw = load_v3(from + k_sh_frame_w) * side;
                d = (fu * cos_phi) * sin_theta + (fv * sin_phi) * sin_theta + fw * cos_theta;
            }
            rd = normalize(d);
            trace_scene<Layout>(scene, s_shade, ro, rd, hit);  // one bounce of path_trace, plugin.mm:232-273
            rays++;
        }

        bool goes_on = false, ended = false;
        v3 throughput = make_v3(1.0f, 1.0f, 1.0f), color = make_v3(0.0f, 0.0f, 0.0f);
        if (live) {
            const bool is_hit = hit.index >= 0;
            const float* sh = s_shade + (is_hit ? hit.index : 0) * k_shade_stride;
            const float side = (is_hit && !hit.front) ? -1.0f : 1.0f;
            const v3 nrm = load_v3(sh + k_sh_normal) * side;  // exact sign flip
            const v3 q_emission = load_v3(sh + k_sh_emission);
            const v3 q_albedo = load_v3(sh + k_sh_albedo);
            const bool is_light = sh[k_sh_is_light] != 0.0f;
            uint32_t bounce = 0;
            if (primary) {
                // First-hit AOVs (plugin.mm:241-249, 300-307).
                v3 a = make_v3(0.0f, 0.0f, 0.0f), nn = a;
                float first_t = 0.0f;
                if (is_hit) {
                    a = is_light ? q_emission : q_albedo;
                    nn = make_v3(nrm.x * 0.5f + 0.5f, nrm.y * 0.5f + 0.5f, nrm.z * 0.5f + 0.5f);
                    first_t = hit.t;
                }
                const float nd = fminf(fmaxf((first_t - 1.0f) / 4.0f, 0.0f), 1.0f);
                const uint32_t sample_index = frame.sample_index + (tag >> 4);
                const bool first = sample_index == 0u;
                const float weight = Batch ? 1.0f / static_cast<float>(sample_index + 1u) : frame.weight;
                const size_t pixel = static_cast<size_t>(pxy >> 16) * frame.width + (pxy & 0xffffu);
                accumulate_rgb1<false>(albedo, pixel, a, first, weight);
                accumulate_rgb1<false>(normal, pixel, nn, first, weight);
                accumulate_grey1<false>(depth, pixel, nd, first, weight);
            } else {
                throughput = make_v3(b_thr[slot], b_thr[k_bq + slot], b_thr[2 * k_bq + slot]);
                color = make_v3(b_col[slot], b_col[k_bq + slot], b_col[2 * k_bq + slot]);
                pxy = b_pxy[slot];
                if constexpr (Batch) tag = b_tag[slot];
                bounce = meta & 0xffu;
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
            done = done || bounce >= max_bounces;  // the shader's last scatter is never observed
            goes_on = !done;
            ended = done;
            if (goes_on) {
                ro = hit.point + nrm * 0.001f;
                meta = bounce | (static_cast<uint32_t>(hit.index) << 8) | (hit.front ? 0u : 0x10000u);
            }
        }
        __syncwarp();  // every lane has read what it needs from the slot it popped

        const uint32_t m_goes_on = __ballot_sync(0xffffffffu, goes_on);
        if (goes_on) {
            const uint32_t at = (b_head + b_count + __popc(m_goes_on & lanes_below)) & (k_bq - 1u);
            b_ro[at] = ro.x, b_ro[k_bq + at] = ro.y, b_ro[2 * k_bq + at] = ro.z;
            b_rng[at] = rng;
            b_meta[at] = meta;
            b_thr[at] = throughput.x, b_thr[k_bq + at] = throughput.y, b_thr[2 * k_bq + at] = throughput.z;
            b_col[at] = color.x, b_col[k_bq + at] = color.y, b_col[2 * k_bq + at] = color.z;
            b_pxy[at] = pxy;
            if constexpr (Batch) b_tag[at] = tag;
        }
        b_count += __popc(m_goes_on);

        [[maybe_unused]] bool again = false, completes = false;  // Batch
        if (ended) {
            const uint32_t sample_index = frame.sample_index + (tag >> 4);
            const float weight = Batch ? 1.0f / static_cast<float>(sample_index + 1u) : frame.weight;
            const size_t pixel = static_cast<size_t>(pxy >> 16) * frame.width + (pxy & 0xffffu);
            accumulate_rgb1<false>(beauty, pixel, color, sample_index == 0u, weight);
            if constexpr (Batch) {
                again = (tag >> 4) + 1u < batch;  // the pixel's next sample may start now
                if (!again) completes = atomicSub(t_left + (tag & 15u), 1u) == 1u;  // the tile's last pixel has its last sample
            }
        }
        if constexpr (Batch) {
            const uint32_t m_again = __ballot_sync(0xffffffffu, again);
            if (again) {
                const uint32_t at = (p_head + p_count + __popc(m_again & lanes_below)) & (k_pq - 1u);
                p_xy[at] = pxy;
                p_tag[at] = tag + 16u;
            }
            p_count += __popc(m_again);
            uint32_t m_completes = __ballot_sync(0xffffffffu, completes);
            if (m_completes != 0u) __syncwarp();  // the tile's pixels were written by lanes of this warp, now and earlier
            while (m_completes != 0u) {
                const int from_lane = __ffs(static_cast<int>(m_completes)) - 1;
                m_completes &= m_completes - 1u;
                const uint32_t done_slot = __shfl_sync(0xffffffffu, tag, from_lane) & 15u;
                if (launch.export_slices != 0u) {  // the tile has its last sample of the launch: out it goes
                    const uint32_t done_tile = t_tile[done_slot];
                    uint32_t tile_y = __umulhi(done_tile, frame.tiles_x_magic);
                    uint32_t tile_x = done_tile - tile_y * tiles_x;
                    if (tile_x >= tiles_x) {
                        tile_x -= tiles_x;
                        tile_y++;
                    }
                    const uint32_t px = tile_x * 8u + (lane & 7u), py = tile_y * 4u + (lane >> 3);
                    const bool inside = px < frame.width && py < frame.height;
                    const uint32_t pixel = py * frame.width + px;
                    const float4* const layers[4] = {beauty, albedo, normal, depth};
                    auto slice_of = [&launch](uint32_t i) {  // slice_lo[s] <= i < slice_lo[s + 1]
                        uint32_t at = 0;
                        while (at + 1u < launch.export_slices && i >= launch.slice_lo[at + 1u]) at++;
                        return at;
                    };
                    if (inside) {
#pragma unroll
                        for (uint32_t l = 0; l < 4u; l++) {
                            const uint32_t i = l * n_px + pixel;
                            const uint32_t slice = slice_of(i);
                            static_cast<float4*>(launch.export_base[slice])[i - launch.slice_lo[slice]] = layers[l][pixel];
                        }
                    }
                    if constexpr (Reduce) {
                        // Fused reduce (batch_params): the four units of this tile have been stored at their owners;
                        // lanes 0..3 report the arrival of layer `lane` (a release on the counter: the barrier before
                        // it makes the release cover every lane's stores), and where this warp's arrival is the G-th,
                        // it sums the unit's G copies in rank order and stores the finished unit.
                        __syncwarp();
                        const uint32_t first_pixel = tile_y * 4u * frame.width + tile_x * 8u;  // always inside the image
                        uint32_t unit_slice = 0, arrivals = 0;
                        if (lane < 4u) {
                            unit_slice = slice_of(lane * n_px + first_pixel);
                            unsigned int* const counter = launch.reduce_counters[unit_slice] + lane * n_tiles + done_tile;
                            arrivals = arrive_release_system(counter) + 1u;
                            if (arrivals == launch.reduce_ranks) {
                                acquire_fence_system();  // the other ranks' copies are visible from here on
                                *counter = 0u;           // nobody else touches it before the buffer's next frame
                            }
                        }
                        uint32_t m_last = __ballot_sync(0xffffffffu, lane < 4u && arrivals == launch.reduce_ranks);
                        if (m_last != 0u) __syncwarp();  // every lane's loads below come after the fence of the lane that arrived
                        while (m_last != 0u) {
                            const uint32_t l = static_cast<uint32_t>(__ffs(static_cast<int>(m_last))) - 1u;
                            m_last &= m_last - 1u;
                            const uint32_t slice = __shfl_sync(0xffffffffu, unit_slice, static_cast<int>(l));
                            if (inside) {
                                const uint32_t within = l * n_px + pixel - launch.slice_lo[slice];
                                const float4* copy = static_cast<const float4*>(launch.reduce_staging[slice]) + within;
                                float4 sum = load_uncached(copy);
                                for (uint32_t g = 1; g < launch.reduce_ranks; g++) {
                                    const float4 v = load_uncached(copy + static_cast<size_t>(g) * launch.reduce_slot_stride);
                                    sum.x = sum.x + v.x;
                                    sum.y = sum.y + v.y;
                                    sum.z = sum.z + v.z;
                                    sum.w = sum.w + v.w;
                                }
                                const float k = launch.reduce_scale;
                                static_cast<float4*>(launch.reduce_out[slice])[within] = make_float4(sum.x * k, sum.y * k, sum.z * k, sum.w * k);
                            }
                        }
                    }
                }
                free_slots |= 1u << done_slot;
            }
        }
        __syncwarp();  // pushes and pixel writes are ordered before the next iteration's pops and reads
    }

    // The last warp to finish resets the device-wide queue for the next launch (stream-ordered).
    if (lane == 0u) {
        if (atomicAdd(tile_counters + 1, 1u) == gridDim.x * k_warps_per_block - 1u) {
            tile_counters[0] = 0u;
            tile_counters[1] = 0u;
        }
    }
    rays = __reduce_add_sync(0xffffffffu, rays);
    if (lane == 0 && rays != 0) atomicAdd(ray_counter, static_cast<unsigned long long>(rays));
}

#if !defined(QUASI_HOST_EMULATION)  // tests/cuda_emulation runs the kernel above on the host and has its own launcher
__global__ void sincos_probe_kernel(const float* __restrict__ phi, float* __restrict__ sin_out,
                                    float* __restrict__ cos_out, uint64_t n) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) sincos_round_once(phi[i], sin_out[i], cos_out[i]);
}

}  // namespace

cudaError_t launch_sincos_probe(const float* phi, float* sin_out, float* cos_out, uint64_t n,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    sincos_probe_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(phi, sin_out, cos_out, n);
    return cudaGetLastError();
}

cudaError_t launch_pathtrace_accumulate(const scene_params& scene, const frame_params& frame, const batch_params& launch,
                                       float* const layers[4], unsigned long long* ray_counter,
                                       unsigned int* tile_counters, int sm_count, cudaStream_t stream, bool generic_scan) {
    const uint32_t n_tiles = ((frame.width + 7u) / 8u) * ((frame.height + 3u) / 4u);
    uint32_t blocks = static_cast<uint32_t>(sm_count) * k_blocks_per_sm;  // resident CTAs only: a warp may wait for another
    const uint32_t needed = (n_tiles + k_warps_per_block - 1u) / k_warps_per_block;
    if (needed < blocks) blocks = needed;
    if (blocks == 0) return cudaSuccess;
    auto* b = reinterpret_cast<float4*>(layers[0]);
    auto* a = reinterpret_cast<float4*>(layers[1]);
    auto* n = reinterpret_cast<float4*>(layers[2]);
    auto* d = reinterpret_cast<float4*>(layers[3]);
    const bool batched = frame.batch_samples > 1u;
    const bool reduce = batched && launch.reduce_ranks > 1u;  // (the fused multi-GPU reduce is an instantiation of its own)
    auto run = [&](auto kernel) {
        kernel<<<blocks, k_block_threads, k_wavefront_smem, stream>>>(scene, frame, launch, b, a, n, d, ray_counter,
                                                                     tile_counters);
        return cudaGetLastError();
    };
    if (!generic_scan && cornell_layout::matches(scene)) {
        return reduce    ? run(pathtrace_accumulate_kernel<16, cornell_layout, true, true>)
               : batched ? run(pathtrace_accumulate_kernel<16, cornell_layout, true>)
                         : run(pathtrace_accumulate_kernel<16, cornell_layout, false>);
    }
    if (!generic_scan && cornell_quads_layout::matches(scene)) {
        return reduce    ? run(pathtrace_accumulate_kernel<16, cornell_quads_layout, true, true>)
               : batched ? run(pathtrace_accumulate_kernel<16, cornell_quads_layout, true>)
                         : run(pathtrace_accumulate_kernel<16, cornell_quads_layout, false>);
    }
    return reduce    ? run(pathtrace_accumulate_kernel<32, dynamic_layout, true, true>)
           : batched ? run(pathtrace_accumulate_kernel<32, dynamic_layout, true>)
                     : run(pathtrace_accumulate_kernel<32, dynamic_layout, false>);
}

}  // namespace Q::cuda
#else
}  // namespace
}  // namespace Q::cuda
#endif  // !QUASI_HOST_EMULATION
