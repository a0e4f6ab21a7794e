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
///  - Persistent warps with path regeneration. The grid is resident CTAs only
///    (4 per SM); warps take 8x4 pixel tiles from a device-wide work queue (one
///    atomic per tile, device_common.cuh:tile_feed). A lane whose
///    path ended writes its pixel and immediately starts the warp's next pixel
///    (ballot + popc compaction of the idle lanes), so every lane of the warp is
///    always inside the same bounce step. In the thread-per-pixel form the warp
///    ran at 18 of 32 lanes (ncu, profiles/r1_v1_*).
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

/// Batched launches only (QUASI_CUDA_MIRROR, plugin.cu): pinned host buffers, one per layer, that
/// receive every pixel's value after the LAST sample of the launch, written by the lane that
/// computed it, so that the device-to-host transfer of a frame is spread over the launch that
/// renders it instead of following it. Null entries: nothing is mirrored.
__device__ float4* g_mirror[4];

__device__ __forceinline__ void mirror_pixel(int layer, size_t pixel, float4 value) {
    float4* to = g_mirror[layer];
    if (to != nullptr) to[pixel] = value;
}

/// Sample-major batched launches only: one word per tile, the number of samples of THIS launch
/// that every pixel of the tile has received (zeroed by the host before the launch).
__device__ unsigned int* g_tile_epochs;

__device__ __forceinline__ unsigned int load_acquire_gpu(const unsigned int* p) {
#if defined(QUASI_HOST_EMULATION)
    return std::atomic_ref<const unsigned int>(*p).load(std::memory_order_acquire);
#else
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void store_release_gpu(unsigned int* p, unsigned int v) {
#if defined(QUASI_HOST_EMULATION)
    std::atomic_ref<unsigned int>(*p).store(v, std::memory_order_release);
#else
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

/// The work queue of a sample-major launch: entry q of n_tiles * k is tile q % n_tiles at sample
/// q / n_tiles, so all tiles get sample 0, then all get sample 1, ...; otherwise tile_feed
/// (device_common.cuh). An entry with sample s > 0 is READY once the tile's epoch says sample
/// s - 1 is complete. Entries are handed out in order, so that sample was handed out n_tiles
/// entries ago and is all but always complete; when it is not, the warp hands out no pixel of the
/// entry and looks again on its next iteration, never spinning while it holds paths itself (on a
/// small image the unfinished sample may be in this very warp). It is what lets ONE launch hold k
/// samples per pixel without a lane being tied to a pixel: the launch ends when the last paths of
/// the last sample end, the same few iterations that end a one-sample launch, once instead of k times.
struct sample_major_feed {
    static constexpr uint32_t k_unfetched = 0xffffffffu;
    unsigned int* counters;
    uint32_t n_tiles, n_entries;
    uint32_t slot0 = 0;
    uint32_t entry0 = k_unfetched, entry1 = k_unfetched;  ///< queue entries of positions slot0, slot0 + 1
    uint32_t tile0 = 0, tile1 = 0, sample0 = 0, sample1 = 0;
    bool ready0 = false, ready1 = false;
    bool dry = false;

    /// Warp-uniform (every lane looks, the warp takes the answer all lanes agree on: the feed's
    /// state must be the same in every lane).
    __device__ __forceinline__ bool is_ready(uint32_t tile, uint32_t sample) const {
        if (sample == 0u) return true;
        return __ballot_sync(0xffffffffu, load_acquire_gpu(g_tile_epochs + tile) >= sample) == 0xffffffffu;
    }
    __device__ __forceinline__ void fetch(uint32_t& entry, uint32_t& tile, uint32_t& sample, bool& ready) {
        uint32_t q = 0;
        if ((threadIdx.x & 31u) == 0u) q = atomicAdd(counters, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        dry = q >= n_entries;
        entry = q;
        ready = false;
        if (dry) return;
        sample = q / n_tiles;
        tile = q - sample * n_tiles;
        ready = is_ready(tile, sample);
    }
    /// Warp-uniform: makes the entries of items [first, last] available and looks again at the
    /// epochs of those that were not ready.
    __device__ __forceinline__ void open(uint32_t first, uint32_t last) {
        if (slot0 < (first >> 5)) {
            slot0 = first >> 5;
            entry0 = entry1, tile0 = tile1, sample0 = sample1, ready0 = ready1;
            entry1 = k_unfetched;
            ready1 = false;
        }
        if (entry0 == k_unfetched && !dry) fetch(entry0, tile0, sample0, ready0);
        if (entry0 < n_entries && !ready0) ready0 = is_ready(tile0, sample0);
        if (!dry && (last >> 5) > slot0 && entry1 == k_unfetched) fetch(entry1, tile1, sample1, ready1);
        if (entry1 < n_entries && !ready1) ready1 = is_ready(tile1, sample1);
    }
    __device__ __forceinline__ bool has(uint32_t item) const { return ((item >> 5) == slot0 ? entry0 : entry1) < n_entries; }
    __device__ __forceinline__ uint32_t tile_of(uint32_t item) const { return (item >> 5) == slot0 ? tile0 : tile1; }
    __device__ __forceinline__ uint32_t sample_of(uint32_t item) const { return (item >> 5) == slot0 ? sample0 : sample1; }
    /// How many of the `wanted` items starting at `first` may be handed out now: up to the first
    /// one that belongs to an entry that is not ready (an entry past the end of the queue counts as
    /// ready: its items are skipped).
    __device__ __forceinline__ uint32_t grantable(uint32_t first, uint32_t wanted) const {
        const bool ok0 = ready0 || entry0 >= n_entries;
        const bool ok1 = ready1 || entry1 >= n_entries;
        if (!ok0) return 0u;
        const uint32_t in_entry0 = 32u - (first & 31u);
        return (wanted <= in_entry0 || ok1) ? wanted : in_entry0;
    }
    /// Warp-uniform: the queue has run out at item `next` (false while an entry is only not ready).
    __device__ __forceinline__ bool exhausted(uint32_t next) {
        open(next, next);
        return dry && entry0 >= n_entries;  // (an entry never fetched because the queue had run out counts as past its end)
    }
    __device__ __forceinline__ void finish(uint32_t n_warps) {
        if ((threadIdx.x & 31u) == 0u) {
            if (atomicAdd(counters + 1, 1u) == n_warps - 1u) {
                counters[0] = 0u;
                counters[1] = 0u;
            }
        }
    }
};

/// `Batch` (QUASI_CUDA_BATCH, plugin.cu): one launch gives every pixel frame.batch_samples
/// consecutive samples instead of one. A lane that ends a path starts the next sample of the
/// SAME pixel, so the samples of a pixel are accumulated in order by one thread (the running
/// mean of plugin.mm:326 depends on the order) and its 64 bytes of history stay in L1/L2
/// between them; what lies between two launches is paid once per batch instead of once per
/// sample (the drain inside a launch scales with the item, which is k samples long now).
/// `SampleMajor` (with Batch): the k samples are k passes over the image inside the one launch
/// (sample_major_feed) instead of k consecutive samples per lane and pixel.
template <int NQ, class Layout, bool Batch = false, bool SampleMajor = false>
__global__ void __launch_bounds__(k_block_threads, k_blocks_per_sm)
pathtrace_accumulate_kernel(const __grid_constant__ scene_params scene,
                            const __grid_constant__ frame_params frame, float4* __restrict__ beauty,
                            float4* __restrict__ albedo, float4* __restrict__ normal,
                            float4* __restrict__ depth, unsigned long long* __restrict__ ray_counter,
                            unsigned int* __restrict__ tile_counters) {
    pdl_launch_dependents();
    __shared__ float s_shade[NQ * k_shade_stride];
    for (int idx = threadIdx.x; idx < NQ * static_cast<int>(k_shade_floats); idx += k_block_threads) {
        const int qi = idx / static_cast<int>(k_shade_floats);
        const int k = idx - qi * static_cast<int>(k_shade_floats);
        s_shade[qi * k_shade_stride + k] =
            reinterpret_cast<const float*>(&scene.quads[qi])[k_shade_offset + k];
    }
    __syncthreads();
    pdl_wait();  // layers, tile queue and counters belong to the previous launch until here

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lanes_below = (1u << lane) - 1u;
    const uint32_t tiles_x = (frame.width + 7u) >> 3;
    const uint32_t n_tiles = tiles_x * ((frame.height + 3u) >> 2);
    const bool first_sample = frame.sample_index == 0;
    const float weight = frame.weight;
    const uint32_t max_bounces = scene.max_bounces;
    const v3 cam_origin = load_v3(frame.position);

    // Path state of this lane.
    bool active = false;
    v3 ro = cam_origin, rd = make_v3(0.0f, 0.0f, -1.0f);
    v3 throughput = make_v3(1.0f, 1.0f, 1.0f), color = make_v3(0.0f, 0.0f, 0.0f);
    uint32_t rng = 0, bounce = 0;
    size_t pixel = 0;
    uint32_t rays = 0;
    uint32_t next_item = 0;  // warp-uniform cursor into this warp's pixel sequence
    tile_feed feed{tile_counters, n_tiles};
    // Where a continuing path scatters from: the quad it hit, complemented when it was hit
    // from behind. Set by the shading step, consumed at the top of the next iteration.
    int scatter = 0;
    // Batch only: which sample of the launch this lane's path is, its pixel, and that sample's
    // accumulation constants (first_sample / weight are per launch otherwise).
    [[maybe_unused]] uint32_t sample_of_item = 0, pixel_xy = 0xffffffffu;
    [[maybe_unused]] bool lane_first = first_sample;
    [[maybe_unused]] float lane_weight = weight;
    [[maybe_unused]] const uint32_t batch = frame.batch_samples > 1u ? frame.batch_samples : 1u;
    // SampleMajor only: the queue, this warp's eight completion counters (dynamic shared memory,
    // zeroed here), and per lane tile << 9 | pixels of the tile inside the image << 3 | counter
    [[maybe_unused]] sample_major_feed sfeed{tile_counters, n_tiles, n_tiles * batch};
    [[maybe_unused]] unsigned int* tile_done = nullptr;
    [[maybe_unused]] uint32_t item_tag = 0;
    if constexpr (SampleMajor) {
#if defined(QUASI_HOST_EMULATION)
        unsigned int* s_done = static_cast<unsigned int*>(emu::dynamic_shared());
#else
        extern __shared__ unsigned int s_done[];
#endif
        tile_done = s_done + (threadIdx.x >> 5) * 8u;
        if (lane < 8u) tile_done[lane] = 0u;
        __syncthreads();
    }

    while (true) {
        // Regeneration: idle lanes take the warp's next pixels, compacted.
        bool fresh = false;
        float uv_x = 0.0f, uv_y = 0.0f;
        if constexpr (Batch && SampleMajor) {
            bool start = false;
            uint32_t px = 0, py = 0;
            const uint32_t idle = __ballot_sync(0xffffffffu, !active);
            if (idle != 0u) {
                sfeed.open(next_item, next_item + __popc(idle) - 1u);
                const uint32_t granted = sfeed.grantable(next_item, __popc(idle));
                if (!active) {
                    const uint32_t item = next_item + __popc(idle & lanes_below);
                    if (item < next_item + granted && sfeed.has(item)) {
                        const uint32_t tile = sfeed.tile_of(item);
                        uint32_t tile_y = __umulhi(tile, frame.tiles_x_magic);
                        uint32_t tile_x = tile - tile_y * tiles_x;
                        if (tile_x >= tiles_x) {
                            tile_x -= tiles_x;
                            tile_y++;
                        }
                        px = tile_x * 8u + (item & 7u);
                        py = tile_y * 4u + ((item >> 3) & 3u);
                        if (px < frame.width && py < frame.height) {
                            const uint32_t in_x = frame.width - tile_x * 8u, in_y = frame.height - tile_y * 4u;
                            const uint32_t valid = (in_x < 8u ? in_x : 8u) * (in_y < 4u ? in_y : 4u);
                            sample_of_item = sfeed.sample_of(item);
                            item_tag = (tile << 9) | (valid << 3) | ((item >> 5) & 7u);
                            start = true;
                            pixel = static_cast<size_t>(py) * frame.width + px;
                            if (!(first_sample && sample_of_item == 0u)) {
                                prefetch_l2(beauty + pixel);
                                prefetch_l2(albedo + pixel);
                                prefetch_l2(normal + pixel);
                                prefetch_l2(depth + pixel);
                            }
                        }
                    }
                }
                next_item += granted;
            }
            if (start) {
                const uint32_t sample_index = frame.sample_index + sample_of_item;
                uv_x = (static_cast<float>(px) + 0.5f) / static_cast<float>(frame.width);
                uv_y = 1.0f - (static_cast<float>(py) + 0.5f) / static_cast<float>(frame.height);
                rng = pcg_hash(px + py * 1920u + (frame.frame_index + sample_of_item * frame.frame_stride) * 1920u * 1080u);
                fresh = true;
                throughput = make_v3(1.0f, 1.0f, 1.0f);
                color = make_v3(0.0f, 0.0f, 0.0f);
                bounce = 0;
                active = true;
                lane_first = sample_index == 0u;
                lane_weight = 1.0f / static_cast<float>(sample_index + 1u);
            }
        } else if constexpr (Batch) {
            // A lane whose path has ended goes on with the next sample of its pixel; only a lane
            // that has given its pixel all its samples takes a new pixel from the queue.
            bool start = false;
            uint32_t px = 0, py = 0;
            if (!active && pixel_xy != 0xffffffffu) {
                if (sample_of_item + 1u < batch) {
                    sample_of_item++;
                    px = pixel_xy & 0xffffu;
                    py = pixel_xy >> 16;
                    start = true;
                } else {
                    pixel_xy = 0xffffffffu;
                }
            }
            const uint32_t idle = __ballot_sync(0xffffffffu, !active && !start);
            if (idle != 0u) {
                feed.open(next_item, next_item + __popc(idle) - 1u);
                if (!active && !start) {
                    const uint32_t item = next_item + __popc(idle & lanes_below);
                    const uint32_t tile = feed.tile_of(item);
                    if (tile < n_tiles) {
                        uint32_t tile_y = __umulhi(tile, frame.tiles_x_magic);
                        uint32_t tile_x = tile - tile_y * tiles_x;
                        if (tile_x >= tiles_x) {
                            tile_x -= tiles_x;
                            tile_y++;
                        }
                        px = tile_x * 8u + (item & 7u);
                        py = tile_y * 4u + ((item >> 3) & 3u);
                        if (px < frame.width && py < frame.height) {
                            pixel_xy = px | (py << 16);  // the ABI's sizes are far below 65536
                            sample_of_item = 0;
                            start = true;
                            pixel = static_cast<size_t>(py) * frame.width + px;
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
            if (start) {
                const uint32_t sample_index = frame.sample_index + sample_of_item;
                uv_x = (static_cast<float>(px) + 0.5f) / static_cast<float>(frame.width);
                uv_y = 1.0f - (static_cast<float>(py) + 0.5f) / static_cast<float>(frame.height);
                rng = pcg_hash(px + py * 1920u + (frame.frame_index + sample_of_item * frame.frame_stride) * 1920u * 1080u);
                fresh = true;
                throughput = make_v3(1.0f, 1.0f, 1.0f);
                color = make_v3(0.0f, 0.0f, 0.0f);
                bounce = 0;
                active = true;
                lane_first = sample_index == 0u;
                lane_weight = 1.0f / static_cast<float>(sample_index + 1u);  // scene_params.hpp:prepare_frame
            }
        } else {
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u) {
            feed.open(next_item, next_item + __popc(idle) - 1u);
            if (!active) {
                const uint32_t item = next_item + __popc(idle & lanes_below);
                const uint32_t tile = feed.tile_of(item);
                if (tile < n_tiles) {
                    // tile / tiles_x by multiply-high with floor(2^32 / tiles_x): at most
                    // one too small for any 32-bit tile, fixed up by one compare.
                    uint32_t tile_y = __umulhi(tile, frame.tiles_x_magic);
                    uint32_t tile_x = tile - tile_y * tiles_x;
                    if (tile_x >= tiles_x) {
                        tile_x -= tiles_x;
                        tile_y++;
                    }
                    const uint32_t px = tile_x * 8u + (item & 7u);
                    const uint32_t py = tile_y * 4u + ((item >> 3) & 3u);
                    if (px < frame.width && py < frame.height) {
                        // pathtrace_frag, plugin.mm:287-309; uv is the closed form of
                        // the rasteriser's interpolation, py = 0 on top.
                        uv_x = (static_cast<float>(px) + 0.5f) / static_cast<float>(frame.width);
                        uv_y = 1.0f - (static_cast<float>(py) + 0.5f) / static_cast<float>(frame.height);
                        rng = pcg_hash(px + py * 1920u + frame.frame_index * 1920u * 1080u);
                        fresh = true;
                        throughput = make_v3(1.0f, 1.0f, 1.0f);
                        color = make_v3(0.0f, 0.0f, 0.0f);
                        bounce = 0;
                        pixel = static_cast<size_t>(py) * frame.width + px;
                        active = true;
                        // The history of this pixel is needed ~1000 cycles from now
                        // (after the first trace); start pulling it out of HBM.
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
        }  // !Batch
        if (!__any_sync(0xffffffffu, active)) {
            if constexpr (SampleMajor) {
                if (sfeed.exhausted(next_item)) break;
            } else {
                if (feed.exhausted(next_item)) break;
            }
            continue;  // only off-image pixels of an edge tile were handed out
        }

        if (active) {
            // Every path needs a direction now: a new one its camera ray (get_camera_ray,
            // plugin.mm:129-153, jitter x drawn before y), a continuing one its scattered ray
            // (cosine_sample_hemisphere, plugin.mm:111-125, r1 before r2, with the per-quad
            // frame). Both take two draws and end in a normalisation; those are issued once
            // for the whole warp instead of once per kind of lane.
            const float draw_a = random_float(rng);
            const float draw_b = random_float(rng);
            v3 d;
            if (fresh) {
                uv_x += (draw_a - 0.5f) * 0.001f;
                uv_y += (draw_b - 0.5f) * 0.001f;
                const v3 target = load_v3(frame.lower_left) + load_v3(frame.horizontal) * uv_x +
                                  load_v3(frame.vertical) * uv_y;
                ro = cam_origin;
                d = target - cam_origin;
            } else {
                const float* from = s_shade + (scatter < 0 ? ~scatter : scatter) * k_shade_stride;
                const float side = scatter < 0 ? -1.0f : 1.0f;
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

            // One bounce of path_trace, plugin.mm:232-273.
            hit_record hit;
            trace_scene<Layout>(scene, s_shade, ro, rd, hit);
            rays++;
            const bool is_hit = hit.index >= 0;
            const float* sh = s_shade + (is_hit ? hit.index : 0) * k_shade_stride;
            const float side = (is_hit && !hit.front) ? -1.0f : 1.0f;
            const v3 n = load_v3(sh + k_sh_normal) * side;  // exact sign flip
            const v3 q_emission = load_v3(sh + k_sh_emission);
            const v3 q_albedo = load_v3(sh + k_sh_albedo);
            const bool is_light = sh[k_sh_is_light] != 0.0f;

            if (bounce == 0) {
                // First-hit AOVs (plugin.mm:241-249, 300-307) go out now, so the
                // lane does not carry them to the end of the path.
                v3 a = make_v3(0.0f, 0.0f, 0.0f), nn = a;
                float first_t = 0.0f;
                if (is_hit) {
                    a = is_light ? q_emission : q_albedo;
                    nn = make_v3(n.x * 0.5f + 0.5f, n.y * 0.5f + 0.5f, n.z * 0.5f + 0.5f);
                    first_t = hit.t;
                }
                const float nd = fminf(fmaxf((first_t - 1.0f) / 4.0f, 0.0f), 1.0f);
                if constexpr (Batch) {
                    constexpr auto accumulate = SampleMajor ? accumulate_pixel_value_cg : accumulate_pixel_value;
                    const float4 va = accumulate(albedo, pixel, make_float4(a.x, a.y, a.z, 1.0f), lane_first, lane_weight);
                    const float4 vn = accumulate(normal, pixel, make_float4(nn.x, nn.y, nn.z, 1.0f), lane_first, lane_weight);
                    const float4 vd = accumulate(depth, pixel, make_float4(nd, nd, nd, 1.0f), lane_first, lane_weight);
                    if (sample_of_item + 1u == batch) {  // the pixel's last sample of this launch
                        mirror_pixel(1, pixel, va);
                        mirror_pixel(2, pixel, vn);
                        mirror_pixel(3, pixel, vd);
                    }
                } else {
                accumulate_pixel(albedo, pixel, make_float4(a.x, a.y, a.z, 1.0f), first_sample, weight);
                accumulate_pixel(normal, pixel, make_float4(nn.x, nn.y, nn.z, 1.0f), first_sample, weight);
                accumulate_pixel(depth, pixel, make_float4(nd, nd, nd, 1.0f), first_sample, weight);
                }
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

            if (!done) {
                ro = hit.point + n * 0.001f;
                scatter = hit.front ? hit.index : ~hit.index;  // the direction is drawn at the top
            } else {
                if constexpr (Batch) {
                    constexpr auto accumulate = SampleMajor ? accumulate_pixel_value_cg : accumulate_pixel_value;
                    const float4 vb = accumulate(beauty, pixel, make_float4(color.x, color.y, color.z, 1.0f), lane_first, lane_weight);
                    if (sample_of_item + 1u == batch) mirror_pixel(0, pixel, vb);
                    if constexpr (SampleMajor) {
                        // this pixel has its sample; the lane that completes the tile publishes the tile's epoch
                        __threadfence_block();
                        const uint32_t counter = item_tag & 7u, valid = (item_tag >> 3) & 63u;
                        if (atomicAdd(tile_done + counter, 1u) + 1u == valid) {
                            tile_done[counter] = 0u;
                            __threadfence();
                            store_release_gpu(g_tile_epochs + (item_tag >> 9), sample_of_item + 1u);
                        }
                    }
                } else {
                accumulate_pixel(beauty, pixel, make_float4(color.x, color.y, color.z, 1.0f),
                                 first_sample, weight);
                }
                active = false;
            }
        }
    }

    if constexpr (SampleMajor) {
        sfeed.finish(gridDim.x * k_warps_per_block);
    } else {
        feed.finish(gridDim.x * k_warps_per_block);
    }
    // One atomic per warp for the ray statistic (SURVEY.md 8d "Ray counting").
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

cudaError_t set_pathtrace_mirror(float* const mirror[4], cudaStream_t stream) {
    float4* table[4];
    for (int i = 0; i < 4; i++) table[i] = mirror ? reinterpret_cast<float4*>(mirror[i]) : nullptr;
    return cudaMemcpyToSymbolAsync(g_mirror, table, sizeof table, 0, cudaMemcpyHostToDevice, stream);
}

cudaError_t set_pathtrace_tile_epochs(unsigned int* epochs, cudaStream_t stream) {
    return cudaMemcpyToSymbolAsync(g_tile_epochs, &epochs, sizeof epochs, 0, cudaMemcpyHostToDevice, stream);
}

cudaError_t launch_pathtrace_accumulate(const scene_params& scene, const frame_params& frame,
                                        float* const layers[4], unsigned long long* ray_counter,
                                        unsigned int* tile_counters, int sm_count, cudaStream_t stream,
                                        bool generic_scan, bool sample_major) {
    const uint32_t n_tiles = ((frame.width + 7u) / 8u) * ((frame.height + 3u) / 4u);
    uint32_t blocks = static_cast<uint32_t>(sm_count) * k_blocks_per_sm;
    const uint32_t needed = (n_tiles + k_warps_per_block - 1u) / k_warps_per_block;
    if (needed < blocks) blocks = needed;
    if (blocks == 0) return cudaSuccess;
    auto* b = reinterpret_cast<float4*>(layers[0]);
    auto* a = reinterpret_cast<float4*>(layers[1]);
    auto* n = reinterpret_cast<float4*>(layers[2]);
    auto* d = reinterpret_cast<float4*>(layers[3]);
#if defined(QUASI_PDL) && QUASI_PDL
    if (!generic_scan && cornell_layout::matches(scene)) {
        return launch_persistent(pathtrace_accumulate_kernel<16, cornell_layout>, blocks, k_block_threads, 0, stream, scene,
                                 frame, b, a, n, d, ray_counter, tile_counters);
    }
    if (!generic_scan && cornell_quads_layout::matches(scene)) {
        return launch_persistent(pathtrace_accumulate_kernel<16, cornell_quads_layout>, blocks, k_block_threads, 0, stream,
                                 scene, frame, b, a, n, d, ray_counter, tile_counters);
    }
    return launch_persistent(pathtrace_accumulate_kernel<32, dynamic_layout>, blocks, k_block_threads, 0, stream, scene, frame,
                             b, a, n, d, ray_counter, tile_counters);
#endif
    if (frame.batch_samples > 1u && sample_major) {  // QUASI_CUDA_BATCH_ORDER=sample: k passes over the image in one launch
        constexpr size_t smem = k_warps_per_block * 8 * sizeof(unsigned int);  // completion counters, eight per warp
        if (!generic_scan && cornell_layout::matches(scene)) {
            pathtrace_accumulate_kernel<16, cornell_layout, true, true>
                <<<blocks, k_block_threads, smem, stream>>>(scene, frame, b, a, n, d, ray_counter, tile_counters);
        } else if (!generic_scan && cornell_quads_layout::matches(scene)) {
            pathtrace_accumulate_kernel<16, cornell_quads_layout, true, true>
                <<<blocks, k_block_threads, smem, stream>>>(scene, frame, b, a, n, d, ray_counter, tile_counters);
        } else {
            pathtrace_accumulate_kernel<32, dynamic_layout, true, true>
                <<<blocks, k_block_threads, smem, stream>>>(scene, frame, b, a, n, d, ray_counter, tile_counters);
        }
        return cudaGetLastError();
    }
    if (frame.batch_samples > 1u) {  // QUASI_CUDA_BATCH: several samples per pixel in one launch
        if (!generic_scan && cornell_layout::matches(scene)) {
            pathtrace_accumulate_kernel<16, cornell_layout, true>
                <<<blocks, k_block_threads, 0, stream>>>(scene, frame, b, a, n, d, ray_counter, tile_counters);
        } else if (!generic_scan && cornell_quads_layout::matches(scene)) {
            pathtrace_accumulate_kernel<16, cornell_quads_layout, true>
                <<<blocks, k_block_threads, 0, stream>>>(scene, frame, b, a, n, d, ray_counter, tile_counters);
        } else {
            pathtrace_accumulate_kernel<32, dynamic_layout, true>
                <<<blocks, k_block_threads, 0, stream>>>(scene, frame, b, a, n, d, ray_counter, tile_counters);
        }
        return cudaGetLastError();
    }
    if (!generic_scan && cornell_layout::matches(scene)) {
        pathtrace_accumulate_kernel<16, cornell_layout>
            <<<blocks, k_block_threads, 0, stream>>>(scene, frame, b, a, n, d, ray_counter, tile_counters);
    } else if (!generic_scan && cornell_quads_layout::matches(scene)) {
        pathtrace_accumulate_kernel<16, cornell_quads_layout>
            <<<blocks, k_block_threads, 0, stream>>>(scene, frame, b, a, n, d, ray_counter, tile_counters);
    } else {
        pathtrace_accumulate_kernel<32, dynamic_layout>
            <<<blocks, k_block_threads, 0, stream>>>(scene, frame, b, a, n, d, ray_counter, tile_counters);
    }
    return cudaGetLastError();
}

}  // namespace Q::cuda
#else
}  // namespace
}  // namespace Q::cuda
#endif  // !QUASI_HOST_EMULATION
