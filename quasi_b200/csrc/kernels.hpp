/// @file kernels.hpp
/// @brief Launchers for the sm_100a kernels of libquasi_cuda.so.

#pragma once

#include "scene_params.hpp"

#include <cuda_runtime_api.h>

#include <cstdint>
#include <functional>

namespace Q::cuda {

inline constexpr uint32_t k_max_shards = 16;
inline constexpr uint32_t k_mesh_stack_entries = 32;  ///< per-lane traversal stack; BVH depth must fit
inline constexpr uint32_t k_mesh_materials = 8;

/// Device view of the mesh scene (mesh_scene.hpp:pack_for_device).
struct alignas(16) mesh_params {
    const float4* nodes;    ///< 2 float4 per node: {min.xyz, left_first} {max.xyz, count}; wide: 8 per node (wide_bvh.hpp)
    const float4* tri_v0;   ///< leaf order: {v0.xyz, material}
    const float4* tri_e1;   ///< leaf order: {e1.xyz, original index}
    const float4* tri_e2;   ///< leaf order: {e2.xyz, slot | material << k_mesh_slot_bits as bits}
    const float4* tri_shade;  ///< leaf order, 3 per triangle: mesh_scene.hpp:shading_record
    uint32_t n_nodes;
    uint32_t n_triangles;
    uint32_t max_bounces;
    uint32_t wide;          ///< 1: `nodes` is the 4-wide tree and n_nodes counts its nodes
    float materials[k_mesh_materials * 8];  ///< per material: albedo.xyz, 0, emission.xyz, 0
    const uint4* pairs;     ///< quantized_bvh.hpp: 2 uint4 per pair of children; null: the float tree only
    float grid_origin[3];   ///< of the quantized boxes
    float grid_step[3];
};

/// @brief frame.batch_samples new samples per pixel (one by default), folded into the four accumulated
/// layers in place (pathtrace.cu:pathtrace_accumulate_kernel).
/// @param launch For frame.batch_samples > 1: the zeroed per-tile epochs (required) and where finished
///        tiles are exported to (optional); ignored for a launch of one sample.
/// @param layers Device pointers to beauty, albedo, normal, depth (RGBA32F, row 0 = top).
/// @param ray_counter Device counter incremented by the number of closest-hit queries.
/// @param tile_counters Two zero-initialised device uints: the device-wide work queue (next entry, warps finished).
/// @param sm_count Multiprocessors of the device; the grid is 4 resident CTAs per SM.
/// @param generic_scan Use the any-scene kernel (one loop per quad shape) even where a kernel
///        unrolled for the scene's layout exists.
cudaError_t launch_pathtrace_accumulate(const scene_params& scene, const frame_params& frame, const batch_params& launch,
                                        float* const layers[4], unsigned long long* ray_counter,
                                        unsigned int* tile_counters, int sm_count, cudaStream_t stream,
                                        bool generic_scan = false);

/// A range of device memory whose lines a launch asks the L2 to keep (cudaAccessPropertyPersisting, launch attribute
/// cudaLaunchAttributeAccessPolicyWindow): the BVH nodes, which every ray reads twenty times, next to layers that are
/// streamed through once per launch. bytes == 0: no window.
struct l2_window {
    const void* base = nullptr;
    size_t bytes = 0;
    float hit_ratio = 1.0f;
};

/// @brief The mesh-scene form of launch_pathtrace_accumulate (BASELINE config 3).
/// @param counters Device array of 3: rays, BVH nodes visited (box tests), triangles tested.
cudaError_t launch_mesh_pathtrace(const mesh_params& scene, const frame_params& frame, float* const layers[4],
                                  unsigned long long* counters, unsigned int* tile_counters, int sm_count,
                                  cudaStream_t stream, const l2_window& keep = {});

/// Device memory of the wavefront form of the mesh scene (mesh_wavefront.cu), sized for `capacity` paths (pixels x
/// samples of a batch): two ray queues, the hits of the queue being traced, the results of every path.
inline constexpr uint32_t k_mesh_slot_bits = 28;  ///< tri_e2[slot].w = slot | material << 28: what a hit is recorded as (mesh_scene.hpp:k_slot_bits)
inline constexpr uint32_t k_mesh_wave_counters = 72;  ///< one per wave (bounce), + 1; QUASI_CUDA_MAX_BOUNCES <= 64
struct mesh_wave_buffers {
    float4* origin[2];        ///< queue 0 / 1, by position: {ray origin, path (pixel + sample * pixels; bit 31: it has a
                              ///< colour entry; all ones: no ray here) as bits}
    float4* direction[2];     ///< {unit direction, RNG state as bits}
    float4* throughput[2];    ///< {throughput, -}: written for waves >= 1 (a camera ray's is 1)
    float4* colour[2];        ///< {colour gathered so far, -}
    float2* hit;              ///< by position in the queue being traced: {t, slot in leaf order as bits (all ones: none)}
    float4* out_beauty;       ///< by path: {beauty, 1}
    float4* out_albedo_depth; ///< by path: {albedo, depth AOV value}
    float4* out_normal;       ///< by path: {normal AOV, 1}
    uint32_t* counts;         ///< [k_mesh_wave_counters] rays of wave w (in queue w & 1), then as many queue heads
    uint64_t capacity;        ///< entries of a queue (every lane of every 8x4 tile, times the samples of a batch)
};
/// Bytes of device memory per path behind a mesh_wave_buffers.
size_t mesh_wave_bytes_per_path();

/// @brief frame.batch_samples samples per pixel of the mesh scene as waves of rays (mesh_wavefront.cu), folded into
/// the four layers in place, in sample order. Enqueue only; 2 * max_bounces + 3 launches.
/// @param launches if not null: receives the number of kernels launched.
/// @param first_hit_layers_enqueued if set: called once the kernel that leaves albedo, normal and depth final has been
///        enqueued on `stream` (after the first shade pass; the beauty layer is final when the last launch ends).
cudaError_t launch_mesh_wavefront(const mesh_params& scene, const frame_params& frame, const mesh_wave_buffers& q,
                                  float* const layers[4], unsigned long long* counters, int sm_count, cudaStream_t stream,
                                  uint32_t* launches = nullptr, const std::function<void()>& first_hit_layers_enqueued = {});

/// Triangles in ORIGINAL order on the device: what build_lbvh reads.
struct lbvh_input {
    const float4* tri_v0;     ///< {v0.xyz, material}
    const float4* tri_e1;     ///< {e1.xyz, -}
    const float4* tri_e2;     ///< {e2.xyz, -}
    const float4* tri_shade;  ///< 3 per triangle (mesh_scene.hpp:shading_record)
    uint32_t n_triangles;
};
/// Device buffers build_lbvh fills: the layout mesh_params describes.
struct lbvh_output {
    float4* nodes;         ///< 2 * (2 n - 1) float4
    float4* tri_v0;        ///< n, leaf order
    float4* tri_e1;        ///< n, w = original index
    float4* tri_e2;        ///< n
    float4* tri_shade;     ///< 3 n
    uint32_t* leaf_order;  ///< n: slot -> original index
    uint32_t* depth;       ///< 1: longest root-to-leaf path in nodes
};
/// @brief Builds a linear BVH over the triangles on the GPU (lbvh_build.cu) and waits for it.
/// @param build_ms If not null: device time of the build's kernels (CUDA events; the
///        allocation of the scratch space is not included).
cudaError_t build_lbvh(const lbvh_input& in, const lbvh_output& out, cudaStream_t stream, float* build_ms = nullptr);

/// @brief Test hook: the kernel's sin/cos of phi in [0, 2*pi], element-wise.
cudaError_t launch_sincos_probe(const float* phi, float* sin_out, float* cos_out, uint64_t n,
                                cudaStream_t stream);

/// @brief output = mix(history, current, 1/(frame_count+1)) over n_pixels float4s.
cudaError_t launch_accumulate(const float* current, const float* history, float* output,
                              uint64_t n_pixels, uint32_t frame_count, cudaStream_t stream);

/// @brief Reinhard + gamma 2.2 into BGRA8 unorm (row 0 = top).
cudaError_t launch_tonemap_bgra8(const float* accumulated, uint32_t width, uint32_t height,
                                 void* bgra, uint32_t pitch_bytes, cudaStream_t stream);

/// @brief out = (shards[0] + shards[1] + ...) * scale, in that order (scale 0 = 1/n);
/// optional fused tonemap.
cudaError_t launch_combine_shards(const float* const* shards, uint32_t n_shards,
                                  uint64_t n_pixels, float scale, float* out, void* bgra,
                                  cudaStream_t stream);

/// Words in a rank's flag block for launch_reduce_slices_sync: arrive[16] (the rest is spare).
inline constexpr uint32_t k_sync_words = 2 * k_max_shards + 1;

/// @brief The layer block slice by slice: float4 i of slice s (slice_lo[s] <= i < slice_lo[s + 1]) to
/// base[s] + (i - slice_lo[s]) (the addressing of batch_params' export), for a frame whose last launch exported nothing.
cudaError_t launch_push_slices(const float* block, uint64_t n_float4, void* const* base, const uint32_t* slice_lo,
                               uint32_t slices, cudaStream_t stream);

/// @brief The owner's half of the multi-GPU exchange (post.cu:reduce_slices_sync_kernel): waits on device-side
/// flags until all `n_shards` ranks have completed the frame `epoch`, then out[i] = (sum over g in rank order
/// of staging[g * float4_per_slice + i]) * scale for i < n_float4 (scale 0 = 1/n_shards).
/// `flags[g]`: rank g's zero-initialised flag block (k_sync_words uint32), addressable from this GPU.
cudaError_t launch_reduce_slices_sync(const float* staging, uint32_t float4_per_slice, uint32_t* const* flags,
                                      uint32_t n_shards, uint32_t rank, uint32_t epoch, uint64_t n_float4, float scale,
                                      float* out, cudaStream_t stream);

}  // namespace Q::cuda
