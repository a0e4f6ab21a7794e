/// @file plugin.cu
/// @brief libquasi_cuda.so: the CUDA backend behind Quasi's plugin C ABI.
///
/// Counterpart of the Metal backend's host half
/// (/root/reference/backends/metal/plugin.mm:389-904). Same ten entry points, same
/// null-tolerance, ownership and per-call semantics; the
/// GPU work is the kernels in pathtrace.cu and post.cu. There is no CPU path: if
/// CUDA fails, create returns NULL and render/readback do nothing.

#include "kernels.hpp"
#include "mesh_scene.hpp"
#include "quantized_bvh.hpp"
#include "wide_bvh.hpp"

#include <quasi_cuda_backend.h>

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace {

static_assert(Q::cuda::k_mesh_slot_bits == Q::cuda::mesh::k_slot_bits, "one definition of a recorded hit");

using Q::cuda::frame_params;
using Q::cuda::scene_params;

constexpr const char* k_name = "CUDA Path Tracer";
constexpr const char* k_description = "Cornell Box path tracer using CUDA (sm_100a)";
constexpr const char* k_author = "Quasi";

/// Pinned host buffers handed out by readback. The ABI's free functions take no
/// plugin handle (plugin_interface.hpp:140, 149), so the pool is per library
/// image. Buffers are recycled: cudaHostAlloc of 33 MB costs milliseconds,
/// reuse costs nothing, and pinned memory lets the D2H copy run at PCIe speed.
class readback_pool {
public:
    float* acquire(size_t bytes) {
        std::lock_guard<std::mutex> lock(mutex_);
        for (auto& e : entries_) {
            if (!e.in_use && e.bytes == bytes) {
                e.in_use = true;
                return e.ptr;
            }
        }
        void* p = nullptr;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        entries_.push_back({static_cast<float*>(p), bytes, true});
        return static_cast<float*>(p);
    }

    void release(float* ptr) {
        std::lock_guard<std::mutex> lock(mutex_);
        for (auto& e : entries_) {
            if (e.ptr == ptr) {
                e.in_use = false;
                return;
            }
        }
    }

    /// Frees idle buffers; called when the last plugin instance goes away.
    void trim() {
        std::lock_guard<std::mutex> lock(mutex_);
        std::vector<entry> keep;
        for (auto& e : entries_) {
            if (e.in_use) {
                keep.push_back(e);
            } else {
                cudaFreeHost(e.ptr);
            }
        }
        entries_.swap(keep);
    }

private:
    struct entry {
        float* ptr;
        size_t bytes;
        bool in_use;
    };
    std::mutex mutex_;
    std::vector<entry> entries_;
};

readback_pool g_pool;
std::mutex g_instances_mutex;
int g_instances = 0;

struct plugin_state {
    Q_plugin_context* context = nullptr;
    int device = 0;
    int sm_count = 148;
    uint32_t shard_index = 0;
    uint32_t shard_count = 1;
    cudaStream_t queue = nullptr;

    Q::cuda::scene_desc scene_desc;
    scene_params scene;

    // Mesh scene (QUASI_CUDA_SCENE=mesh, BASELINE config 3); empty for the demo scene.
    bool use_mesh = false;
    Q::cuda::mesh::scene mesh;
    Q::cuda::mesh_params mesh_device{};
    bool mesh_wavefront = false;     ///< headless calls of the mesh scene are collected and traced as waves (mesh_wavefront.cu)
    uint32_t mesh_wave_min = 4;      ///< fewer collected calls than this go through the one-kernel form, one launch each
    void* wave_block = nullptr;      ///< the one allocation behind `wave`
    Q::cuda::mesh_wave_buffers wave{};
    Q::cuda::l2_window mesh_keep{};  ///< QUASI_CUDA_MESH_L2_KEEP: the part of the mesh the launches ask the L2 to keep
    void* mesh_buffers[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  ///< [5]: the 4-wide tree, if asked for; [6]: the quantized pairs
    uint32_t wide_nodes = 0;

    float* layers[Q_AOV_COUNT] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long* ray_counter = nullptr;
    cudaEvent_t last_render = nullptr;
    bool render_pending = false;

    uint32_t frame_count = 0;
    uint32_t last_width = 0;
    uint32_t last_height = 0;

    uint64_t paths = 0;
    uint64_t kernel_launches = 0;

    bool generic_scan = false;  ///< QUASI_CUDA_SCAN=generic: the any-scene kernel

    // QUASI_CUDA_BATCH=k (default 64; demo scene, headless calls): render calls are collected and every
    // k of them become ONE launch that gives each pixel k samples, a pixel's samples one after the other in
    // the warp that took its tile (pathtrace.cu, Batch). Anything that looks at the layers, and any call with another
    // camera or size, launches what has been collected first, so the only observable difference is time.
    uint32_t batch = 1;
    // QUASI_CUDA_MIRROR (default on, with a batch): a batched launch exports every tile that has received its
    // last sample into pinned host buffers (batch_params), so a readback that follows finds the frame already
    // on the host: the 64 B/px device-to-host transfer is spread over the launch instead of following it.
    bool mirror_enabled = false;
    bool mirror_valid = false;         ///< the buffers hold what the layers hold
    // Mesh scene (wavefront form): the mirror is filled by copies on a stream of its own, the three first-hit layers
    // while the later waves are still being traced, the beauty layer after the last launch; `mirror_copied` follows
    // the last of them, and a readback that uses the mirror waits for it as well as for the render.
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t layers_ready = nullptr, mirror_copied = nullptr;
    bool mirror_by_copy = false;
    float* mirror[Q_AOV_COUNT] = {nullptr, nullptr, nullptr, nullptr};
    // Q_cuda_plugin_set_export: a multi-GPU host has the batched launches push every slice of the layer
    // block to the GPU that owns it (peer memory) instead; takes the place of the mirror while set.
    bool external_export = false;
    Q_cuda_export_table export_table{};
    bool external_reduce = false;  ///< Q_cuda_plugin_set_reduce: the export with the multi-GPU reduce fused in
    Q_cuda_reduce_table reduce_table{};
    uint32_t reduce_calls_seen = 0;  ///< render calls since Q_cuda_plugin_set_reduce (reduce_table.render_calls)
    bool last_launch_exported = false;  ///< the most recent launch sent every tile of the layers to the export target
    uint32_t pending = 0;              ///< render calls collected and not yet launched
    uint32_t pending_first_sample = 0; ///< frame_count of the first of them
    Q_render_frame pending_frame{};    ///< that call's frame (camera, size, stream)
    Q_cuda_persistent_state* persistent = nullptr;  ///< host-owned, Q_gpu_context::layer (may be null)
    uint32_t scene_tag = 0;
};

void log_msg(const plugin_state* state, const char* msg) {
    if (state->context && state->context->log) {
        state->context->log(state->context->host_data, msg);
    }
}

bool check(const plugin_state* state, cudaError_t err, const char* what) {
    if (err == cudaSuccess) return true;
    char buf[256];
    std::snprintf(buf, sizeof buf, "CUDA error in %s: %s", what, cudaGetErrorString(err));
    log_msg(state, buf);
    (void)cudaGetLastError();
    return false;
}

void free_layers(plugin_state* state) {
    if (state->copy_stream) cudaStreamSynchronize(state->copy_stream);  // the mirror copies read the layers
    if (state->layers[0]) cudaFree(state->layers[0]);
    for (auto& l : state->layers) l = nullptr;
}

void drop_mirror(plugin_state* state);

/// Counterpart of create_textures (plugin.mm:491-535): four RGBA32F layers,
/// accumulated in place, so one allocation of 64 B/px instead of twelve textures
/// of 16 B/px. The layers are contiguous (beauty, albedo, normal, depth) so a
/// multi-GPU host can reduce all four with a single collective.
bool create_layers(plugin_state* state, uint32_t width, uint32_t height) {
    if (state->render_pending) cudaEventSynchronize(state->last_render);
    free_layers(state);
    drop_mirror(state);
    size_t floats = static_cast<size_t>(width) * height * 4;
    if (floats != 0) {
        float* block = nullptr;
        if (!check(state, cudaMalloc(reinterpret_cast<void**>(&block),
                                     floats * Q_AOV_COUNT * sizeof(float)), "cudaMalloc(layers)")) {
            return false;
        }
        for (int i = 0; i < Q_AOV_COUNT; i++) state->layers[i] = block + floats * i;
    }
    state->last_width = width;
    state->last_height = height;
    state->frame_count = 0;
    return true;
}

/// Takes over the layer block a previous instance left with the host
/// (Q_cuda_persistent_state), if it belongs to this viewport, scene and shard.
/// A block that does not fit is released here, so the host never holds a stale one.
bool adopt_layers(plugin_state* state, uint32_t width, uint32_t height) {
    Q_cuda_persistent_state* p = state->persistent;
    if (!p || !p->valid || !p->layers) return false;
    const uint64_t bytes = static_cast<uint64_t>(width) * height * 4 * Q_AOV_COUNT * sizeof(float);
    const bool fits = p->width == width && p->height == height && p->bytes == bytes && bytes != 0 &&
                      p->scene_tag == state->scene_tag && p->shard_index == state->shard_index &&
                      p->shard_count == state->shard_count && p->device_ordinal == state->device;
    auto* block = static_cast<float*>(p->layers);
    p->valid = 0;
    p->layers = nullptr;
    if (!fits) {
        cudaFree(block);
        return false;
    }
    const size_t floats = static_cast<size_t>(width) * height * 4;
    for (int i = 0; i < Q_AOV_COUNT; i++) state->layers[i] = block + floats * i;
    state->last_width = width;
    state->last_height = height;
    state->frame_count = p->frame_count;
    log_msg(state, "Accumulation adopted from the host's persistent state");
    return true;
}

/// Hands the layer block to the host instead of freeing it (see adopt_layers).
bool leave_layers_with_host(plugin_state* state) {
    Q_cuda_persistent_state* p = state->persistent;
    if (!p || !state->layers[0] || state->frame_count == 0) return false;
    p->layers = state->layers[0];
    p->bytes = static_cast<uint64_t>(state->last_width) * state->last_height * 4 * Q_AOV_COUNT * sizeof(float);
    p->width = state->last_width;
    p->height = state->last_height;
    p->frame_count = state->frame_count;
    p->scene_tag = state->scene_tag;
    p->shard_index = state->shard_index;
    p->shard_count = state->shard_count;
    p->device_ordinal = state->device;
    p->valid = 1;
    for (auto& l : state->layers) l = nullptr;
    return true;
}

void free_wave(plugin_state* state) {
    if (state->wave_block) cudaFree(state->wave_block);
    state->wave_block = nullptr;
    state->wave = {};
}

/// Device memory for `paths` paths in flight as waves; keeps what is there if it is large enough.
bool ensure_wave(plugin_state* state, uint64_t paths) {
    if (state->wave_block && state->wave.capacity >= paths) return true;
    free_wave(state);
    const size_t counter_bytes = 2 * Q::cuda::k_mesh_wave_counters * sizeof(uint32_t);
    const size_t bytes = static_cast<size_t>(paths) * Q::cuda::mesh_wave_bytes_per_path() + 256 + counter_bytes;
    if (cudaMalloc(&state->wave_block, bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        state->wave_block = nullptr;
        return false;
    }
    auto* at = static_cast<float4*>(state->wave_block);
    auto take = [&at, paths]() {
        float4* p = at;
        at += paths;
        return p;
    };
    Q::cuda::mesh_wave_buffers& w = state->wave;
    for (int i = 0; i < 2; i++) {
        w.origin[i] = take();
        w.direction[i] = take();
        w.throughput[i] = take();
        w.colour[i] = take();
    }
    w.out_beauty = take();
    w.out_albedo_depth = take();
    w.out_normal = take();
    w.hit = reinterpret_cast<float2*>(at);
    w.counts = reinterpret_cast<uint32_t*>(w.hit + paths);
    w.capacity = paths;
    return true;
}

void free_mesh(plugin_state* state) {
    free_wave(state);
    if (state->mesh_keep.bytes != 0) {  // the lines the launches asked the L2 to keep are ordinary lines again
        cudaCtxResetPersistingL2Cache();
        state->mesh_keep = {};
    }
    for (auto& b : state->mesh_buffers) {
        if (b) cudaFree(b);
        b = nullptr;
    }
}

void finish_mesh_params(plugin_state* state, uint32_t max_bounces) {
    namespace mesh = Q::cuda::mesh;
    auto& d = state->mesh_device;
    d.nodes = static_cast<const float4*>(state->mesh_buffers[0]);
    d.tri_v0 = static_cast<const float4*>(state->mesh_buffers[1]);
    d.tri_e1 = static_cast<const float4*>(state->mesh_buffers[2]);
    d.tri_e2 = static_cast<const float4*>(state->mesh_buffers[3]);
    d.tri_shade = static_cast<const float4*>(state->mesh_buffers[4]);
    d.n_nodes = static_cast<uint32_t>(state->mesh.nodes.size());
    d.wide = 0;
    if (state->mesh_buffers[5]) {
        d.nodes = static_cast<const float4*>(state->mesh_buffers[5]);
        d.n_nodes = state->wide_nodes;
        d.wide = 1;
    }
    d.n_triangles = static_cast<uint32_t>(state->mesh.triangles.size());
    d.max_bounces = max_bounces;
    for (uint32_t m = 0; m < mesh::k_material_count; m++) {
        for (int c = 0; c < 3; c++) {
            d.materials[m * 8 + c] = state->mesh.materials[m].albedo[c];
            d.materials[m * 8 + 4 + c] = state->mesh.materials[m].emission[c];
        }
    }
}

/// BVH on the GPU (lbvh_build.cu): uploads the triangles in original order, builds,
/// gathers into leaf order on the device; the nodes and the leaf order come back to
/// the host only for Q_cuda_plugin_mesh_view. Returns false (nothing kept) if the
/// build fails or the tree is deeper than the traversal stack.
bool build_mesh_on_gpu(plugin_state* state) {
    namespace mesh = Q::cuda::mesh;
    const size_t n = state->mesh.triangles.size();
    std::vector<mesh::float4_t> original[4];
    mesh::pack_original_order(state->mesh, original[0], original[1], original[2], original[3]);
    void* input[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* order = nullptr;
    uint32_t* depth = nullptr;
    auto cleanup = [&](bool keep) {
        for (auto& b : input) {
            if (b) cudaFree(b);
        }
        if (order) cudaFree(order);
        if (depth) cudaFree(depth);
        if (!keep) free_mesh(state);
        return keep;
    };
    bool ok = true;
    for (int i = 0; i < 4 && ok; i++) {
        const size_t bytes = original[i].size() * sizeof(mesh::float4_t);
        ok = check(state, cudaMalloc(&input[i], bytes), "cudaMalloc(lbvh input)") &&
             check(state, cudaMemcpyAsync(input[i], original[i].data(), bytes, cudaMemcpyHostToDevice, state->queue),
                   "upload(lbvh input)");
    }
    const size_t out_bytes[5] = {sizeof(float4) * 2 * (2 * n - 1), sizeof(float4) * n, sizeof(float4) * n,
                                 sizeof(float4) * n, sizeof(float4) * 3 * n};
    for (int i = 0; i < 5 && ok; i++) ok = check(state, cudaMalloc(&state->mesh_buffers[i], out_bytes[i]), "cudaMalloc(mesh)");
    ok = ok && check(state, cudaMalloc(reinterpret_cast<void**>(&order), sizeof(uint32_t) * n), "cudaMalloc(order)") &&
         check(state, cudaMalloc(reinterpret_cast<void**>(&depth), sizeof(uint32_t)), "cudaMalloc(depth)");
    if (!ok) return cleanup(false);

    Q::cuda::lbvh_input in{static_cast<const float4*>(input[0]), static_cast<const float4*>(input[1]),
                           static_cast<const float4*>(input[2]), static_cast<const float4*>(input[3]),
                           static_cast<uint32_t>(n)};
    Q::cuda::lbvh_output out{static_cast<float4*>(state->mesh_buffers[0]), static_cast<float4*>(state->mesh_buffers[1]),
                             static_cast<float4*>(state->mesh_buffers[2]), static_cast<float4*>(state->mesh_buffers[3]),
                             static_cast<float4*>(state->mesh_buffers[4]), order, depth};
    float build_ms = 0.0f;
    if (!check(state, Q::cuda::build_lbvh(in, out, state->queue, &build_ms), "build_lbvh")) return cleanup(false);

    state->mesh.nodes.resize(2 * n - 1);
    state->mesh.leaf_order.resize(n);
    ok = check(state, cudaMemcpy(&state->mesh.depth, depth, sizeof(uint32_t), cudaMemcpyDeviceToHost), "download(depth)") &&
         check(state, cudaMemcpy(state->mesh.nodes.data(), state->mesh_buffers[0], out_bytes[0], cudaMemcpyDeviceToHost),
               "download(nodes)") &&
         check(state, cudaMemcpy(state->mesh.leaf_order.data(), order, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost),
               "download(order)");
    if (!ok) return cleanup(false);
    char buf[200];
    if (state->mesh.depth > Q::cuda::k_mesh_stack_entries) {
        std::snprintf(buf, sizeof buf, "GPU-built BVH has depth %u, the traversal stack holds %u: using the host builder",
                      state->mesh.depth, Q::cuda::k_mesh_stack_entries);
        log_msg(state, buf);
        return cleanup(false);
    }
    std::snprintf(buf, sizeof buf, "Mesh BVH built on the GPU in %.3f ms (LBVH, %zu node slots, depth %u)", build_ms,
                  2 * n - 1, state->mesh.depth);
    log_msg(state, buf);
    return cleanup(true);
}

uint32_t env_u32(const char* name, uint32_t fallback, uint32_t lo, uint32_t hi);

/// Generates the terrain-box scene, builds its BVH (binned SAH on the host by default,
/// LBVH on the GPU with QUASI_CUDA_MESH_BUILDER=lbvh) and leaves both on the device.
bool create_mesh(plugin_state* state, uint32_t grid, uint32_t seed, uint32_t max_bounces) {
    namespace mesh = Q::cuda::mesh;
    mesh::generate_triangles(grid, seed, state->mesh);
    const char* builder = std::getenv("QUASI_CUDA_MESH_BUILDER");
    bool on_gpu = builder && std::strcmp(builder, "lbvh") == 0 && state->mesh.triangles.size() > 4 &&
                  build_mesh_on_gpu(state);
    if (!on_gpu) {
        mesh::build_bvh(state->mesh);
        if (state->mesh.depth > Q::cuda::k_mesh_stack_entries) {
            log_msg(state, "mesh BVH is deeper than the traversal stack");
            return false;
        }
        std::vector<mesh::float4_t> packed[5];
        mesh::pack_for_device(state->mesh, packed[0], packed[1], packed[2], packed[3], packed[4]);
        for (int i = 0; i < 5; i++) {
            const size_t bytes = packed[i].size() * sizeof(mesh::float4_t);
            if (!check(state, cudaMalloc(&state->mesh_buffers[i], bytes), "cudaMalloc(mesh)") ||
                !check(state, cudaMemcpyAsync(state->mesh_buffers[i], packed[i].data(), bytes,
                                              cudaMemcpyHostToDevice, state->queue), "upload(mesh)")) {
                return false;
            }
        }
        if (!check(state, cudaStreamSynchronize(state->queue), "upload(mesh) sync")) return false;
    }
    // QUASI_CUDA_MESH_WIDE=1: the kernel walks the 4-wide tree collapsed from the binary one
    // (wide_bvh.hpp), whichever builder made that; the binary tree stays what
    // Q_cuda_plugin_mesh_view shows.
    if (env_u32("QUASI_CUDA_MESH_WIDE", 0, 0, 1) != 0) {
        // one entry of the stack stays free: the wide visit stores before it knows whether to keep
        constexpr uint32_t limit = Q::cuda::k_mesh_stack_entries - 1;
        const mesh::wide_tree wide = mesh::collapse_wide(state->mesh, limit);
        const size_t bytes = wide.nodes.size() * sizeof(mesh::float4_t);
        if (wide.node_count == 0 || wide.stack_need > limit) {
            log_msg(state, "4-wide tree does not fit the traversal stack: using the binary tree");
        } else if (!check(state, cudaMalloc(&state->mesh_buffers[5], bytes), "cudaMalloc(wide nodes)") ||
                   !check(state, cudaMemcpy(state->mesh_buffers[5], wide.nodes.data(), bytes, cudaMemcpyHostToDevice),
                          "upload(wide nodes)")) {
            return false;
        } else {
            state->wide_nodes = wide.node_count;
            char msg[200];
            std::snprintf(msg, sizeof msg, "Mesh BVH collapsed to %u 4-wide nodes (%.2f children each), depth %u, stack bound %u",
                          wide.node_count, static_cast<double>(wide.children) / wide.node_count, wide.depth, wide.stack_need);
            log_msg(state, msg);
        }
    }
    finish_mesh_params(state, max_bounces);
    // The binary tree once more with both children of a node in one 32-byte sector (quantized_bvh.hpp), for the
    // wavefront trace kernel; QUASI_CUDA_MESH_QUANT=0 keeps it on the float tree.
    if (state->mesh_wavefront && state->mesh_device.wide == 0 && env_u32("QUASI_CUDA_MESH_QUANT", 1, 0, 1) != 0) {
        const mesh::quantized_tree grid = mesh::quantize_pairs(state->mesh);
        const size_t bytes = grid.pairs.size() * sizeof(uint32_t);
        if (bytes == 0) {
            log_msg(state, "mesh BVH is not numbered in pairs: the trace kernel walks the float tree");
        } else if (!check(state, cudaMalloc(&state->mesh_buffers[6], bytes), "cudaMalloc(quantized pairs)") ||
                   !check(state, cudaMemcpy(state->mesh_buffers[6], grid.pairs.data(), bytes, cudaMemcpyHostToDevice),
                          "upload(quantized pairs)")) {
            return false;
        } else {
            state->mesh_device.pairs = static_cast<const uint4*>(state->mesh_buffers[6]);
            for (int a = 0; a < 3; a++) {
                state->mesh_device.grid_origin[a] = grid.origin[a];
                state->mesh_device.grid_step[a] = grid.step[a];
            }
        }
    }
    // QUASI_CUDA_MESH_L2_KEEP=<percent>: the launches ask the L2 to keep the BVH nodes (a persisting access window
    // over the node array with that hit ratio), up to the share of the L2 the device lets be set aside.
    if (const uint32_t percent = env_u32("QUASI_CUDA_MESH_L2_KEEP", 0, 0, 100); percent != 0) {
        int most = 0, window = 0;
        cudaDeviceGetAttribute(&most, cudaDevAttrMaxPersistingL2CacheSize, state->device);
        cudaDeviceGetAttribute(&window, cudaDevAttrMaxAccessPolicyWindowSize, state->device);
        const size_t node_bytes = static_cast<size_t>(state->mesh_device.n_nodes) * (state->mesh_device.wide ? 128u : 32u);
        const size_t aside = std::min<size_t>(static_cast<size_t>(most), node_bytes);
        if (aside != 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, aside) == cudaSuccess) {
            state->mesh_keep.base = state->mesh_device.nodes;
            state->mesh_keep.bytes = std::min<size_t>(node_bytes, static_cast<size_t>(window));
            state->mesh_keep.hit_ratio = std::min(1.0f, static_cast<float>(aside) / static_cast<float>(state->mesh_keep.bytes)) *
                                         static_cast<float>(percent) / 100.0f;
            char msg[200];
            std::snprintf(msg, sizeof msg, "Mesh nodes: %zu MB asked to stay in the L2 (set aside %zu of at most %d MB, hit ratio %.2f)",
                          state->mesh_keep.bytes >> 20, aside >> 20, most >> 20, static_cast<double>(state->mesh_keep.hit_ratio));
            log_msg(state, msg);
        } else {
            (void)cudaGetLastError();
            log_msg(state, "QUASI_CUDA_MESH_L2_KEEP: this device sets no L2 aside");
        }
    }
    char buf[160];
    std::snprintf(buf, sizeof buf, "Mesh scene: %u triangles, %u BVH nodes, depth %u", state->mesh_device.n_triangles,
                  static_cast<uint32_t>(state->mesh.nodes.size()), state->mesh.depth);
    log_msg(state, buf);
    return true;
}

uint32_t env_u32(const char* name, uint32_t fallback, uint32_t lo, uint32_t hi) {
    const char* s = std::getenv(name);
    if (!s || !*s) return fallback;
    long v = std::strtol(s, nullptr, 10);
    if (v < static_cast<long>(lo)) v = lo;
    if (v > static_cast<long>(hi)) v = hi;
    return static_cast<uint32_t>(v);
}

Q_aov_buffer download_layer(plugin_state* state, int layer) {
    Q_aov_buffer buf{};
    const float* src = state->layers[layer];
    if (!src || state->frame_count == 0) return buf;
    size_t bytes = static_cast<size_t>(state->last_width) * state->last_height * 4 * sizeof(float);
    float* dst = g_pool.acquire(bytes);
    if (!dst) return buf;
    if (!check(state, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, state->queue),
               "cudaMemcpyAsync(readback)")) {
        g_pool.release(dst);
        return buf;
    }
    buf.data = dst;
    buf.width = state->last_width;
    buf.height = state->last_height;
    buf.channels = 4;
    return buf;
}

/// Orders `stream` after the most recent render without blocking the host.
void order_after_render(plugin_state* state, cudaStream_t stream) {
    if (state->render_pending) cudaStreamWaitEvent(stream, state->last_render, 0);
    // (mesh scene: what follows may write the layers that the last batch's mirror copies read)
    if (state->mirror_by_copy && state->mirror_copied) cudaStreamWaitEvent(stream, state->mirror_copied, 0);
}


void drop_mirror(plugin_state* state) {
    if (state->copy_stream) cudaStreamSynchronize(state->copy_stream);  // nothing is still being copied into them
    state->mirror_by_copy = false;
    for (auto& m : state->mirror) {
        if (m) g_pool.release(m);
        m = nullptr;
    }
    state->mirror_valid = false;
}

/// Launches the render calls collected under QUASI_CUDA_BATCH as one kernel. The arguments are
/// the ones Q_plugin_render computes for a single call, for the first collected call; sample s of
/// the launch is what the (s+1)-th collected call would have rendered.
/// `stream`: the stream of the render call that triggers the launch; null when something else does (a readback, the
/// statistics, destroy): then the launch goes to the plugin's own queue, because the stream of an earlier render
/// call may no longer exist (a host may create one per frame). Either way it is ordered after the previous launch.
/// If the launch cannot be made, the collected calls are taken back (frame_count and the path statistic), so that
/// the samples that follow carry the weights of what is really in the layers.
/// The stream and events behind the mesh scene's mirror copies, made on first use.
bool mirror_copies_ready(plugin_state* state) {
    if (state->copy_stream) return true;
    if (cudaStreamCreateWithFlags(&state->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&state->layers_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&state->mirror_copied, cudaEventDisableTiming) != cudaSuccess) {
        (void)cudaGetLastError();
        if (state->copy_stream) cudaStreamDestroy(state->copy_stream);
        if (state->layers_ready) cudaEventDestroy(state->layers_ready);
        if (state->mirror_copied) cudaEventDestroy(state->mirror_copied);
        state->copy_stream = nullptr;
        state->layers_ready = state->mirror_copied = nullptr;
        return false;
    }
    return true;
}

/// Whether a fused-reduce table (Q_cuda_plugin_set_reduce) was cut for this viewport: the slices cover the four
/// layers and each starts on a row of tiles of one layer, so that a tile's 32 pixels in a layer have one owner.
bool reduce_table_fits(const Q_cuda_reduce_table& t, uint32_t width, uint32_t height) {
    if (width == 0 || height == 0 || t.slice_lo[t.ranks] != 4ull * width * height) return false;
    for (uint32_t i = 0; i < t.ranks; i++) {
        const uint32_t lo = t.slice_lo[i];
        if (lo % width != 0 || ((lo / width) % height) % 4 != 0) return false;
    }
    return true;
}

void flush_pending(plugin_state* state, cudaStream_t stream = nullptr) {
    if (state->pending == 0) return;
    const Q_render_frame& f = state->pending_frame;
    if (stream == nullptr) stream = state->queue;
    const uint32_t collected = state->pending;
    state->pending = 0;
    state->mirror_valid = false;
    auto take_back = [state, collected, &f]() {
        state->frame_count -= collected < state->frame_count ? collected : state->frame_count;
        state->paths -= static_cast<uint64_t>(collected) * f.width * f.height;
        state->mirror_valid = false;
    };
    frame_params fp;
    const uint32_t frame_index = state->shard_index + state->pending_first_sample * state->shard_count;
    Q::cuda::prepare_frame(f.camera.position, f.camera.target, f.camera.up, f.camera.fov, f.width, f.height, frame_index,
                           state->pending_first_sample, fp);
    fp.batch_samples = collected;
    fp.frame_stride = state->shard_count;
    if (state->use_mesh) {
        // The mesh scene: the collected calls as waves of rays (mesh_wavefront.cu); without the memory for that,
        // one launch of the one-kernel form per call, exactly what Q_plugin_render would have made.
        order_after_render(state, stream);
        const uint64_t n_px = static_cast<uint64_t>(f.width) * f.height;
        cudaError_t launched = cudaSuccess;
        const uint64_t entries = static_cast<uint64_t>((f.width + 7u) / 8u) * ((f.height + 3u) / 4u) * 32u * collected;
        bool mirrored = false;
        // (a wave has a ramp and a tail of its own, and a batch is 2 * bounces + 3 launches: measured per 16 samples at
        // 1080p, batches of 1 take 37 ms, of 4 26.1 ms, of 16 16.9 ms, the one-kernel form 26.3 ms)
        if (collected >= state->mesh_wave_min && n_px * collected < (uint64_t{1} << 31) && ensure_wave(state, entries)) {
            // The readback mirror (QUASI_CUDA_MIRROR), the wavefront's way: a copy engine takes albedo, normal and depth
            // to pinned host memory as soon as the first shade pass has made them final, under the waves that
            // follow, and the beauty layer after the last launch; neither copy is on the render's stream.
            const size_t layer_bytes = static_cast<size_t>(n_px) * 4 * sizeof(float);
            bool mirror = state->mirror_enabled && !state->external_export && !state->external_reduce && mirror_copies_ready(state);
            for (int i = 0; i < Q_AOV_COUNT && mirror; i++) {
                if (!state->mirror[i]) state->mirror[i] = g_pool.acquire(layer_bytes);
                mirror = state->mirror[i] != nullptr;
            }
            auto copy_first_hit_layers = [&] {
                bool ok = cudaEventRecord(state->layers_ready, stream) == cudaSuccess &&
                          cudaStreamWaitEvent(state->copy_stream, state->layers_ready, 0) == cudaSuccess;
                for (int i = 1; i < Q_AOV_COUNT && ok; i++) {
                    ok = cudaMemcpyAsync(state->mirror[i], state->layers[i], layer_bytes, cudaMemcpyDeviceToHost,
                                         state->copy_stream) == cudaSuccess;
                }
                mirror = ok;
            };
            uint32_t made = 0;
            launched = Q::cuda::launch_mesh_wavefront(state->mesh_device, fp, state->wave, state->layers, state->ray_counter,
                                                      state->sm_count, stream, &made,
                                                      mirror ? std::function<void()>(copy_first_hit_layers) : std::function<void()>());
            state->kernel_launches += made;
            if (launched == cudaSuccess && mirror) {
                mirrored = cudaEventRecord(state->layers_ready, stream) == cudaSuccess &&
                           cudaStreamWaitEvent(state->copy_stream, state->layers_ready, 0) == cudaSuccess &&
                           cudaMemcpyAsync(state->mirror[0], state->layers[0], layer_bytes, cudaMemcpyDeviceToHost,
                                           state->copy_stream) == cudaSuccess &&
                           cudaEventRecord(state->mirror_copied, state->copy_stream) == cudaSuccess;
                if (!mirrored) (void)cudaGetLastError();
            }
        } else {
            auto* tile_counters = reinterpret_cast<unsigned int*>(state->ray_counter + 3);
            for (uint32_t s = 0; s < collected && launched == cudaSuccess; s++) {
                frame_params one;
                Q::cuda::prepare_frame(f.camera.position, f.camera.target, f.camera.up, f.camera.fov, f.width, f.height,
                                       frame_index + s * state->shard_count, state->pending_first_sample + s, one);
                launched = Q::cuda::launch_mesh_pathtrace(state->mesh_device, one, state->layers, state->ray_counter,
                                                          tile_counters, state->sm_count, stream, state->mesh_keep);
                state->kernel_launches++;
            }
        }
        if (!check(state, launched, "mesh pathtrace (batch)")) {
            take_back();
            return;
        }
        state->last_launch_exported = false;
        state->mirror_valid = mirrored;
        state->mirror_by_copy = mirrored;
        cudaEventRecord(state->last_render, stream);
        state->render_pending = true;
        return;
    }
    Q::cuda::batch_params launch{};
    bool mirrored = false;
    if (collected > 1u) {
        const uint64_t n_px = static_cast<uint64_t>(f.width) * f.height;
        const uint32_t total = static_cast<uint32_t>(4 * n_px);
        if (state->external_reduce && state->reduce_calls_seen < state->reduce_table.render_calls) {
            // not the frame's last launch: its tiles are not final, nothing is exported and nothing arrives
        } else if (state->external_reduce && n_px < (uint64_t{1} << 30) && reduce_table_fits(state->reduce_table, f.width, f.height)) {
            const Q_cuda_reduce_table& t = state->reduce_table;
            launch.export_slices = t.ranks;
            launch.reduce_ranks = t.ranks;
            launch.reduce_slot_stride = t.slot_stride;
            launch.reduce_scale = 1.0f / static_cast<float>(t.ranks);
            for (uint32_t i = 0; i <= t.ranks; i++) launch.slice_lo[i] = t.slice_lo[i];
            for (uint32_t i = 0; i < t.ranks; i++) {
                launch.reduce_staging[i] = t.staging[i];
                launch.reduce_counters[i] = t.counters[i];
                launch.reduce_out[i] = t.out[i];
                launch.export_base[i] = static_cast<float4*>(t.staging[i]) + static_cast<size_t>(t.rank) * t.slot_stride;
            }
        } else if (state->external_export && n_px < (uint64_t{1} << 30)) {
            const Q_cuda_export_table& t = state->export_table;
            if (t.slices != 0) {  // (a table of no slices: nothing leaves the GPU during the launch)
                launch.export_slices = t.slices;
                for (uint32_t i = 0; i < t.slices; i++) {
                    launch.export_base[i] = t.base[i];
                    const uint64_t lo = static_cast<uint64_t>(i) * t.float4_per_slice;
                    launch.slice_lo[i] = lo < total ? static_cast<uint32_t>(lo) : total;
                }
                launch.slice_lo[t.slices] = total;
            }
        } else if (state->mirror_enabled && !state->external_export && !state->external_reduce && n_px < (uint64_t{1} << 30)) {
            const size_t bytes = static_cast<size_t>(n_px) * 4 * sizeof(float);
            bool have = true;
            for (int i = 0; i < Q_AOV_COUNT && have; i++) {
                // buffers handed to a caller are never written again: a readback takes them away, fresh ones come here
                if (!state->mirror[i]) state->mirror[i] = g_pool.acquire(bytes);
                have = state->mirror[i] != nullptr;
            }
            if (have) {  // one slice per layer
                launch.export_slices = Q_AOV_COUNT;
                for (int i = 0; i < Q_AOV_COUNT; i++) {
                    launch.export_base[i] = state->mirror[i];
                    launch.slice_lo[i] = static_cast<uint32_t>(i * n_px);
                }
                launch.slice_lo[Q_AOV_COUNT] = total;
                mirrored = true;
            }
        }
    }
    Q::cuda::set_decision_margins(state->scene, f.camera.position);
    order_after_render(state, stream);
    auto* tile_counters = reinterpret_cast<unsigned int*>(state->ray_counter + 3);
    const cudaError_t launched = Q::cuda::launch_pathtrace_accumulate(state->scene, fp, launch, state->layers, state->ray_counter,
                                                                      tile_counters, state->sm_count, stream, state->generic_scan);
    if (!check(state, launched, "pathtrace_accumulate (batch)")) {
        take_back();
        return;
    }
    state->kernel_launches++;
    state->mirror_valid = mirrored;
    state->last_launch_exported = launch.export_slices != 0u;
    cudaEventRecord(state->last_render, stream);
    state->render_pending = true;
}

/// Q_plugin_render under QUASI_CUDA_BATCH: the bookkeeping of one call (plugin.mm:637-646, 791),
/// the launch deferred until `batch` calls with the same camera, size and stream have come in.
void render_batched(plugin_state* state, const Q_render_frame* frame) {
    // The stream may differ from call to call (a host with several frames in flight hands out one
    // per frame): the launch goes to the stream of the last collected call, ordered after the
    // previous launch by order_after_render like any call on a new stream.
    const bool same_view = state->pending != 0 && frame->width == state->pending_frame.width &&
                           frame->height == state->pending_frame.height && frame->camera_dirty == 0 &&
                           std::memcmp(&frame->camera, &state->pending_frame.camera, sizeof frame->camera) == 0;
    auto stream = static_cast<cudaStream_t>(frame->command_buffer);
    if (!same_view) flush_pending(state, stream);
    if (frame->width != state->last_width || frame->height != state->last_height) {
        if (!create_layers(state, frame->width, frame->height)) return;
    }
    if (frame->width == 0 || frame->height == 0) return;
    if (frame->camera_dirty) state->frame_count = 0;
    if (state->pending == 0) {
        state->pending_frame = *frame;
        state->pending_first_sample = state->frame_count;
    }
    state->pending_frame.command_buffer = frame->command_buffer;
    state->pending++;
    state->reduce_calls_seen++;
    state->paths += static_cast<uint64_t>(frame->width) * frame->height;
    state->frame_count++;
    if (state->pending >= state->batch) flush_pending(state, stream);
}

}  // namespace

extern "C" {

/// plugin.mm:543-545
Q_PLUGIN_EXPORT uint32_t Q_plugin_abi_version(void) { return Q_PLUGIN_ABI_VERSION; }

/// plugin.mm:547-554
Q_PLUGIN_EXPORT Q_plugin_info Q_plugin_get_info(void) {
    Q_plugin_info info;
    info.name = k_name;
    info.version = {1, 0, 0};
    info.description = k_description;
    info.author = k_author;
    return info;
}

/// plugin.mm:556-592
Q_PLUGIN_EXPORT Q_plugin_handle* Q_plugin_create(Q_plugin_context* ctx) {
    if (!ctx || !ctx->gpu) return nullptr;

    if (ctx->gpu->backend != Q_GPU_BACKEND_CUDA) {
        if (ctx->log) ctx->log(ctx->host_data, "CUDA backend requires CUDA GPU context");
        return nullptr;
    }

    auto* state = new (std::nothrow) plugin_state{};
    if (!state) return nullptr;
    state->context = ctx;
    state->queue = static_cast<cudaStream_t>(ctx->gpu->queue);

    const auto* desc = static_cast<const Q_cuda_device_desc*>(ctx->gpu->device);
    if (desc && desc->struct_size >= sizeof(Q_cuda_device_desc)) {
        state->device = desc->device_ordinal;
        state->shard_count = desc->shard_count ? desc->shard_count : 1u;
        state->shard_index = desc->shard_index % state->shard_count;
    } else if (cudaGetDevice(&state->device) != cudaSuccess) {
        state->device = 0;
    }

    bool ok = check(state, cudaSetDevice(state->device), "cudaSetDevice");
    ok = ok && check(state, cudaDeviceGetAttribute(&state->sm_count, cudaDevAttrMultiProcessorCount,
                                                   state->device), "cudaDeviceGetAttribute");
    ok = ok && check(state, cudaEventCreateWithFlags(&state->last_render, cudaEventDisableTiming),
                     "cudaEventCreate");
    // counters: rays, BVH nodes visited, triangles tested; then the tile work queue (2 x uint32)
    ok = ok && check(state, cudaMalloc(reinterpret_cast<void**>(&state->ray_counter),
                                       4 * sizeof(unsigned long long)), "cudaMalloc(counter)");
    ok = ok && check(state, cudaMemsetAsync(state->ray_counter, 0, 4 * sizeof(unsigned long long),
                                            state->queue), "cudaMemset(counter)");
    ok = ok && check(state, cudaStreamSynchronize(state->queue), "cudaMemset(counter) sync");
    if (ok) {
        // The plugin builds its own scene (plugin.mm:581-584). The ABI has no
        // slot for options, so they come from the environment.
        state->scene_desc = Q::cuda::make_cornell_box();
        uint32_t bounces = env_u32("QUASI_CUDA_MAX_BOUNCES", Q::cuda::k_max_bounces, 1, 64);
        // QUASI_CUDA_SCAN: "generic" = the any-scene kernel, "boxes" = the five faces of a block
        // in one slab test (scan_decide.hpp:scan_box), "generic-boxes" = both; default: every
        // quad on its own in the kernel unrolled for the demo scene, the fastest form.
        const char* scan = std::getenv("QUASI_CUDA_SCAN");
        state->generic_scan = scan && std::strncmp(scan, "generic", 7) == 0;
        const bool group_boxes = scan && std::strstr(scan, "boxes") != nullptr;
        Q::cuda::prepare_scene(state->scene_desc, bounces, state->scene, group_boxes);
        state->batch = env_u32("QUASI_CUDA_BATCH", 64, 1, 1024);
        state->mirror_enabled = env_u32("QUASI_CUDA_MIRROR", 1, 0, 1) != 0;
        const char* scene_name = std::getenv("QUASI_CUDA_SCENE");
        if (scene_name && std::strcmp(scene_name, "mesh") == 0) {
            state->use_mesh = true;
            // QUASI_CUDA_MESH_WAVEFRONT=0: one launch of the one-kernel path tracer per call (mesh_trace.cu); default:
            // headless calls are collected, QUASI_CUDA_MESH_BATCH at a time, and traced as waves (mesh_wavefront.cu)
            state->mesh_wavefront = env_u32("QUASI_CUDA_MESH_WAVEFRONT", 1, 0, 1) != 0;
            state->batch = state->mesh_wavefront ? env_u32("QUASI_CUDA_MESH_BATCH", 16, 1, 64) : 1;
            state->mesh_wave_min = env_u32("QUASI_CUDA_MESH_WAVEFRONT_MIN", 4, 1, 64);
            ok = create_mesh(state, env_u32("QUASI_CUDA_MESH_GRID", 707, 1, 4000),
                             env_u32("QUASI_CUDA_MESH_SEED", 1, 0, 65535), bounces);
        }
        state->scene_tag = state->use_mesh
                               ? 0x6d000000u ^ (env_u32("QUASI_CUDA_MESH_GRID", 707, 1, 4000) << 8) ^
                                     (env_u32("QUASI_CUDA_MESH_SEED", 1, 0, 65535) << 20) ^ bounces
                               : 0x71000000u ^ bounces;
        auto* persistent = static_cast<Q_cuda_persistent_state*>(ctx->gpu->layer);
        if (persistent && persistent->struct_size >= sizeof(Q_cuda_persistent_state)) state->persistent = persistent;
        ok = ok && (adopt_layers(state, ctx->viewport_width, ctx->viewport_height) ||
                    create_layers(state, ctx->viewport_width, ctx->viewport_height));
    }
    if (!ok) {
        free_mesh(state);
        free_layers(state);
        if (state->ray_counter) cudaFree(state->ray_counter);
        if (state->last_render) cudaEventDestroy(state->last_render);
        delete state;
        return nullptr;
    }

    {
        std::lock_guard<std::mutex> lock(g_instances_mutex);
        g_instances++;
    }
    log_msg(state, "Cornell Box path tracer initialized (CUDA)");
    return reinterpret_cast<Q_plugin_handle*>(state);
}

/// plugin.mm:594-617. Also drains the GPU: after this returns the library may be
/// dlclose'd by the hot-reload manager (manager.hpp:350-360).
Q_PLUGIN_EXPORT void Q_plugin_destroy(Q_plugin_handle* handle) {
    if (!handle) return;
    auto* state = reinterpret_cast<plugin_state*>(handle);
    log_msg(state, "Path tracer destroyed");

    cudaSetDevice(state->device);
    flush_pending(state);
    if (state->render_pending) cudaEventSynchronize(state->last_render);
    cudaStreamSynchronize(state->queue);
    free_mesh(state);
    if (!leave_layers_with_host(state)) free_layers(state);
    drop_mirror(state);
    if (state->copy_stream) cudaStreamDestroy(state->copy_stream);
    if (state->layers_ready) cudaEventDestroy(state->layers_ready);
    if (state->mirror_copied) cudaEventDestroy(state->mirror_copied);
    if (state->ray_counter) cudaFree(state->ray_counter);
    if (state->last_render) cudaEventDestroy(state->last_render);
    delete state;

    bool last;
    {
        std::lock_guard<std::mutex> lock(g_instances_mutex);
        last = --g_instances == 0;
    }
    if (last) g_pool.trim();
}

/// plugin.mm:619-622
Q_PLUGIN_EXPORT void Q_plugin_update(Q_plugin_handle* handle, float delta_time) {
    (void)handle;
    (void)delta_time;
}

/// plugin.mm:624-792
Q_PLUGIN_EXPORT void Q_plugin_render(Q_plugin_handle* handle, Q_render_frame* frame) {
    if (!handle || !frame) return;
    auto* state = reinterpret_cast<plugin_state*>(handle);

    // The Metal backend needs a drawable and a command buffer (plugin.mm:632).
    // Here the command buffer is the frame's stream and is required; the
    // drawable is optional because a headless host has nothing to present.
    auto stream = static_cast<cudaStream_t>(frame->command_buffer);
    if (!stream) return;
    const auto* drawable = static_cast<const Q_cuda_drawable*>(frame->drawable);

    if (cudaSetDevice(state->device) != cudaSuccess) return;

    if ((state->batch > 1 || state->mesh_wavefront) && !(drawable && drawable->pixels) && frame->width < 65536u &&
        frame->height < 65536u) {  // the batched kernel keeps a lane's pixel as two 16-bit coordinates
        render_batched(state, frame);
        return;
    }
    flush_pending(state, stream);  // a call that presents, after headless ones that were being collected
    state->mirror_valid = false;

    const uint32_t width = frame->width;
    const uint32_t height = frame->height;

    // plugin.mm:637-640
    if (width != state->last_width || height != state->last_height) {
        if (!create_layers(state, width, height)) return;
    }
    if (width == 0 || height == 0) return;

    // plugin.mm:642-646
    if (frame->camera_dirty) state->frame_count = 0;

    // plugin.mm:648-677; frame index per shard: SURVEY.md 8e.
    frame_params fp;
    const uint32_t frame_index = state->shard_index + state->frame_count * state->shard_count;
    Q::cuda::prepare_frame(frame->camera.position, frame->camera.target, frame->camera.up,
                           frame->camera.fov, width, height, frame_index, state->frame_count, fp);

    Q::cuda::set_decision_margins(state->scene, frame->camera.position);

    // Frames may arrive on different streams (one per frame in flight); the
    // layers are shared, so order this frame after the previous one.
    order_after_render(state, stream);

    auto* tile_counters = reinterpret_cast<unsigned int*>(state->ray_counter + 3);
    const cudaError_t launched =
        state->use_mesh ? Q::cuda::launch_mesh_pathtrace(state->mesh_device, fp, state->layers, state->ray_counter,
                                                         tile_counters, state->sm_count, stream, state->mesh_keep)
                        : Q::cuda::launch_pathtrace_accumulate(state->scene, fp, Q::cuda::batch_params{}, state->layers,
                                                               state->ray_counter, tile_counters, state->sm_count,
                                                               stream, state->generic_scan);
    if (!check(state, launched, "pathtrace_accumulate")) return;
    state->kernel_launches++;
    state->last_launch_exported = false;

    // plugin.mm:776-788
    if (drawable && drawable->pixels) {
        uint32_t pitch = drawable->pitch_bytes ? drawable->pitch_bytes : 4u * width;
        if (check(state, Q::cuda::launch_tonemap_bgra8(state->layers[Q_AOV_BEAUTY], width, height,
                                                       drawable->pixels, pitch, stream),
                  "tonemap")) {
            state->kernel_launches++;
        }
    }

    cudaEventRecord(state->last_render, stream);
    state->render_pending = true;

    state->paths += static_cast<uint64_t>(width) * height;
    state->frame_count++;  // plugin.mm:791
}

/// plugin.mm:846-863
Q_PLUGIN_EXPORT Q_readback_result Q_plugin_readback(Q_plugin_handle* handle) {
    Q_readback_result result{};
    if (!handle) return result;
    auto* state = reinterpret_cast<plugin_state*>(handle);
    if (cudaSetDevice(state->device) != cudaSuccess) return result;
    flush_pending(state);

    if (state->mirror_valid && state->render_pending &&
        check(state, cudaEventSynchronize(state->last_render), "readback sync (mirror)") &&
        (!state->mirror_by_copy || check(state, cudaEventSynchronize(state->mirror_copied), "readback sync (mirror copies)"))) {
        // the batched launch has put the frame into pinned host buffers itself: hand the beauty layer over
        result.data = state->mirror[Q_AOV_BEAUTY];
        result.width = state->last_width;
        result.height = state->last_height;
        result.channels = 4;
        state->mirror[Q_AOV_BEAUTY] = nullptr;  // the caller's now; the next batch gets a fresh one
        state->mirror_valid = false;
        log_msg(state, "HDR readback complete (mirrored during the render)");
        return result;
    }
    order_after_render(state, state->queue);
    Q_aov_buffer buf = download_layer(state, Q_AOV_BEAUTY);
    if (buf.data && !check(state, cudaStreamSynchronize(state->queue), "readback sync")) {
        g_pool.release(buf.data);
        buf = Q_aov_buffer{};
    }
    result.data = buf.data;
    result.width = buf.width;
    result.height = buf.height;
    result.channels = buf.channels;
    if (buf.data) log_msg(state, "HDR readback complete");
    return result;
}

/// plugin.mm:865-870
Q_PLUGIN_EXPORT void Q_plugin_readback_free(Q_readback_result* result) {
    if (result && result->data) {
        g_pool.release(result->data);
        result->data = nullptr;
    }
}

/// plugin.mm:872-892. The four copies are enqueued back to back and waited for
/// once, instead of four blit + waitUntilCompleted round trips.
Q_PLUGIN_EXPORT Q_readback_aov_result Q_plugin_readback_aov(Q_plugin_handle* handle) {
    Q_readback_aov_result result{};
    if (!handle) return result;
    auto* state = reinterpret_cast<plugin_state*>(handle);
    if (cudaSetDevice(state->device) != cudaSuccess) return result;
    flush_pending(state);

    if (state->mirror_valid) {
        // the batched launch has put the frame into pinned host buffers itself: wait for it, hand them over
        if (state->render_pending && check(state, cudaEventSynchronize(state->last_render), "readback sync (mirror)") &&
            (!state->mirror_by_copy || check(state, cudaEventSynchronize(state->mirror_copied), "readback sync (mirror copies)"))) {
            for (int i = 0; i < Q_AOV_COUNT; i++) {
                result.buffers[i] = Q_aov_buffer{state->mirror[i], state->last_width, state->last_height, 4};
                state->mirror[i] = nullptr;  // the caller's now; the next batch gets fresh ones
            }
            state->mirror_valid = false;
            log_msg(state, "AOV readback complete (mirrored during the render)");
            return result;
        }
        state->mirror_valid = false;
    }

    order_after_render(state, state->queue);
    for (int i = 0; i < Q_AOV_COUNT; i++) result.buffers[i] = download_layer(state, i);
    if (!check(state, cudaStreamSynchronize(state->queue), "readback sync")) {
        for (auto& b : result.buffers) {
            if (b.data) g_pool.release(b.data);
            b = Q_aov_buffer{};
        }
        return result;
    }
    log_msg(state, "AOV readback complete");
    return result;
}

/// plugin.mm:894-902
Q_PLUGIN_EXPORT void Q_plugin_readback_aov_free(Q_readback_aov_result* result) {
    if (!result) return;
    for (auto& b : result->buffers) {
        if (b.data) {
            g_pool.release(b.data);
            b.data = nullptr;
        }
    }
}

/* ---- CUDA-only extension entry points (quasi_cuda_backend.h) ---- */

Q_PLUGIN_EXPORT int Q_cuda_plugin_stats(Q_plugin_handle* handle, Q_cuda_stats* out) {
    if (!handle || !out) return -1;
    auto* state = reinterpret_cast<plugin_state*>(handle);
    if (cudaSetDevice(state->device) != cudaSuccess) return -1;
    flush_pending(state);
    if (state->render_pending) cudaEventSynchronize(state->last_render);
    unsigned long long counters[3] = {0, 0, 0};
    order_after_render(state, state->queue);
    if (!check(state, cudaMemcpyAsync(counters, state->ray_counter, sizeof counters,
                                      cudaMemcpyDeviceToHost, state->queue), "stats copy") ||
        !check(state, cudaStreamSynchronize(state->queue), "stats sync")) {
        return -1;
    }
    std::memset(out, 0, sizeof *out);
    out->rays = counters[0];
    out->nodes_visited = counters[1];
    out->triangles_tested = counters[2];
    out->paths = state->paths;
    out->kernel_launches = state->kernel_launches;
    out->frame_count = state->frame_count;
    out->width = state->last_width;
    out->height = state->last_height;
    return 0;
}

Q_PLUGIN_EXPORT int Q_cuda_plugin_layers(Q_plugin_handle* handle, void* stream,
                                         Q_cuda_layers* out) {
    if (!handle || !out) return -1;
    auto* state = reinterpret_cast<plugin_state*>(handle);
    if (cudaSetDevice(state->device) != cudaSuccess) return -1;
    flush_pending(state);
    order_after_render(state, static_cast<cudaStream_t>(stream));
    std::memset(out, 0, sizeof *out);
    for (int i = 0; i < Q_AOV_COUNT; i++) out->device_ptr[i] = state->layers[i];
    out->width = state->last_width;
    out->height = state->last_height;
    out->frame_count = state->frame_count;
    return 0;
}

Q_PLUGIN_EXPORT int Q_cuda_plugin_set_export(Q_plugin_handle* handle, const Q_cuda_export_table* table) {
    if (!handle) return -1;
    auto* state = reinterpret_cast<plugin_state*>(handle);
    if (cudaSetDevice(state->device) != cudaSuccess) return -1;
    if (table != nullptr) {
        if (table->struct_size < sizeof(Q_cuda_export_table) || table->slices > Q_CUDA_MAX_EXPORT_SLICES ||
            (table->slices != 0 && table->float4_per_slice == 0)) {
            return -1;
        }
        for (uint32_t i = 0; i < table->slices; i++) {
            if (table->base[i] == nullptr) return -1;
        }
    }
    flush_pending(state);  // what has been collected goes where it was meant to go
    state->external_export = table != nullptr;
    if (table != nullptr) {
        state->export_table = *table;
        state->external_reduce = false;
    }
    state->mirror_valid = false;
    return 0;
}

Q_PLUGIN_EXPORT int Q_cuda_plugin_set_reduce(Q_plugin_handle* handle, const Q_cuda_reduce_table* table) {
    if (!handle) return -1;
    auto* state = reinterpret_cast<plugin_state*>(handle);
    if (cudaSetDevice(state->device) != cudaSuccess) return -1;
    if (table != nullptr) {
        if (table->struct_size < sizeof(Q_cuda_reduce_table) || table->ranks < 2 || table->ranks > Q_CUDA_MAX_EXPORT_SLICES ||
            table->rank >= table->ranks || table->slice_lo[0] != 0) {
            return -1;
        }
        for (uint32_t i = 0; i < table->ranks; i++) {
            const uint32_t lo = table->slice_lo[i], hi = table->slice_lo[i + 1];
            if (table->staging[i] == nullptr || table->counters[i] == nullptr || table->out[i] == nullptr || hi < lo ||
                hi - lo > table->slot_stride) {
                return -1;
            }
        }
    }
    flush_pending(state);  // what has been collected goes where it was meant to go
    state->external_reduce = table != nullptr;
    state->reduce_calls_seen = 0;
    if (table != nullptr) {
        state->reduce_table = *table;
        state->external_export = false;
    }
    state->mirror_valid = false;
    return 0;
}

Q_PLUGIN_EXPORT int Q_cuda_plugin_flush(Q_plugin_handle* handle, void* stream) {
    if (!handle) return -1;
    auto* state = reinterpret_cast<plugin_state*>(handle);
    if (cudaSetDevice(state->device) != cudaSuccess) return -1;
    flush_pending(state, static_cast<cudaStream_t>(stream));
    return state->last_launch_exported ? 1 : 0;
}

Q_PLUGIN_EXPORT int Q_cuda_plugin_mesh_view(Q_plugin_handle* handle, Q_cuda_mesh_view* out) {
    if (!handle || !out) return -1;
    auto* state = reinterpret_cast<plugin_state*>(handle);
    if (!state->use_mesh) return -1;
    std::memset(out, 0, sizeof *out);
    out->triangles = state->mesh.triangles.data();
    out->nodes = state->mesh.nodes.data();
    out->leaf_order = state->mesh.leaf_order.data();
    out->n_triangles = static_cast<uint32_t>(state->mesh.triangles.size());
    out->n_nodes = static_cast<uint32_t>(state->mesh.nodes.size());
    out->depth = state->mesh.depth;
    return 0;
}

Q_PLUGIN_EXPORT int Q_cuda_accumulate(const float* current, const float* history, float* output,
                                      uint64_t n_pixels, uint32_t frame_count, void* stream) {
    if (!current || !history || !output) return -1;
    return Q::cuda::launch_accumulate(current, history, output, n_pixels, frame_count,
                                      static_cast<cudaStream_t>(stream)) == cudaSuccess ? 0 : -1;
}

Q_PLUGIN_EXPORT int Q_cuda_tonemap_bgra8(const float* accumulated, uint32_t width,
                                         uint32_t height, void* bgra, uint32_t pitch_bytes,
                                         void* stream) {
    if (!accumulated || !bgra) return -1;
    if (pitch_bytes == 0) pitch_bytes = 4u * width;
    return Q::cuda::launch_tonemap_bgra8(accumulated, width, height, bgra, pitch_bytes,
                                         static_cast<cudaStream_t>(stream)) == cudaSuccess ? 0 : -1;
}

Q_PLUGIN_EXPORT int Q_cuda_combine_shards(const float* const* shards, uint32_t n_shards,
                                          uint64_t n_pixels, float scale, float* out, void* bgra,
                                          void* stream) {
    if (!shards || !out) return -1;
    return Q::cuda::launch_combine_shards(shards, n_shards, n_pixels, scale, out, bgra,
                                          static_cast<cudaStream_t>(stream)) == cudaSuccess ? 0 : -1;
}

Q_PLUGIN_EXPORT int Q_cuda_push_slices(const float* block, uint64_t n_float4, const Q_cuda_export_table* table,
                                       void* stream) {
    if (!block || !table || table->struct_size < sizeof(Q_cuda_export_table) || table->slices == 0 ||
        table->slices > Q_CUDA_MAX_EXPORT_SLICES || table->float4_per_slice == 0 || n_float4 >= (uint64_t{1} << 32)) {
        return -1;
    }
    uint32_t slice_lo[Q_CUDA_MAX_EXPORT_SLICES + 1];
    for (uint32_t i = 0; i <= table->slices; i++) {
        const uint64_t lo = static_cast<uint64_t>(i) * table->float4_per_slice;
        slice_lo[i] = lo < n_float4 && i < table->slices ? static_cast<uint32_t>(lo) : static_cast<uint32_t>(n_float4);
    }
    return Q::cuda::launch_push_slices(block, n_float4, table->base, slice_lo, table->slices,
                                       static_cast<cudaStream_t>(stream)) == cudaSuccess ? 0 : -1;
}

Q_PLUGIN_EXPORT int Q_cuda_push_slices_reduce(const float* block, uint64_t n_float4, const Q_cuda_reduce_table* table,
                                              void* stream) {
    if (!block || !table || table->struct_size < sizeof(Q_cuda_reduce_table) || table->ranks < 2 ||
        table->ranks > Q_CUDA_MAX_EXPORT_SLICES || table->rank >= table->ranks) {
        return -1;
    }
    void* base[Q_CUDA_MAX_EXPORT_SLICES];
    for (uint32_t i = 0; i < table->ranks; i++) {
        if (table->staging[i] == nullptr) return -1;
        base[i] = static_cast<float4*>(table->staging[i]) + static_cast<size_t>(table->rank) * table->slot_stride;
    }
    return Q::cuda::launch_push_slices(block, n_float4, base, table->slice_lo, table->ranks,
                                       static_cast<cudaStream_t>(stream)) == cudaSuccess ? 0 : -1;
}

Q_PLUGIN_EXPORT int Q_cuda_reduce_slices_sync(const float* staging, uint32_t float4_per_slice, uint32_t* const* flags,
                                              uint32_t n_shards, uint32_t rank, uint32_t epoch, uint64_t n_float4,
                                              float scale, float* out, void* stream) {
    if (!staging || !flags || !out || epoch == 0) return -1;
    return Q::cuda::launch_reduce_slices_sync(staging, float4_per_slice, flags, n_shards, rank, epoch, n_float4, scale, out,
                                              static_cast<cudaStream_t>(stream)) == cudaSuccess ? 0 : -1;
}

Q_PLUGIN_EXPORT int Q_cuda_host_register(void* host_ptr, uint64_t bytes, void** out_device_ptr) {
    if (!host_ptr || !out_device_ptr || bytes == 0) return -1;
    *out_device_ptr = nullptr;
    if (cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1;
    }
    if (cudaHostGetDevicePointer(out_device_ptr, host_ptr, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        cudaHostUnregister(host_ptr);
        return -1;
    }
    return 0;
}

Q_PLUGIN_EXPORT int Q_cuda_host_unregister(void* host_ptr) {
    if (!host_ptr) return 0;
    if (cudaHostUnregister(host_ptr) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1;
    }
    return 0;
}

Q_PLUGIN_EXPORT int Q_cuda_memset_async(void* device_ptr, int value, uint64_t bytes, void* stream) {
    if (!device_ptr) return -1;
    return cudaMemsetAsync(device_ptr, value, bytes, static_cast<cudaStream_t>(stream)) == cudaSuccess ? 0 : -1;
}

Q_PLUGIN_EXPORT int Q_cuda_memcpy_async(void* dst, const void* src, uint64_t bytes, void* stream) {
    if (!dst || !src) return -1;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)) == cudaSuccess ? 0 : -1;
}

Q_PLUGIN_EXPORT int Q_cuda_probe_sincos(const float* phi, float* sin_out, float* cos_out,
                                        uint64_t n, void* stream) {
    if (!phi || !sin_out || !cos_out) return -1;
    return Q::cuda::launch_sincos_probe(phi, sin_out, cos_out, n, static_cast<cudaStream_t>(stream)) ==
                   cudaSuccess ? 0 : -1;
}

Q_PLUGIN_EXPORT int Q_cuda_alloc(uint64_t bytes, void** out_device_ptr) {
    if (!out_device_ptr) return -1;
    *out_device_ptr = nullptr;
    if (cudaMalloc(out_device_ptr, bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1;
    }
    return 0;
}

Q_PLUGIN_EXPORT int Q_cuda_free(void* device_ptr) {
    if (!device_ptr) return 0;
    return cudaFree(device_ptr) == cudaSuccess ? 0 : -1;
}

Q_PLUGIN_EXPORT int Q_cuda_ipc_export(void* device_ptr, Q_cuda_ipc_handle* out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(Q_cuda_ipc_handle));
    if (!device_ptr || !out) return -1;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, device_ptr) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1;
    }
    std::memcpy(out->bytes, &h, sizeof h);
    return 0;
}

Q_PLUGIN_EXPORT int Q_cuda_ipc_open(const Q_cuda_ipc_handle* handle, void** out_device_ptr) {
    if (!handle || !out_device_ptr) return -1;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle->bytes, sizeof h);
    *out_device_ptr = nullptr;
    if (cudaIpcOpenMemHandle(out_device_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1;
    }
    return 0;
}

Q_PLUGIN_EXPORT int Q_cuda_ipc_close(void* device_ptr) {
    if (!device_ptr) return 0;
    return cudaIpcCloseMemHandle(device_ptr) == cudaSuccess ? 0 : -1;
}

}  // extern "C"
