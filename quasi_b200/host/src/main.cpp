/// @file main.cpp
/// @brief Headless Linux host: loads a backend plugin and renders to EXR.
///
/// Behaviour follows the reference host's batch mode
/// (/root/reference/src/quasi/host/main.cpp:120-327): positional plugin path,
/// `--render N` renders N frames (N samples per pixel) and saves an EXR with
/// AOVs; the per-frame order is update -> begin_frame -> fill camera -> render ->
/// end_frame; the camera is the orbit controller's default (main.cpp:37-45).
/// There is no window: GLFW/Cocoa (window.cpp) are replaced by nothing, the Metal
/// context by Q::gpu::cuda::context.
///
/// Unlike the reference host, this one actually uses the pieces the reference
/// ships as libraries only: the plugin is owned by
/// Q::plugin::manager, and the frame loop of every GPU is a coroutine on one
/// Q::async::scheduler, so `--watch` hot-reloads the backend mid-run and
/// `--gpus G` drives G GPUs (one stream set each, sample-sharded) from one thread.

#include <quasi/gpu/cuda/context.hpp>
#include <quasi/io/exr_writer.hpp>
#include <quasi/plugin/plugin.hpp>

#include <cuda_runtime_api.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <memory>
#include <string>
#include <string_view>
#include <vector>

namespace {

using clock_type = std::chrono::steady_clock;

void plugin_log(void* /*host_data*/, const char* message) { std::printf("[Plugin] %s\n", message); }

bool g_shutdown_requested = false;
void plugin_request_shutdown(void* /*host_data*/) { g_shutdown_requested = true; }

/// The orbit controller of main.cpp:37-116 without the mouse: default pose only.
struct orbit_camera {
    float target[3] = {0.0f, 1.0f, 0.0f};
    float distance = 3.5f;
    float azimuth = 0.0f;
    float elevation = 0.0f;
    float fov = 40.0f;
    float position[3] = {0.0f, 1.0f, 3.5f};
    bool dirty = true;

    void update_position() {
        constexpr float max_elev = 1.5f;
        elevation = std::clamp(elevation, -max_elev, max_elev);
        float cos_elev = std::cos(elevation);
        position[0] = target[0] + distance * std::sin(azimuth) * cos_elev;
        position[1] = target[1] + distance * std::sin(elevation);
        position[2] = target[2] + distance * std::cos(azimuth) * cos_elev;
    }

    void fill_camera(Q_camera& cam) const {
        for (int i = 0; i < 3; i++) {
            cam.position[i] = position[i];
            cam.target[i] = target[i];
        }
        cam.up[0] = 0.0f;
        cam.up[1] = 1.0f;
        cam.up[2] = 0.0f;
        cam.fov = fov;
    }
};

struct options {
    std::filesystem::path plugin_path;
    std::filesystem::path out_path;
    int render_frames = 0;       ///< --render N; 0 with --progressive = run by time
    double progressive_s = 0.0;  ///< --progressive SECONDS: 1 spp per frame + tonemap, by wall time
    uint32_t width = 720, height = 720;  ///< the reference window is 720x720 (main.cpp:155)
    int gpus = 1;
    int first_device = 0;
    bool watch = false;
    bool present = false;
    bool json = false;
    bool keep_accumulation = false;  ///< --keep-accumulation: a reloaded backend continues the image
    int repeat = 1;                  ///< --repeat R: render the N-frame set R times (accumulation restarted each time)
    bool readback_each = false;      ///< --readback-each: every set ends with the blocking AOV readback (end to end)
};

struct gpu_worker {
    int index = 0;
    Q::gpu::cuda::context context;
    std::unique_ptr<Q::plugin::manager> manager;
    orbit_camera camera;
    uint64_t frames = 0;
    bool finished = false;
    std::vector<float> frame_gap_ms;  ///< wall time between consecutive frame submissions
    clock_type::time_point last_submit{};
    std::vector<float> set_ms;        ///< --repeat: wall time of every N-frame set, first submission to idle (+ readback)
};

/// The main loop of main.cpp:243-323 for one GPU, as a coroutine: it never blocks,
/// so one scheduler thread keeps every GPU and the reload watcher going.
Q::async::task<void> frame_loop(gpu_worker& w, const options& opt, clock_type::time_point deadline) {
    auto last_time = clock_type::now();
    auto set_start = last_time;
    uint64_t frames_in_set = 0;
    int sets_done = 0;
    for (;;) {
        if (g_shutdown_requested) break;
        if (opt.render_frames > 0 && frames_in_set >= static_cast<uint64_t>(opt.render_frames)) {
            // One frame set is complete: wait for the GPU (and, with --readback-each, bring the four layers to the
            // host exactly as the save path does), note the wall time, start the next set from sample 0.
            co_await w.context.wait_idle();
            if (opt.readback_each) {
                if (auto* plugin = w.manager->plugin(); plugin && plugin->supports_readback_aov()) {
                    auto rb = plugin->readback_aov();
                    plugin->readback_aov_free(&rb);
                }
            }
            w.set_ms.push_back(std::chrono::duration<float, std::milli>(clock_type::now() - set_start).count());
            if (++sets_done >= opt.repeat) break;
            frames_in_set = 0;
            w.camera.dirty = true;
            set_start = clock_type::now();
            continue;
        }
        if (opt.progressive_s > 0.0 && clock_type::now() >= deadline) break;

        if (!w.manager->is_loaded() || w.manager->is_reloading()) {
            // Mid-reload: submit nothing, so that the pre_unload hook sees the GPU
            // go idle and the old library can be closed.
            co_await Q::async::yield();
            continue;
        }
        co_await w.context.wait_slot();

        auto now = clock_type::now();
        float delta_time = std::chrono::duration<float>(now - last_time).count();
        last_time = now;
        w.manager->update(delta_time);

        auto frame = w.context.begin_frame();
        if (frame) {
            w.camera.fill_camera(frame->camera);
            frame->camera_dirty = w.camera.dirty ? 1u : 0u;
            w.camera.dirty = false;
            w.manager->render(&*frame);
            w.context.end_frame(*frame);
            ++w.frames;
            ++frames_in_set;
            if (w.last_submit != clock_type::time_point{}) {
                w.frame_gap_ms.push_back(std::chrono::duration<float, std::milli>(now - w.last_submit).count());
            }
            w.last_submit = now;
        }
        co_await Q::async::yield();
    }
    co_await w.context.wait_idle();
    w.finished = true;
}

float percentile(std::vector<float> v, double p) {
    if (v.empty()) return 0.0f;
    std::sort(v.begin(), v.end());
    size_t i = static_cast<size_t>(p * static_cast<double>(v.size() - 1) + 0.5);
    return v[std::min(i, v.size() - 1)];
}

void usage(const char* argv0) {
    std::fprintf(stderr,
                 "Usage: %s [--render N] [--progressive SECONDS] [--width W] [--height H] [--out FILE.exr]\n"
                 "          [--gpus G] [--device D] [--watch] [--keep-accumulation] [--present] [--json]\n"
                 "          [--repeat R] [--readback-each] <plugin.so>\n",
                 argv0);
}

bool parse(int argc, char* argv[], options& opt) {
    for (int i = 1; i < argc; ++i) {
        std::string_view arg = argv[i];
        auto value = [&](const char* name) -> const char* {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "%s needs a value\n", name);
                return nullptr;
            }
            return argv[++i];
        };
        if (arg == "--render") {
            const char* v = value("--render");
            if (!v) return false;
            opt.render_frames = std::atoi(v);
        } else if (arg == "--progressive") {
            const char* v = value("--progressive");
            if (!v) return false;
            opt.progressive_s = std::atof(v);
        } else if (arg == "--width") {
            const char* v = value("--width");
            if (!v) return false;
            opt.width = static_cast<uint32_t>(std::atoi(v));
        } else if (arg == "--height") {
            const char* v = value("--height");
            if (!v) return false;
            opt.height = static_cast<uint32_t>(std::atoi(v));
        } else if (arg == "--out") {
            const char* v = value("--out");
            if (!v) return false;
            opt.out_path = v;
        } else if (arg == "--gpus") {
            const char* v = value("--gpus");
            if (!v) return false;
            opt.gpus = std::max(1, std::atoi(v));
        } else if (arg == "--device") {
            const char* v = value("--device");
            if (!v) return false;
            opt.first_device = std::atoi(v);
        } else if (arg == "--watch") {
            opt.watch = true;
        } else if (arg == "--present") {
            opt.present = true;
        } else if (arg == "--json") {
            opt.json = true;
        } else if (arg == "--keep-accumulation") {
            opt.keep_accumulation = true;
        } else if (arg == "--repeat") {
            const char* v = value("--repeat");
            if (!v) return false;
            opt.repeat = std::max(1, std::atoi(v));
        } else if (arg == "--readback-each") {
            opt.readback_each = true;
        } else if (!arg.empty() && arg[0] != '-') {
            opt.plugin_path = arg;
        } else {
            std::fprintf(stderr, "Unknown option: %s\n", argv[i]);
            return false;
        }
    }
    return true;
}

/// Sample-shard combine in one process: rank-ordered mean of the per-GPU layer
/// blocks on GPU 0 via peer access, then one D2H copy. Returns host floats
/// (4 layers x width x height x 4), empty on failure.
std::vector<float> combine_on_root(std::vector<std::unique_ptr<gpu_worker>>& workers, uint32_t width,
                                   uint32_t height) {
    using layers_fn = int (*)(Q_plugin_handle*, void*, Q_cuda_layers*);
    using combine_fn = int (*)(const float* const*, uint32_t, uint64_t, float, float*, void*, void*);
    const uint64_t n_float4 = static_cast<uint64_t>(width) * height * 4;  // four contiguous layers
    std::vector<const float*> blocks;
    for (auto& w : workers) {
        // Device pointers of this GPU's contiguous 4-layer block, ordered after its renders.
        auto get_layers = w->manager->extension<layers_fn>("Q_cuda_plugin_layers");
        auto* plugin = w->manager->plugin();
        if (!get_layers || !plugin) return {};
        Q_cuda_layers l{};
        if (get_layers(plugin->native_handle(), w->context.queue(), &l) != 0 || !l.device_ptr[0]) return {};
        blocks.push_back(l.device_ptr[0]);
    }
    auto combine = workers[0]->manager->extension<combine_fn>("Q_cuda_combine_shards");
    if (!combine) return {};

    const int root = workers[0]->context.device_ordinal();
    cudaSetDevice(root);
    for (size_t g = 1; g < workers.size(); g++) {
        cudaError_t e = cudaDeviceEnablePeerAccess(workers[g]->context.device_ordinal(), 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            std::fprintf(stderr, "[Host] no peer access %d -> %d: %s\n", root,
                         workers[g]->context.device_ordinal(), cudaGetErrorString(e));
            return {};
        }
        (void)cudaGetLastError();
    }
    float* result = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&result), n_float4 * 16) != cudaSuccess) return {};
    std::vector<float> host(n_float4 * 4);
    bool ok = combine(blocks.data(), static_cast<uint32_t>(blocks.size()), n_float4, 0.0f, result, nullptr,
                      workers[0]->context.queue()) == 0;
    ok = ok && cudaMemcpyAsync(host.data(), result, n_float4 * 16, cudaMemcpyDeviceToHost,
                               static_cast<cudaStream_t>(workers[0]->context.queue())) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(static_cast<cudaStream_t>(workers[0]->context.queue())) == cudaSuccess;
    cudaFree(result);
    if (!ok) host.clear();
    return host;
}

/// The sample-shard exchange of one frame as part of the render (quasi_cuda_backend.h: Q_cuda_plugin_set_reduce),
/// for several GPUs driven by this process: the layer block is cut into G slices along rows of 8x4 tiles, GPU s owns
/// slice s (a staging area of G slots and an arrival counter per tile and layer, in its own memory, reached by the
/// others through peer access), and the batched render kernels store every finished tile at its owners, count its
/// arrivals and, where an arrival is the G-th, sum the G copies in rank order and store the finished tile into
/// `frame`, pinned host memory every GPU has mapped. When the last GPU's last launch ends, the frame is on the host:
/// no combine pass, no copy. One frame per set-up (--repeat uses combine_on_root).
struct fused_exchange {
    std::vector<int> devices;
    std::vector<void*> staging, counters;
    float* frame = nullptr;
    uint64_t frame_floats = 0;
    bool active = false;

    /// slice_lo[0..G] (quasi_b200/multi_gpu.py:tile_row_slices): 4 * ceil(h / 4) rows of tiles over the four layers,
    /// slice s starts at row ceil(s * rows / G).
    static std::vector<uint32_t> slices(uint32_t w, uint32_t h, uint32_t ranks) {
        const uint32_t tiles_y = (h + 3u) / 4u, rows = 4u * tiles_y;
        std::vector<uint32_t> lo(ranks + 1u);
        for (uint32_t s = 0; s <= ranks; s++) {
            const uint32_t row = (s * rows + ranks - 1u) / ranks;
            lo[s] = s == ranks ? 4u * w * h : ((row / tiles_y) * h + (row % tiles_y) * 4u) * w;
        }
        return lo;
    }

    bool setup(std::vector<std::unique_ptr<gpu_worker>>& workers, uint32_t w, uint32_t h, uint32_t render_calls) {
        using set_reduce_fn = int (*)(Q_plugin_handle*, const Q_cuda_reduce_table*);
        const uint32_t ranks = static_cast<uint32_t>(workers.size());
        if (ranks < 2 || ranks > Q_CUDA_MAX_EXPORT_SLICES || w == 0 || h == 0) return false;
        for (auto& wk : workers) devices.push_back(wk->context.device_ordinal());
        for (int a : devices) {
            if (cudaSetDevice(a) != cudaSuccess) return false;
            for (int b : devices) {
                if (a == b) continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    (void)cudaGetLastError();
                    return false;
                }
                (void)cudaGetLastError();
            }
        }
        const std::vector<uint32_t> lo = slices(w, h, ranks);
        uint32_t stride = 1;
        for (uint32_t s = 0; s < ranks; s++) stride = std::max(stride, lo[s + 1] - lo[s]);
        const size_t n_tiles = static_cast<size_t>((w + 7u) / 8u) * ((h + 3u) / 4u);
        frame_floats = static_cast<uint64_t>(w) * h * 16u;
        if (cudaHostAlloc(reinterpret_cast<void**>(&frame), frame_floats * sizeof(float),
                          cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            (void)cudaGetLastError();
            frame = nullptr;
            return false;
        }
        for (int d : devices) {
            void *st = nullptr, *ct = nullptr;
            if (cudaSetDevice(d) != cudaSuccess || cudaMalloc(&st, static_cast<size_t>(ranks) * stride * 16u) != cudaSuccess ||
                cudaMalloc(&ct, 4u * n_tiles * sizeof(uint32_t)) != cudaSuccess ||
                cudaMemset(ct, 0, 4u * n_tiles * sizeof(uint32_t)) != cudaSuccess) {
                (void)cudaGetLastError();
                if (st) cudaFree(st);
                if (ct) cudaFree(ct);
                return false;
            }
            staging.push_back(st);
            counters.push_back(ct);
        }
        for (uint32_t g = 0; g < ranks; g++) {
            auto set_reduce = workers[g]->manager->extension<set_reduce_fn>("Q_cuda_plugin_set_reduce");
            auto* plugin = workers[g]->manager->plugin();
            if (!set_reduce || !plugin) return false;
            Q_cuda_reduce_table t{};
            t.struct_size = sizeof t;
            t.ranks = ranks;
            t.rank = g;
            t.slot_stride = stride;
            t.render_calls = render_calls;
            for (uint32_t s = 0; s <= ranks; s++) t.slice_lo[s] = lo[s];
            for (uint32_t s = 0; s < ranks; s++) {
                t.staging[s] = staging[s];
                t.counters[s] = static_cast<uint32_t*>(counters[s]);
                t.out[s] = frame + static_cast<size_t>(lo[s]) * 4u;
            }
            if (set_reduce(plugin->native_handle(), &t) != 0) return false;
        }
        active = true;
        return true;
    }

    /// After the frame's render calls: launches what the plugins still hold; true if every GPU's last launch carried
    /// the exchange, i.e. `frame` is the finished frame once the GPUs are idle (they are waited for here).
    bool finish(std::vector<std::unique_ptr<gpu_worker>>& workers) {
        using flush_fn = int (*)(Q_plugin_handle*, void*);
        bool all = active;
        for (auto& wk : workers) {
            auto flush = wk->manager->extension<flush_fn>("Q_cuda_plugin_flush");
            auto* plugin = wk->manager->plugin();
            if (!flush || !plugin || flush(plugin->native_handle(), wk->context.queue()) != 1) all = false;
        }
        for (int d : devices) {
            cudaSetDevice(d);
            cudaDeviceSynchronize();
        }
        return all;
    }

    void release(std::vector<std::unique_ptr<gpu_worker>>& workers) {
        using set_reduce_fn = int (*)(Q_plugin_handle*, const Q_cuda_reduce_table*);
        for (auto& wk : workers) {
            auto set_reduce = wk->manager->extension<set_reduce_fn>("Q_cuda_plugin_set_reduce");
            if (auto* plugin = wk->manager->plugin(); set_reduce && plugin) (void)set_reduce(plugin->native_handle(), nullptr);
        }
        for (size_t i = 0; i < staging.size(); i++) {
            cudaSetDevice(devices[i]);
            cudaDeviceSynchronize();
            cudaFree(staging[i]);
            cudaFree(counters[i]);
        }
        staging.clear();
        counters.clear();
        if (frame) cudaFreeHost(frame);
        frame = nullptr;
        active = false;
    }
};

}  // namespace

int main(int argc, char* argv[]) {
    options opt;
    if (!parse(argc, argv, opt)) {
        usage(argv[0]);
        return EXIT_FAILURE;
    }
    if (opt.plugin_path.empty()) {
        if (const char* env = std::getenv("QUASI_PLUGIN")) opt.plugin_path = env;
    }
    if (opt.plugin_path.empty()) {
        std::fprintf(stderr, "No plugin given\n");
        usage(argv[0]);
        return EXIT_FAILURE;
    }
    if (opt.render_frames <= 0 && opt.progressive_s <= 0.0) opt.render_frames = 1;
    if (opt.progressive_s > 0.0) opt.present = true;  // interactive mode tonemaps every frame

    const int available = Q::gpu::cuda::device_count();
    if (available <= 0) {
        std::fprintf(stderr, "Failed to create CUDA context: %s\n",
                     Q::gpu::cuda::to_string(Q::gpu::cuda::context_error::device_not_found));
        return EXIT_FAILURE;
    }
    if (opt.first_device + opt.gpus > available) {
        std::fprintf(stderr, "Requested %d GPU(s) from device %d, %d available\n", opt.gpus, opt.first_device,
                     available);
        return EXIT_FAILURE;
    }

    // One context + one plugin instance per GPU.
    std::vector<std::unique_ptr<gpu_worker>> workers;
    for (int g = 0; g < opt.gpus; g++) {
        auto w = std::make_unique<gpu_worker>();
        w->index = g;
        auto ctx = Q::gpu::cuda::context::create(
            opt.width, opt.height,
            {.device_ordinal = opt.first_device + g, .shard_index = static_cast<uint32_t>(g),
             .shard_count = static_cast<uint32_t>(opt.gpus), .present = opt.present,
             .keep_accumulation = opt.keep_accumulation});
        if (!ctx) {
            std::fprintf(stderr, "Failed to create CUDA context: %s\n", Q::gpu::cuda::to_string(ctx.error()));
            return EXIT_FAILURE;
        }
        w->context = std::move(*ctx);
        w->camera.update_position();

        gpu_worker* raw = w.get();
        Q::plugin::reload_hooks hooks;
        hooks.pre_unload = [raw]() -> Q::async::task<void> {
            co_await raw->context.wait_idle();  // nothing of the old library may still be running
        };
        hooks.post_load = [raw, keep = opt.keep_accumulation]() -> Q::async::task<void> {
            // The new instance starts from sample 0 (the reference's behaviour), unless the
            // context offered it the old instance's layers (Q_cuda_persistent_state).
            if (!keep) raw->camera.dirty = true;
            co_return;
        };
        hooks.on_error = [](const std::string& msg) { std::fprintf(stderr, "[Host] reload error: %s\n", msg.c_str()); };
        w->manager = std::make_unique<Q::plugin::manager>(opt.plugin_path, std::move(hooks));
        w->manager->set_viewport(opt.width, opt.height);
        w->manager->set_host_data(nullptr);
        w->manager->set_gpu_context(w->context.gpu());
        w->manager->set_log_callback(plugin_log);
        w->manager->set_shutdown_callback(plugin_request_shutdown);
        auto loaded = w->manager->load_sync();
        if (!loaded) {
            std::fprintf(stderr, "Failed to load plugin: %s\n",
                         std::string{Q::plugin::to_string(loaded.error())}.c_str());
            return EXIT_FAILURE;
        }
        workers.push_back(std::move(w));
    }
    // With several GPUs the frame is combined on GPU 0 from the layers themselves; nobody reads a single GPU's
    // frame back, so its readback mirror (the batched launches' export into pinned host memory) stays off.
    // (After a hot reload the new instance has its mirror on again until the next run: a cost, not an error.)
    auto disable_mirrors = [&] {
        if (workers.size() < 2) return;
        using set_export_fn = int (*)(Q_plugin_handle*, const Q_cuda_export_table*);
        for (auto& w : workers) {
            auto set_export = w->manager->extension<set_export_fn>("Q_cuda_plugin_set_export");
            auto* plugin = w->manager->plugin();
            if (!set_export || !plugin) continue;
            Q_cuda_export_table none{};
            none.struct_size = sizeof none;
            (void)set_export(plugin->native_handle(), &none);
        }
    };
    disable_mirrors();
    // One frame on several GPUs: the exchange rides inside the render (fused_exchange); QUASI_HOST_EXCHANGE=pull keeps
    // the combine pass on GPU 0 after the render, which is also what any failure to set the fused form up falls back to.
    fused_exchange exchange;
    const char* exchange_mode = std::getenv("QUASI_HOST_EXCHANGE");
    if (workers.size() > 1 && opt.render_frames > 0 && opt.repeat == 1 && !opt.watch && !opt.present &&
        !(exchange_mode && std::strcmp(exchange_mode, "pull") == 0)) {
        if (exchange.setup(workers, opt.width, opt.height, static_cast<uint32_t>(opt.render_frames))) {
            std::printf("[Host] Shard exchange fused into the render kernels (%zu GPUs)\n", workers.size());
        } else {
            exchange.release(workers);
            disable_mirrors();
        }
    }
    if (auto info = workers[0]->manager->info()) {
        std::printf("Loaded plugin: %s v%u.%u.%u\n", info->name, info->version.major, info->version.minor,
                    info->version.patch);
        std::printf("Description: %s\n", info->description);
    }
    if (opt.render_frames > 0) {
        std::printf("[Host] Batch mode: rendering %d frames then saving EXR (%ux%u, %d GPU%s)\n",
                    opt.render_frames, opt.width, opt.height, opt.gpus, opt.gpus > 1 ? "s" : "");
    }

    // Everything below runs on one scheduler thread.
    Q::async::scheduler scheduler;
    const auto start = clock_type::now();
    const auto deadline = start + std::chrono::duration_cast<clock_type::duration>(
                                      std::chrono::duration<double>(opt.progressive_s));
    for (auto& w : workers) scheduler.spawn(frame_loop(*w, opt, deadline));
    if (opt.watch) {
        for (auto& w : workers) scheduler.spawn(w->manager->watch_and_reload_loop());
    }
    auto all_finished = [&] {
        return std::all_of(workers.begin(), workers.end(), [](const auto& w) { return w->finished; });
    };
    while (!all_finished()) scheduler.tick();

    // Statistics (Mrays/s needs the backend's ray counter; the reference has none).
    uint64_t rays = 0, paths = 0, frames = 0, launches = 0;
    for (auto& w : workers) {
        Q_cuda_stats s{};
        if (auto* p = w->manager->plugin(); p && p->cuda_stats(&s)) {
            rays += s.rays;
            paths += s.paths;
            launches += s.kernel_launches;
        }
        frames += w->frames;
    }
    // (taken after the statistics: asking for them launches the render calls a plugin still holds and waits for them)
    const double elapsed_s = std::chrono::duration<double>(clock_type::now() - start).count();
    std::printf("[Host] %llu frames, %llu paths, %llu rays in %.3f s: %.1f Mrays/s, %.3f ms/frame\n",
                static_cast<unsigned long long>(frames), static_cast<unsigned long long>(paths),
                static_cast<unsigned long long>(rays), elapsed_s, static_cast<double>(rays) / elapsed_s / 1e6,
                frames ? elapsed_s * 1e3 / static_cast<double>(workers[0]->frames) : 0.0);

    // Save (main.cpp:277-321): AOV readback when supported, else beauty only.
    int exit_code = EXIT_SUCCESS;
    auto out_path = opt.out_path.empty() ? Q::io::make_timestamped_path(".") : opt.out_path;
    if (opt.render_frames > 0) {
        if (workers.size() == 1) {
            auto* plugin = workers[0]->manager->plugin();
            if (plugin && plugin->supports_readback_aov()) {
                std::printf("[Host] Saving EXR with AOVs (%llu samples)...\n",
                            static_cast<unsigned long long>(workers[0]->frames));
                auto rb = plugin->readback_aov();
                if (rb.buffers[Q_AOV_BEAUTY].data) {
                    auto written = Q::io::write_exr(out_path, rb);
                    if (written) {
                        std::printf("[Host] Saved: %s\n", out_path.c_str());
                    } else {
                        std::fprintf(stderr, "[Host] EXR write failed: %s\n", Q::io::to_string(written.error()));
                        exit_code = EXIT_FAILURE;
                    }
                    plugin->readback_aov_free(&rb);
                } else {
                    std::fprintf(stderr, "[Host] AOV readback returned no data\n");
                    exit_code = EXIT_FAILURE;
                }
            } else if (plugin && plugin->supports_readback()) {
                auto rb = plugin->readback();
                if (rb.data) {
                    auto written = Q::io::write_exr(out_path, rb);
                    if (written) std::printf("[Host] Saved: %s\n", out_path.c_str());
                    plugin->readback_free(&rb);
                } else {
                    std::fprintf(stderr, "[Host] Readback returned no data\n");
                    exit_code = EXIT_FAILURE;
                }
            } else {
                std::printf("[Host] Plugin does not support HDR readback\n");
            }
        } else {
            std::vector<float> host;
            if (exchange.active && exchange.finish(workers)) {
                std::printf("[Host] Frame of %zu sample shards reduced inside the render kernels\n", workers.size());
                host.assign(exchange.frame, exchange.frame + exchange.frame_floats);
            } else {
                if (exchange.active) {  // e.g. a last launch of a single sample: nothing was exported
                    exchange.release(workers);
                    disable_mirrors();
                }
                std::printf("[Host] Combining %zu sample shards on GPU %d...\n", workers.size(),
                            workers[0]->context.device_ordinal());
                host = combine_on_root(workers, opt.width, opt.height);
            }
            if (host.empty()) {
                std::fprintf(stderr, "[Host] Shard combine failed\n");
                exit_code = EXIT_FAILURE;
            } else {
                const size_t layer = static_cast<size_t>(opt.width) * opt.height * 4;
                Q_readback_aov_result rb{};
                for (int i = 0; i < Q_AOV_COUNT; i++) {
                    rb.buffers[i] = {host.data() + layer * static_cast<size_t>(i), opt.width, opt.height, 4};
                }
                auto written = Q::io::write_exr(out_path, rb);
                if (written) {
                    std::printf("[Host] Saved: %s\n", out_path.c_str());
                } else {
                    exit_code = EXIT_FAILURE;
                }
            }
        }
    }

    if (opt.json || opt.progressive_s > 0.0) {
        const auto& gaps = workers[0]->frame_gap_ms;
        const auto& st = workers[0]->manager->stats();
        std::string set_ms_list;  // wall time of every frame set of GPU 0 (--repeat)
        for (float ms : workers[0]->set_ms) {
            char item[32];
            std::snprintf(item, sizeof item, "%s%.4f", set_ms_list.empty() ? "" : ", ", ms);
            set_ms_list += item;
        }
        std::printf(
            "{\"frames\": %llu, \"seconds\": %.4f, \"rays\": %llu, \"mrays_per_s\": %.2f, \"kernel_launches\": %llu, "
            "\"frame_ms_p50\": %.4f, \"frame_ms_p95\": %.4f, \"frame_ms_p99\": %.4f, \"frame_ms_max\": %.3f, "
            "\"reloads\": %llu, \"reload_failures\": %llu, \"last_reload_ms\": %.2f, \"gpus\": %d, "
            "\"width\": %u, \"height\": %u, \"set_ms\": [%s]}\n",
            static_cast<unsigned long long>(frames), elapsed_s, static_cast<unsigned long long>(rays),
            static_cast<double>(rays) / elapsed_s / 1e6, static_cast<unsigned long long>(launches),
            percentile(gaps, 0.50), percentile(gaps, 0.95), percentile(gaps, 0.99),
            gaps.empty() ? 0.0f : *std::max_element(gaps.begin(), gaps.end()),
            static_cast<unsigned long long>(st.success_count), static_cast<unsigned long long>(st.failure_count),
            st.last_reload_time * 1000.0f, opt.gpus, opt.width, opt.height, set_ms_list.c_str());
    }

    std::printf("Shutting down...\n");
    exchange.release(workers);
    workers.clear();  // managers destroy their plugins before the contexts go away
    return exit_code;
}
