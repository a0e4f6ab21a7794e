/// @file ref_loader_host.cpp
/// @brief Drop-in proof: libquasi_cuda.so driven by the REFERENCE's own, unmodified host classes.
///
/// Test infrastructure only. This TU is compiled with -I/root/reference/src and nothing of this
/// repository's host re-implementations: Q::plugin::dynamic_library
/// (/root/reference/src/quasi/plugin/dynamic_library.hpp:67-178), Q::plugin::loader
/// (loader.hpp:65-137: symbol resolution, ABI check, create) and Q::plugin::manager
/// (manager.hpp:122-135 load_sync, :145-157 render, :209-268 reload) are the reference's. The only
/// things it takes from this repository are what a host must know about a CUDA backend
/// (include/quasi_cuda_backend.h: the device descriptor behind Q_gpu_context::device) and the one
/// appended enumerator Q_GPU_BACKEND_CUDA = 4, which the reference's types.hpp does not have yet.
///
///   ref_loader_host <plugin.so> <width> <height> <frames> <out.bin>
///
/// Part 1 (loader): open, load, render <frames> calls on one stream exactly as main.cpp:243-270
/// does (update, fill the frame, render), read the four AOV layers back and write them as raw
/// float32 (beauty, albedo, normal, depth; each H x W x 4) to <out.bin>; the Python test compares
/// them with the oracle bit for bit.
/// Part 2 (manager): the same library under the reference's hot-reload manager: load_sync, four
/// render calls, one reload_async driven to completion, four more render calls. Prints what the
/// manager and the plugin logged so the test can check the sequence.

#include <quasi/plugin/dynamic_library.hpp>
#include <quasi/plugin/loader.hpp>
#include <quasi/plugin/manager.hpp>

#define QUASI_HOST_HAS_REFERENCE_HEADERS 1  // the ABI types come from the reference headers above
#include <quasi_cuda_backend.h>  // this repository: what Q_gpu_context::device means for CUDA

#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace {

int g_log_lines = 0;
void host_log(void*, const char* message) {
    g_log_lines++;
    std::printf("[Plugin] %s\n", message);
}

void fill_default_camera(Q_render_frame& frame) {  // the orbit controller's default pose, main.cpp:37-45
    const float position[3] = {0.0f, 1.0f, 3.5f}, target[3] = {0.0f, 1.0f, 0.0f}, up[3] = {0.0f, 1.0f, 0.0f};
    for (int i = 0; i < 3; i++) {
        frame.camera.position[i] = position[i];
        frame.camera.target[i] = target[i];
        frame.camera.up[i] = up[i];
    }
    frame.camera.fov = 40.0f;
}

}  // namespace

int main(int argc, char* argv[]) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s <plugin.so> <width> <height> <frames> <out.bin>\n", argv[0]);
        return 2;
    }
    const std::string plugin_path = argv[1];
    const uint32_t width = static_cast<uint32_t>(std::atoi(argv[2]));
    const uint32_t height = static_cast<uint32_t>(std::atoi(argv[3]));
    const int frames = std::atoi(argv[4]);
    const std::string out_path = argv[5];

    int devices = 0;
    if (cudaGetDeviceCount(&devices) != cudaSuccess || devices == 0) {
        std::fprintf(stderr, "no CUDA device\n");
        return 3;
    }
    cudaSetDevice(0);
    cudaStream_t queue = nullptr, frame_stream = nullptr;
    cudaStreamCreateWithFlags(&queue, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&frame_stream, cudaStreamNonBlocking);

    Q_cuda_device_desc desc{};
    desc.struct_size = sizeof desc;
    desc.device_ordinal = 0;
    desc.shard_index = 0;
    desc.shard_count = 1;
    Q::gpu::gpu_context gpu{};
    gpu.backend = Q_GPU_BACKEND_CUDA;  // = 4, appended to types.hpp:14-19 (quasi_cuda_backend.h)
    gpu.device = &desc;
    gpu.queue = queue;
    gpu.layer = nullptr;

    // ---- part 1: dynamic_library + loader, as the reference's main.cpp:206-232 uses them ----
    {
        auto lib = Q::plugin::dynamic_library::open(plugin_path);
        if (!lib) {
            std::fprintf(stderr, "dynamic_library::open failed: %s\n", Q::plugin::dynamic_library::last_error());
            return 4;
        }
        Q::plugin::plugin_context ctx{};
        ctx.viewport_width = width;
        ctx.viewport_height = height;
        ctx.host_data = nullptr;
        ctx.gpu = &gpu;
        ctx.log = host_log;
        ctx.request_shutdown = nullptr;
        auto plugin = Q::plugin::loader::load(*lib, &ctx);
        if (!plugin) {
            std::fprintf(stderr, "loader::load failed: %d\n", static_cast<int>(plugin.error()));
            return 5;
        }
        const auto info = plugin->info();
        std::printf("[Host] Loaded plugin: %s v%u.%u.%u (ABI %u)\n", info.name, info.version.major, info.version.minor,
                    info.version.patch, plugin->abi_version());
        for (int f = 0; f < frames; f++) {
            plugin->update(0.0f);
            Q::gpu::render_frame frame{};
            frame.drawable = nullptr;  // headless
            frame.command_buffer = frame_stream;
            frame.width = width;
            frame.height = height;
            fill_default_camera(frame);
            frame.camera_dirty = f == 0 ? 1u : 0u;
            plugin->render(&frame);
        }
        if (!plugin->supports_readback_aov()) {
            std::fprintf(stderr, "plugin has no AOV readback\n");
            return 6;
        }
        auto rb = plugin->readback_aov();
        if (!rb.buffers[0].data) {
            std::fprintf(stderr, "AOV readback returned no data\n");
            return 7;
        }
        std::FILE* out = std::fopen(out_path.c_str(), "wb");
        if (!out) return 8;
        for (int layer = 0; layer < 4; layer++) {
            const auto& b = rb.buffers[layer];
            if (b.width != width || b.height != height || b.channels != 4) {
                std::fprintf(stderr, "layer %d has the wrong shape\n", layer);
                return 9;
            }
            std::fwrite(b.data, sizeof(float), static_cast<size_t>(width) * height * 4, out);
        }
        std::fclose(out);
        plugin->readback_aov_free(&rb);
        std::printf("[Host] loader path: %d frames rendered, AOVs written to %s\n", frames, out_path.c_str());
        plugin->destroy();
    }

    // ---- part 2: the reference's hot-reload manager around the same library ----
    {
        int errors = 0;
        Q::plugin::reload_hooks hooks;
        hooks.pre_unload = [&]() -> Q::async::task<void> {
            cudaStreamSynchronize(frame_stream);  // nothing of the old library may still be running
            co_return;
        };
        hooks.on_error = [&](const std::string& msg) {
            errors++;
            std::fprintf(stderr, "[Host] reload error: %s\n", msg.c_str());
        };
        Q::plugin::manager manager{plugin_path, std::move(hooks)};
        manager.set_viewport(width, height);
        manager.set_host_data(nullptr);
        manager.set_gpu_context(&gpu);
        manager.set_log_callback(host_log);
        auto loaded = manager.load_sync();
        if (!loaded) {
            std::fprintf(stderr, "manager::load_sync failed: %s\n", std::string{Q::plugin::to_string(loaded.error())}.c_str());
            return 10;
        }
        auto render_some = [&](int n, bool restart) {
            for (int f = 0; f < n; f++) {
                manager.update(0.0f);
                Q::gpu::render_frame frame{};
                frame.command_buffer = frame_stream;
                frame.width = width;
                frame.height = height;
                fill_default_camera(frame);
                frame.camera_dirty = (restart && f == 0) ? 1u : 0u;
                manager.render(&frame);
            }
        };
        render_some(4, true);
        {
            auto reload = manager.reload_async();
            while (!reload.done()) reload.resume();
        }
        if (!manager.is_loaded()) {
            std::fprintf(stderr, "manager lost the plugin over the reload\n");
            return 11;
        }
        render_some(4, true);
        cudaStreamSynchronize(frame_stream);
        const auto& st = manager.stats();
        std::printf("[Host] manager path: reloads %llu ok %llu failed %llu errors %d log lines %d\n",
                    static_cast<unsigned long long>(st.reload_count), static_cast<unsigned long long>(st.success_count),
                    static_cast<unsigned long long>(st.failure_count), errors, g_log_lines);
        if (st.success_count != 1 || st.failure_count != 0 || errors != 0) return 12;
    }
    cudaStreamDestroy(frame_stream);
    cudaStreamDestroy(queue);
    std::printf("[Host] ok\n");
    return 0;
}
