// ref_driver.cpp - host loops around the reference's own shader and scene code.
//
// TEST INFRASTRUCTURE ONLY (see oracle/quasi_oracle.h). Built only where
// /root/reference exists, into oracle/_ref/libquasi_ref.so. It plays the part
// of the Metal runtime for the shader compiled in ref_msl_kernels.cpp: it
// launches one "fragment" per pixel and one "thread" per accumulate texel, and
// it packs GpuSceneUniforms the way Q_plugin_render does
// (/root/reference/backends/metal/plugin.mm:648-700) from the reference's own
// make_cornell_box() (/root/reference/src/quasi/scene/cornell_box.hpp:95-134),
// which is compiled from where it lies.
#include <quasi/scene/cornell_box.hpp>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {
unsigned ref_msl_sizeof_uniforms(void);
void ref_msl_pathtrace_frag(const void* uniforms, unsigned px, unsigned py, unsigned width,
                            unsigned height, float* out16);
void ref_msl_accumulate_row(const float* current, const float* history, float* output,
                            unsigned width, unsigned height, unsigned frame_count, unsigned gy);
void ref_msl_tonemap_frag(const float* accumulated, unsigned width, unsigned height,
                          unsigned px, unsigned py, float* out4);
}

namespace {

constexpr uint32_t k_max_quads = 32;

// Byte layout of GpuSceneUniforms, plugin.mm:354-387.
struct uniforms_block {
    float cam_position[3];
    float cam_fov;
    float cam_direction[3];
    float cam_aspect;
    float cam_up[3];
    float cam_pad;
    uint32_t quad_count;
    uint32_t frame_count;
    uint32_t light_index;
    float pad;
    float quads[k_max_quads][12];
    float materials[k_max_quads][8];
};
static_assert(sizeof(uniforms_block) == 2624);

int pick_threads(int threads, uint32_t height) {
    if (threads <= 0) threads = static_cast<int>(std::thread::hardware_concurrency());
    threads = std::max(1, std::min<int>(threads, static_cast<int>(height)));
    return threads;
}

template <class F>
void for_rows(uint32_t height, int threads, F&& body) {
    threads = pick_threads(threads, height);
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) {
        pool.emplace_back([&, t] {
            for (uint32_t y = static_cast<uint32_t>(t); y < height; y += threads) body(y);
        });
    }
    for (uint32_t y = 0; y < height; y += threads) body(y);
    for (auto& th : pool) th.join();
}

}  // namespace

#define REF_EXPORT extern "C" __attribute__((visibility("default")))

/// Scene block exactly as the Metal plugin would upload it for a viewport of
/// the given aspect (plugin.mm:582-584, 676-700). Camera and frame_count are
/// left for ref_pack_camera / the caller.
REF_EXPORT void ref_make_cornell_box(float aspect, void* out_uniforms) {
    auto* u = static_cast<uniforms_block*>(out_uniforms);
    Q::scene::cornell_box_scene scene = Q::scene::make_cornell_box(aspect);
    u->quad_count = static_cast<uint32_t>(std::min(scene.quads.size(), size_t(k_max_quads)));
    u->light_index = static_cast<uint32_t>(scene.light_index);
    std::memset(u->quads, 0, sizeof u->quads);
    std::memset(u->materials, 0, sizeof u->materials);
    for (size_t i = 0; i < u->quad_count; i++) {
        const auto& q = scene.quads[i];
        float* g = u->quads[i];
        g[0] = q.geometry.origin.x; g[1] = q.geometry.origin.y; g[2] = q.geometry.origin.z;
        g[4] = q.geometry.u.x; g[5] = q.geometry.u.y; g[6] = q.geometry.u.z;
        g[8] = q.geometry.v.x; g[9] = q.geometry.v.y; g[10] = q.geometry.v.z;
        float* m = u->materials[i];
        m[0] = q.mat.albedo.x; m[1] = q.mat.albedo.y; m[2] = q.mat.albedo.z;
        m[3] = q.mat.roughness;
        m[4] = q.mat.emission.x; m[5] = q.mat.emission.y; m[6] = q.mat.emission.z;
        m[7] = q.mat.metallic;
    }
}

/// Default camera of the reference scene (cornell_box.hpp:127-131), returned as
/// position[3], direction[3], up[3], fov, aspect.
REF_EXPORT void ref_cornell_box_camera(float aspect, float* out11) {
    Q::scene::cornell_box_scene scene = Q::scene::make_cornell_box(aspect);
    const auto& c = scene.cam;
    float v[11] = {c.position.x, c.position.y, c.position.z, c.direction.x, c.direction.y,
                   c.direction.z, c.up.x, c.up.y, c.up.z, c.fov, c.aspect};
    std::memcpy(out11, v, sizeof v);
}

/// Camera block from a host Q_camera, plugin.mm:648-674.
REF_EXPORT void ref_pack_camera(void* uniforms, const float* position, const float* target,
                                const float* up, float fov, uint32_t width, uint32_t height) {
    auto* u = static_cast<uniforms_block*>(uniforms);
    float aspect = static_cast<float>(width) / static_cast<float>(height);
    float dx = target[0] - position[0];
    float dy = target[1] - position[1];
    float dz = target[2] - position[2];
    float len = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (len > 0.0f) {
        dx /= len; dy /= len; dz /= len;
    }
    for (int i = 0; i < 3; i++) {
        u->cam_position[i] = position[i];
        u->cam_up[i] = up[i];
    }
    u->cam_direction[0] = dx;
    u->cam_direction[1] = dy;
    u->cam_direction[2] = dz;
    u->cam_fov = fov;
    u->cam_aspect = aspect;
    u->cam_pad = 0.0f;
}

/// PASS 1 of Q_plugin_render (plugin.mm:715-738): one pathtrace_frag per pixel
/// into the four MRT layers.
REF_EXPORT void ref_render_sample(const void* uniforms, uint32_t width, uint32_t height,
                                  float* beauty, float* albedo, float* normal, float* depth,
                                  int threads) {
    float* layer[4] = {beauty, albedo, normal, depth};
    for_rows(height, threads, [&](uint32_t py) {
        for (uint32_t px = 0; px < width; px++) {
            float f[16];
            ref_msl_pathtrace_frag(uniforms, px, py, width, height, f);
            size_t o = (static_cast<size_t>(py) * width + px) * 4;
            for (int l = 0; l < 4; l++) std::memcpy(layer[l] + o, f + l * 4, 16);
        }
    });
}

/// PASS 2 for one layer (plugin.mm:741-774): accumulate over a width x height grid.
REF_EXPORT void ref_accumulate(const float* current, const float* history, float* output,
                               uint32_t width, uint32_t height, uint32_t frame_count,
                               int threads) {
    for_rows(height, threads, [&](uint32_t gy) {
        ref_msl_accumulate_row(current, history, output, width, height, frame_count, gy);
    });
}

/// n_frames Q_plugin_render calls from a reset, ping-ponging like the plugin
/// (plugin.mm:703-712, 790-791). History before call 0 is zero. The RNG frame
/// index of call i is frame_first + i*frame_stride; the accumulate weight uses i.
REF_EXPORT void ref_render_frames(const void* uniforms, uint32_t width, uint32_t height,
                                  uint32_t frame_first, uint32_t frame_stride,
                                  uint32_t n_frames, float* const* layers, int threads) {
    size_t n = static_cast<size_t>(width) * height * 4;
    std::vector<float> cur[4], ping[4], pong[4];
    for (int l = 0; l < 4; l++) {
        cur[l].assign(n, 0.0f);
        ping[l].assign(n, 0.0f);
        pong[l].assign(n, 0.0f);
    }
    uniforms_block u;
    std::memcpy(&u, uniforms, sizeof u);
    bool flip = true;
    for (uint32_t i = 0; i < n_frames; i++) {
        u.frame_count = frame_first + i * frame_stride;
        ref_render_sample(&u, width, height, cur[0].data(), cur[1].data(), cur[2].data(),
                          cur[3].data(), threads);
        for (int l = 0; l < 4; l++) {
            float* rd = flip ? ping[l].data() : pong[l].data();
            float* wr = flip ? pong[l].data() : ping[l].data();
            ref_accumulate(cur[l].data(), rd, wr, width, height, i, threads);
        }
        flip = !flip;
    }
    for (int l = 0; l < 4; l++) {
        const float* last = flip ? ping[l].data() : pong[l].data();
        std::memcpy(layers[l], last, n * sizeof(float));
    }
}

/// PASS 3 (plugin.mm:777-788) into a BGRA8Unorm image: clamp, *255, round to
/// nearest even, as Metal's unorm conversion does.
REF_EXPORT void ref_tonemap_bgra8(const float* accumulated, uint32_t width, uint32_t height,
                                  uint8_t* bgra, int threads) {
    for_rows(height, threads, [&](uint32_t py) {
        for (uint32_t px = 0; px < width; px++) {
            float c[4];
            ref_msl_tonemap_frag(accumulated, width, height, px, py, c);
            uint8_t* o = bgra + (static_cast<size_t>(py) * width + px) * 4;
            auto q = [](float v) {
                v = std::fmin(std::fmax(v, 0.0f), 1.0f);
                return static_cast<uint8_t>(std::rint(v * 255.0f));
            };
            o[0] = q(c[2]);
            o[1] = q(c[1]);
            o[2] = q(c[0]);
            o[3] = q(c[3]);
        }
    });
}

REF_EXPORT int ref_hardware_threads(void) {
    return static_cast<int>(std::thread::hardware_concurrency());
}
