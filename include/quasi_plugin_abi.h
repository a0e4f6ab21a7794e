/*
 * quasi_plugin_abi.h - the C ABI between a Quasi host and a backend plugin.
 *
 * This is the drop-in boundary of libquasi_cuda.so. It declares, in plain C,
 * exactly the types and the ten entry points that the reference host binds
 * with dlsym (Q::plugin::loader, /root/reference/src/quasi/plugin/loader.hpp:
 * 72-128), so that libquasi_cuda.so loads wherever libquasi_metal.dylib does.
 * Layouts and names are fixed by the reference headers cited on each item;
 * only the wording here is ours. A C++ host may include the reference's own
 * plugin_interface.hpp instead: the two are layout-identical (checked by
 * tests/test_abi_layout.py).
 *
 * All calls come from the host's single main thread. No exceptions and no
 * error codes cross this boundary: create returns NULL on failure, every other
 * entry point tolerates NULL arguments silently.
 */
#ifndef QUASI_PLUGIN_ABI_H
#define QUASI_PLUGIN_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- GPU context types: /root/reference/src/quasi/gpu/types.hpp ---------- */

/* types.hpp:14-19. A uint32_t-sized enum in the reference. The CUDA value is
 * appended so that existing values keep their meaning. */
typedef uint32_t Q_gpu_backend;
enum {
    Q_GPU_BACKEND_NONE = 0,
    Q_GPU_BACKEND_METAL = 1,
    Q_GPU_BACKEND_VULKAN = 2,
    Q_GPU_BACKEND_WEBGPU = 3,
    Q_GPU_BACKEND_CUDA = 4 /* new; see quasi_cuda_backend.h for the handles */
};

/* types.hpp:37-42. Long-lived device resources, borrowed from the host. */
typedef struct Q_gpu_context {
    Q_gpu_backend backend;
    void* device; /* CUDA: const Q_cuda_device_desc*          */
    void* queue;  /* CUDA: cudaStream_t (the primary stream)   */
    void* layer;  /* CUDA: unused, may be NULL                 */
} Q_gpu_context;

/* types.hpp:45-50 */
typedef struct Q_camera {
    float position[3];
    float target[3];
    float up[3];
    float fov; /* vertical, degrees */
} Q_camera;

/* types.hpp:63-70. Valid for one frame only. */
typedef struct Q_render_frame {
    void* drawable;       /* CUDA: const Q_cuda_drawable*, NULL when headless */
    void* command_buffer; /* CUDA: cudaStream_t the frame is enqueued on      */
    uint32_t width;
    uint32_t height;
    Q_camera camera;
    uint32_t camera_dirty; /* non-zero: restart accumulation */
} Q_render_frame;

/* ---- plugin types: /root/reference/src/quasi/plugin/plugin_interface.hpp -- */

/* plugin_interface.hpp:27 */
typedef struct Q_plugin_handle Q_plugin_handle;

/* plugin_interface.hpp:30-34 */
typedef struct Q_plugin_version {
    uint32_t major;
    uint32_t minor;
    uint32_t patch;
} Q_plugin_version;

/* plugin_interface.hpp:37-42. Strings have static storage in the plugin. */
typedef struct Q_plugin_info {
    const char* name;
    Q_plugin_version version;
    const char* description;
    const char* author;
} Q_plugin_info;

/* plugin_interface.hpp:48-59. The plugin keeps this pointer; the host must
 * keep the struct alive until Q_plugin_destroy returns. */
typedef struct Q_plugin_context {
    uint32_t viewport_width;
    uint32_t viewport_height;
    void* host_data;
    Q_gpu_context* gpu;
    void (*log)(void* host_data, const char* message);
    void (*request_shutdown)(void* host_data);
} Q_plugin_context;

/* plugin_interface.hpp:65-70. RGBA float, row-major, row 0 = top. */
typedef struct Q_readback_result {
    float* data;
    uint32_t width;
    uint32_t height;
    uint32_t channels; /* always 4 */
} Q_readback_result;

/* plugin_interface.hpp:73-79 */
typedef uint32_t Q_aov_type;
enum {
    Q_AOV_BEAUTY = 0,
    Q_AOV_ALBEDO = 1,
    Q_AOV_NORMAL = 2,
    Q_AOV_DEPTH = 3,
    Q_AOV_COUNT = 4
};

/* plugin_interface.hpp:82-87 */
typedef struct Q_aov_buffer {
    float* data; /* NULL if unavailable */
    uint32_t width;
    uint32_t height;
    uint32_t channels;
} Q_aov_buffer;

/* plugin_interface.hpp:90-92 */
typedef struct Q_readback_aov_result {
    Q_aov_buffer buffers[Q_AOV_COUNT];
} Q_readback_aov_result;

/* plugin_interface.hpp:183. Hosts accept 1..3 (loader.hpp:109-112). */
#define Q_PLUGIN_ABI_VERSION 3u

/* ---- entry points: plugin_interface.hpp:108-148 --------------------------
 * The first six are required by the loader (loader.hpp:72-106), the last four
 * are optional (loader.hpp:115-128). libquasi_cuda.so exports all ten. */

#if defined(__GNUC__)
#define Q_PLUGIN_EXPORT __attribute__((visibility("default")))
#else
#define Q_PLUGIN_EXPORT
#endif

/* plugin_interface.hpp:109 */
Q_PLUGIN_EXPORT uint32_t Q_plugin_abi_version(void);

/* plugin_interface.hpp:112 */
Q_PLUGIN_EXPORT Q_plugin_info Q_plugin_get_info(void);

/* plugin_interface.hpp:117. NULL on: NULL ctx, NULL ctx->gpu, a backend other
 * than Q_GPU_BACKEND_CUDA, or any CUDA failure (reported through ctx->log). */
Q_PLUGIN_EXPORT Q_plugin_handle* Q_plugin_create(Q_plugin_context* ctx);

/* plugin_interface.hpp:121. Waits for enqueued work, frees every allocation. */
Q_PLUGIN_EXPORT void Q_plugin_destroy(Q_plugin_handle* handle);

/* plugin_interface.hpp:126. No-op in both backends. */
Q_PLUGIN_EXPORT void Q_plugin_update(Q_plugin_handle* handle, float delta_time);

/* plugin_interface.hpp:131. Enqueues ONE new sample per pixel on
 * frame->command_buffer and returns without waiting for the GPU. A size change
 * reallocates and restarts accumulation; camera_dirty restarts it. */
Q_PLUGIN_EXPORT void Q_plugin_render(Q_plugin_handle* handle, Q_render_frame* frame);

/* plugin_interface.hpp:136. Blocking. Mean of the samples since the last
 * restart; data == NULL on failure or before the first render. */
Q_PLUGIN_EXPORT Q_readback_result Q_plugin_readback(Q_plugin_handle* handle);

/* plugin_interface.hpp:140. The only way to release a readback buffer. */
Q_PLUGIN_EXPORT void Q_plugin_readback_free(Q_readback_result* result);

/* plugin_interface.hpp:145 */
Q_PLUGIN_EXPORT Q_readback_aov_result Q_plugin_readback_aov(Q_plugin_handle* handle);

/* plugin_interface.hpp:149 */
Q_PLUGIN_EXPORT void Q_plugin_readback_aov_free(Q_readback_aov_result* result);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* QUASI_PLUGIN_ABI_H */
