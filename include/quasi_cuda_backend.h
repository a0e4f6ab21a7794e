/*
 * quasi_cuda_backend.h - what the opaque handles of quasi_plugin_abi.h mean
 * when Q_gpu_context::backend == Q_GPU_BACKEND_CUDA, plus the optional
 * CUDA-only entry points of libquasi_cuda.so.
 *
 * The ten Q_plugin_* functions are the whole contract with an unmodified
 * Quasi host. Everything here is additive: a host that knows nothing about
 * CUDA fills the three descriptors below in its gpu/cuda context (the
 * counterpart of Q::gpu::metal::context,
 * /root/reference/src/quasi/gpu/metal/context.hpp:54-103) and never calls a
 * Q_cuda_* function. The Q_cuda_* functions exist for what the reference ABI
 * has no slot for (SURVEY.md section 0 fact 5, section 8b "out-of-band"):
 * ray statistics for the Mrays/s metric, and device-side access to the
 * accumulation layers so a multi-GPU host can reduce them over NVLink without
 * a round trip through host memory.
 *
 * No CUDA header is needed to include this file: streams and device pointers
 * are passed as void*.
 */
#ifndef QUASI_CUDA_BACKEND_H
#define QUASI_CUDA_BACKEND_H

/* A host built inside the Quasi tree already has the ABI types from the reference's own
 * <quasi/plugin/plugin_interface.hpp> (layout-identical to quasi_plugin_abi.h, checked in
 * tests/test_abi_cpu.py). It defines QUASI_HOST_HAS_REFERENCE_HEADERS before including this file
 * and gets only the CUDA additions from here (tests/ref_loader_host.cpp does exactly that). */
#if defined(QUASI_HOST_HAS_REFERENCE_HEADERS)
#include <quasi/plugin/plugin_interface.hpp>
#ifndef Q_PLUGIN_EXPORT
#define Q_PLUGIN_EXPORT
#endif
#ifndef Q_GPU_BACKEND_CUDA_DEFINED
#define Q_GPU_BACKEND_CUDA_DEFINED
/* appended to the reference's enum (types.hpp:14-19) until the tree carries it */
static const Q_gpu_backend Q_GPU_BACKEND_CUDA = (Q_gpu_backend)4;
#endif
#else
#include "quasi_plugin_abi.h"
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Q_gpu_context::device points at one of these, owned by the host for the
 * lifetime of the plugin. Replaces id<MTLDevice>
 * (/root/reference/backends/metal/plugin.mm:570). */
typedef struct Q_cuda_device_desc {
    uint32_t struct_size;  /* sizeof(Q_cuda_device_desc), for forward compatibility */
    int32_t device_ordinal; /* cudaSetDevice() ordinal the plugin must use */
    /* Sample sharding. This GPU renders RNG frame
     * indices shard_index + i*shard_count for its i-th render call since the
     * last restart. A single-GPU host passes {0, 1}, which is the reference's
     * behaviour (frame index == call index, plugin.mm:677,791). */
    uint32_t shard_index;
    uint32_t shard_count;
} Q_cuda_device_desc;

/* Q_render_frame::drawable points at one of these, or is NULL when there is
 * nothing to present (headless EXR rendering): the tonemap pass is skipped
 * then. Replaces id<CAMetalDrawable> (plugin.mm:629, 777-788). */
typedef struct Q_cuda_drawable {
    void* pixels;         /* BGRA8 unorm, row 0 = top; device or pinned-host memory */
    uint32_t pitch_bytes; /* bytes between rows, >= 4*width, multiple of 4 */
    uint32_t reserved;
} Q_cuda_drawable;

/* Q_gpu_context::layer (CAMetalLayer* for Metal, plugin.mm:573) points at one of
 * these, or is NULL. It lets the accumulation survive Q_plugin_destroy /
 * Q_plugin_create, i.e. a hot reload (SURVEY.md section 5 "checkpoint / resume",
 * section 8f rank 4): in the reference all state lives in the plugin handle and a
 * reload restarts at sample 0 (plugin.mm:594-617).
 *   destroy  with a state attached the backend does NOT free its layer block: it
 *            records it here (pointer, size, sample count, the scene and shard it
 *            belongs to). The block now belongs to the host, which releases it
 *            with cudaFree (or Q_cuda_free) when it is done.
 *   create   if the recorded block fits (same viewport, scene tag and shard) the
 *            backend adopts it and continues at sample `frame_count` (the RNG frame
 *            index continues, so the image is bit-identical to an uninterrupted
 *            run); otherwise it frees the block and starts from zero.
 * The host still decides whether to continue: camera_dirty on the next frame
 * restarts the accumulation as always. Zero-initialise it; set struct_size. */
typedef struct Q_cuda_persistent_state {
    uint32_t struct_size;  /* sizeof(Q_cuda_persistent_state) */
    uint32_t valid;        /* written by the backend: the fields below describe a block */
    void* layers;          /* device: 4 contiguous RGBA32F layers (beauty, albedo, normal, depth) */
    uint64_t bytes;        /* 64 * width * height */
    uint32_t width;
    uint32_t height;
    uint32_t frame_count;  /* samples accumulated per pixel by this shard */
    uint32_t scene_tag;    /* hash of scene kind and options */
    uint32_t shard_index;
    uint32_t shard_count;
    int32_t device_ordinal;
    uint32_t reserved;
} Q_cuda_persistent_state;

/* Counters since Q_plugin_create (not reset by camera_dirty). */
typedef struct Q_cuda_stats {
    uint64_t rays;            /* closest-hit queries = trace_scene calls (plugin.mm:199) */
    uint64_t paths;           /* camera rays = pixel samples */
    uint64_t kernel_launches; /* kernels this plugin enqueued */
    uint32_t frame_count;     /* render calls since the last restart */
    uint32_t width;
    uint32_t height;
    uint32_t reserved;
    uint64_t nodes_visited;    /* mesh scene: BVH box tests (32 B each); 0 for the quad scene */
    uint64_t triangles_tested; /* mesh scene: triangle tests (48 B each) */
} Q_cuda_stats;

/* Host-side view of the mesh scene the plugin built (QUASI_CUDA_SCENE=mesh), for
 * checking the builder against the oracle's: triangles in original order (48 B),
 * nodes (32 B), leaf_order (slot -> original index). Pointers stay valid until
 * Q_plugin_destroy. Returns -1 for the quad scene. */
typedef struct Q_cuda_mesh_view {
    const void* triangles;
    const void* nodes;
    const uint32_t* leaf_order;
    uint32_t n_triangles;
    uint32_t n_nodes;
    uint32_t depth;
    uint32_t reserved;
} Q_cuda_mesh_view;
Q_PLUGIN_EXPORT int Q_cuda_plugin_mesh_view(Q_plugin_handle* handle, Q_cuda_mesh_view* out);

/* Device-side view of the accumulated RGBA32F layers (index = Q_aov_type),
 * each width*height*4 floats, row 0 = top, holding the mean of this GPU's
 * frame_count samples. The four layers are contiguous in that order, so
 * device_ptr[0] addresses all 16*width*height floats. Valid until the next
 * resize or Q_plugin_destroy. */
typedef struct Q_cuda_layers {
    float* device_ptr[Q_AOV_COUNT];
    uint32_t width;
    uint32_t height;
    uint32_t frame_count;
    uint32_t reserved;
} Q_cuda_layers;

/* Where batched launches (QUASI_CUDA_BATCH, the default) send every 8x4 tile of the four layers as soon
 * as it has received the last sample of the launch, rows of 8 pixels (128 bytes) at a time, while the
 * rest of the image is still being rendered. The contiguous layer block (4 * width * height float4:
 * beauty, albedo, normal, depth) is cut into `slices` pieces of `float4_per_slice` float4 (the last
 * one may be shorter); float4 i of the block is stored at base[i / float4_per_slice] +
 * (i % float4_per_slice). A multi-GPU host points base[g] into GPU g's staging memory (peer-mapped,
 * Q_cuda_ipc_open): every rank pushes each slice of its layers to the slice's owner over NVLink as
 * part of the render, and the exchange that follows the frame is a local sum (Q_cuda_reduce_slices).
 * Pinned host memory works as well (that is what the plugin's own readback mirror is). */
#define Q_CUDA_MAX_EXPORT_SLICES 16u
typedef struct Q_cuda_export_table {
    uint32_t struct_size;       /* sizeof(Q_cuda_export_table) */
    uint32_t slices;            /* 1 .. Q_CUDA_MAX_EXPORT_SLICES; 0: export nothing (and keep the readback mirror off) */
    uint32_t float4_per_slice;  /* > 0 */
    uint32_t reserved;
    void* base[Q_CUDA_MAX_EXPORT_SLICES];
} Q_cuda_export_table;

/* Sets (table != NULL) or clears the export target of the launches that follow; render calls already
 * collected are launched first, with the previous target. While a table is set the plugin's own
 * readback mirror is off (Q_plugin_readback* copy from the device as usual). Returns 0, or -1 for a
 * malformed table. The memory behind base[] must stay valid until the launches have completed. */
Q_PLUGIN_EXPORT int Q_cuda_plugin_set_export(Q_plugin_handle* handle, const Q_cuda_export_table* table);

/* The export with the reduce fused in (multi-GPU, G ranks). The layer block is cut into G slices along rows of
 * 8x4 tiles: slice s is the float4 [slice_lo[s], slice_lo[s + 1]) with every slice_lo[s] = (layer * height +
 * 4 * tile_row) * width, slice_lo[G] = 4 * width * height, so that the 32 pixels of a tile in one layer (a unit)
 * have one owner. Rank s owns slice s: a staging area of G slots of `slot_stride` float4 (slot g receives rank
 * g's copy of the slice) and 4 * tiles arrival counters (uint32, [layer * tiles + tile], zeroed before the first frame: the launch that completes a unit zeroes its counter again; tiles =
 * ceil(width / 8) * ceil(height / 4)), both mapped by every rank (Q_cuda_ipc_open). A batched launch stores every
 * finished tile into its slot at the owners, then adds one to the units' counters; the launch whose arrival is the
 * G-th for a unit, on whichever GPU, loads the G copies, sums them in rank order, scales by 1/G and stores
 * float4 i of the unit to out[s] + (i - slice_lo[s]), s being the unit's slice (out[s]: the place of slice s in
 * pinned host memory every rank has mapped, or in its owner's device memory): the exchange and the frame's way
 * to the host are part of the render. Every rank must make the same render calls per frame, and
 * a buffer (staging + counters) may be used again only when every rank's launches of the frame before have
 * completed (Q_cuda_reduce_slices_sync with n_float4 = 0 is that barrier; two buffers alternate in
 * quasi_b200/multi_gpu.py). A frame whose last launch is a single sample exports nothing (Q_cuda_plugin_flush
 * returns 0): such a frame goes through Q_cuda_push_slices + Q_cuda_reduce_slices_sync on every rank. */
typedef struct Q_cuda_reduce_table {
    uint32_t struct_size; /* sizeof(Q_cuda_reduce_table) */
    uint32_t ranks;       /* G, 2 .. Q_CUDA_MAX_EXPORT_SLICES */
    uint32_t rank;        /* this rank: its slot */
    uint32_t slot_stride; /* float4 between two slots of a staging area, >= the longest slice */
    uint32_t slice_lo[Q_CUDA_MAX_EXPORT_SLICES + 1];
    uint32_t render_calls; /* render calls the frame will make after this set: only the launch that holds the last of
                            * them exports and reduces (a frame of several launches: one arrival per rank and unit);
                            * 0: every launch does (one launch per frame) */
    void* staging[Q_CUDA_MAX_EXPORT_SLICES];      /* [s]: slot 0 of rank s's staging area for this frame */
    uint32_t* counters[Q_CUDA_MAX_EXPORT_SLICES]; /* [s]: rank s's counters for this frame's buffer */
    void* out[Q_CUDA_MAX_EXPORT_SLICES];          /* [s]: where float4 slice_lo[s] of the finished frame goes */
} Q_cuda_reduce_table;

/* Sets (table != NULL) or clears the fused export + reduce for the launches that follow; replaces a table set
 * with Q_cuda_plugin_set_export and vice versa. Returns 0, or -1 for a malformed table. A table whose slices do
 * not fit the viewport of a launch (not the four layers, or a slice that starts inside a row of tiles) is not
 * used by that launch, which then exports nothing (Q_cuda_plugin_flush returns 0). */
Q_PLUGIN_EXPORT int Q_cuda_plugin_set_reduce(Q_plugin_handle* handle, const Q_cuda_reduce_table* table);

/* Launches the render calls the plugin has collected (QUASI_CUDA_BATCH) on `stream` (NULL: the
 * plugin's own queue) without waiting for them. A host that keeps the device pointers of
 * Q_cuda_plugin_layers across frames calls this before it reads the layers on the device; every
 * Q_plugin_* and Q_cuda_plugin_* call that looks at the layers does it by itself.
 * Returns 1 if the most recent launch exported the whole frame to the export target (a batched
 * launch with a table set, or the readback mirror), 0 if not (a launch of a single sample exports
 * nothing: a multi-GPU host then sends the layers with Q_cuda_push_slices), -1 on a NULL handle. */
Q_PLUGIN_EXPORT int Q_cuda_plugin_flush(Q_plugin_handle* handle, void* stream);

/* Blocking: waits for enqueued renders so that the counters are final. */
Q_PLUGIN_EXPORT int Q_cuda_plugin_stats(Q_plugin_handle* handle, Q_cuda_stats* out);

/* Does not block; work already enqueued on the frame stream is ordered before
 * anything enqueued later on `stream` (a cudaStream_t; NULL = legacy default). */
Q_PLUGIN_EXPORT int Q_cuda_plugin_layers(Q_plugin_handle* handle, void* stream,
                                         Q_cuda_layers* out);

/* Stand-alone passes on caller-owned DEVICE buffers, enqueued on `stream`.
 * They are the CUDA forms of the reference's accumulate and tonemap shaders
 * (plugin.mm:313-330, 334-350) and of the post-reduce scale. */
Q_PLUGIN_EXPORT int Q_cuda_accumulate(const float* current, const float* history, float* output,
                                      uint64_t n_pixels, uint32_t frame_count, void* stream);
Q_PLUGIN_EXPORT int Q_cuda_tonemap_bgra8(const float* accumulated, uint32_t width,
                                         uint32_t height, void* bgra, uint32_t pitch_bytes,
                                         void* stream);
/* out = (sum over g of shards[g]) * scale, summed in order g = 0..n-1;
 * scale == 0 means 1/n_shards. shards[] are device pointers (peer-mapped
 * pointers are fine; n_shards <= 16); optionally also writes the tonemapped
 * BGRA8 image of `out` (tightly packed) when bgra != NULL. After an NCCL sum
 * the same entry point does the scale (+tonemap) with n_shards == 1. */
Q_PLUGIN_EXPORT int Q_cuda_combine_shards(const float* const* shards, uint32_t n_shards,
                                          uint64_t n_pixels, float scale, float* out, void* bgra,
                                          void* stream);

/* Test hook: the path tracer's sin/cos (correctly rounded binary32, evaluated in
 * binary64) of n device floats phi[i] in [0, 2*pi]. */
Q_PLUGIN_EXPORT int Q_cuda_probe_sincos(const float* phi, float* sin_out, float* cos_out,
                                        uint64_t n, void* stream);

/* ---- multi-process peer access (one process per GPU) --------
 * The accumulation layers of every rank are made addressable from every other
 * rank with CUDA IPC, so the sample-shard reduce can be ONE kernel per GPU that
 * loads the peers' layers over NVLink and stores its slice of the result
 * straight into the root GPU's buffer (Q_cuda_combine_shards on peer-mapped
 * pointers): a reduce-scatter and a gather with no staging copy. */
typedef struct Q_cuda_ipc_handle {
    unsigned char bytes[64]; /* cudaIpcMemHandle_t */
} Q_cuda_ipc_handle;

/* Plain device memory owned by the library (cudaMalloc), exportable below. */
Q_PLUGIN_EXPORT int Q_cuda_alloc(uint64_t bytes, void** out_device_ptr);
Q_PLUGIN_EXPORT int Q_cuda_free(void* device_ptr);
Q_PLUGIN_EXPORT int Q_cuda_memset_async(void* device_ptr, int value, uint64_t bytes, void* stream);
/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream): device, peer-mapped or pinned host memory either side. */
Q_PLUGIN_EXPORT int Q_cuda_memcpy_async(void* dst, const void* src, uint64_t bytes, void* stream);
/* Export a Q_cuda_alloc'd buffer, or the plugin's contiguous layer block
 * (device_ptr[0] of Q_cuda_plugin_layers). */
Q_PLUGIN_EXPORT int Q_cuda_ipc_export(void* device_ptr, Q_cuda_ipc_handle* out);
/* Map another process's buffer into this process; enables peer access lazily. */
Q_PLUGIN_EXPORT int Q_cuda_ipc_open(const Q_cuda_ipc_handle* handle, void** out_device_ptr);
Q_PLUGIN_EXPORT int Q_cuda_ipc_close(void* device_ptr);

/* ---- the multi-GPU exchange, push-based -----------------------------------------
 * The layer block of 4 * width * height float4 is cut into G slices of P float4; rank s owns slice s
 * and holds a staging area of G * P float4 (device memory, exported with Q_cuda_ipc_export) in which
 * slot g receives rank g's copy of slice s. Every rank sends slice s of its own layers to
 * staging_s + g * P (Q_cuda_plugin_set_export: the batched render kernel does it tile by tile while it
 * renders; or Q_cuda_push_slices afterwards), then runs Q_cuda_reduce_slices_sync on its own slice.
 * The summation order is the rank order, so the result is bit-identical to
 * oracle/quasi_oracle.c:qo_combine_shards whatever the timing. */

/* block[i] -> table->base[i / P][i % P] for i < n_float4, enqueued on `stream`. */
Q_PLUGIN_EXPORT int Q_cuda_push_slices(const float* block, uint64_t n_float4, const Q_cuda_export_table* table,
                                       void* stream);

/* The same into this rank's slots of a Q_cuda_reduce_table (the slices cut along rows of tiles). */
Q_PLUGIN_EXPORT int Q_cuda_push_slices_reduce(const float* block, uint64_t n_float4, const Q_cuda_reduce_table* table,
                                              void* stream);

/* One launch, enqueue only. Device-side flags replace host barriers: the kernel stores `epoch` into
 * flags[g][rank] of every rank g (its frame, and everything it pushed, is complete: the launch follows the
 * render in stream order), waits until flags[rank][g] >= epoch for every g, then
 * out[i] = (staging[i] + staging[P + i] + ... + staging[(G-1) P + i]) * scale for i < n_float4
 * (scale 0 = 1/G). `out` may be pinned host memory (Q_cuda_host_register): every GPU then sends its
 * slice of the finished frame to the host over its own PCIe link. `flags[g]`: rank g's flag block
 * (Q_CUDA_SYNC_WORDS zeroed uint32, Q_cuda_alloc + Q_cuda_ipc_export); `epoch` counts frames from 1 and
 * is the same on every rank. The staging area must be double-buffered by frame parity (post.cu). */
#define Q_CUDA_SYNC_WORDS 33u
Q_PLUGIN_EXPORT int Q_cuda_reduce_slices_sync(const float* staging, uint32_t float4_per_slice,
                                              uint32_t* const* flags, uint32_t n_shards, uint32_t rank,
                                              uint32_t epoch, uint64_t n_float4, float scale, float* out,
                                              void* stream);

/* Pins `bytes` of host memory the caller owns (e.g. a shared-memory segment every rank has mapped) and
 * returns the address kernels and copies of this process use for it. */
Q_PLUGIN_EXPORT int Q_cuda_host_register(void* host_ptr, uint64_t bytes, void** out_device_ptr);
Q_PLUGIN_EXPORT int Q_cuda_host_unregister(void* host_ptr);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* QUASI_CUDA_BACKEND_H */
