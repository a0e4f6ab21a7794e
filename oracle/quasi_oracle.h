/*
 * quasi_oracle.h - CPU oracle for the Quasi render hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the shipped
 * product. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * reported CPU baseline. libquasi_cuda.so never links or calls it.
 *
 * What it is: a scalar, plain-C restatement of the Metal shader the reference
 * renders with (/root/reference/backends/metal/plugin.mm:98-350) plus the
 * host-side uniform packing (plugin.mm:648-700) and the demo scene
 * (/root/reference/src/quasi/scene/cornell_box.hpp:32-134). Each function cites
 * the reference lines it follows.
 *
 * Parity pin: the reference has no test, golden image or known-answer vector
 * for this path, and the shader only runs on Metal. The
 * pin is therefore (1) the integer known-answer vectors derived from the
 * shader text (tests/golden/pcg_kat.json) and (2) oracle/_ref: the reference's
 * own MSL source text, extracted from plugin.mm at build time and compiled
 * unmodified with g++ against a Metal-stdlib shim (oracle/msl_shim/), plus the
 * reference's own make_cornell_box() compiled from where it lies. This oracle
 * is checked bit-for-bit against both (tests/test_oracle_vs_ref.py) and the
 * resulting images are committed under tests/golden/.
 *
 * Floating-point definition (the reference compiles with fast-math on Apple
 * GPUs, which is not reproducible):
 *   - every +,-,*,/ and sqrt is one IEEE-754 binary32 operation, round to
 *     nearest even, no FMA contraction, evaluated in source order;
 *   - vector helpers follow /root/reference/src/quasi/math/vec.hpp:88-112:
 *       dot(a,b)   = (a.x*b.x + a.y*b.y) + a.z*b.z
 *       cross(a,b) = {a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x}
 *       normalize(v) = {v.x/l, v.y/l, v.z/l}, l = sqrt(dot(v,v))
 *   - rsqrt(x) = 1.0f / sqrtf(x); mix(a,b,t) = a + (b-a)*t;
 *   - sin, cos, tan, pow are the CORRECTLY ROUNDED binary32 results
 *     (evaluated in binary64 and rounded once).
 */
#ifndef QUASI_ORACLE_H
#define QUASI_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QO_MAX_QUADS 32   /* plugin.mm:21,30 */
#define QO_MAX_BOUNCES 5  /* plugin.mm:22,31 */

/* Same bytes as GpuCamera / GpuQuad / GpuMaterial / GpuSceneUniforms,
 * plugin.mm:354-387 (2624 bytes). */
typedef struct {
    float position[3];
    float fov;
    float direction[3];
    float aspect;
    float up[3];
    float pad_;
} qo_camera;

typedef struct {
    float origin[3];
    float p0_;
    float u[3];
    float p1_;
    float v[3];
    float p2_;
} qo_quad;

typedef struct {
    float albedo[3];
    float roughness;
    float emission[3];
    float metallic;
} qo_material;

typedef struct {
    qo_camera camera;
    uint32_t quad_count;
    uint32_t frame_count;
    uint32_t light_index;
    float pad_;
    qo_quad quads[QO_MAX_QUADS];
    qo_material materials[QO_MAX_QUADS];
} qo_scene_uniforms;

/* One path-traced sample of one pixel: the four MRT outputs of
 * pathtrace_frag (plugin.mm:280-309), each RGBA. */
typedef struct {
    float beauty[4];
    float albedo[4];
    float normal[4];
    float depth[4];
} qo_fragment;

/* plugin.mm:98-102 */
uint32_t qo_pcg_hash(uint32_t input);
/* plugin.mm:104-107 */
float qo_random_float(uint32_t* state);
/* plugin.mm:295 */
uint32_t qo_seed(uint32_t px, uint32_t py, uint32_t frame_count);

/* cornell_box.hpp:95-134. Fills quad_count, light_index, quads, materials;
 * leaves camera and frame_count untouched. */
void qo_make_cornell_box(qo_scene_uniforms* scene);

/* plugin.mm:648-674: host camera (position, target, up, fov) -> camera block. */
void qo_pack_camera(qo_scene_uniforms* scene, const float position[3],
                    const float target[3], const float up[3], float fov,
                    uint32_t width, uint32_t height);

/* plugin.mm:287-309 for one pixel; py = 0 is the top row. Returns the number
 * of trace_scene calls ("rays"). */
uint32_t qo_pathtrace_pixel(const qo_scene_uniforms* scene, uint32_t px,
                            uint32_t py, uint32_t width, uint32_t height,
                            qo_fragment* out);

/* One full 1-spp pass (the PASS 1 draw, plugin.mm:715-738) into four RGBA32F
 * layers of width*height*4 floats, row 0 = top. threads <= 0 uses every
 * online core. Returns the ray count. */
uint64_t qo_render_sample(const qo_scene_uniforms* scene, uint32_t width,
                          uint32_t height, float* beauty, float* albedo,
                          float* normal, float* depth, int threads);

/* plugin.mm:313-330 over n_px pixels (float4 each). out may alias history. */
void qo_accumulate(const float* current, const float* history, float* out,
                   uint64_t n_px, uint32_t frame_count);

/* n_frames calls of Q_plugin_render (plugin.mm:624-792) starting from a reset:
 * call i uses RNG frame index frame_first + i*frame_stride and accumulation
 * weight 1/(i+1). History before call 0 is defined as zero.
 * layers[4] are the accumulated RGBA32F images (beauty, albedo, normal,
 * depth). Returns the ray count. */
uint64_t qo_render_frames(const qo_scene_uniforms* scene, uint32_t width,
                          uint32_t height, uint32_t frame_first,
                          uint32_t frame_stride, uint32_t n_frames,
                          float* const layers[4], int threads);

/* plugin.mm:334-350 -> BGRA8Unorm, row 0 = top (the Y flip in the shader
 * cancels the rasteriser's). */
void qo_tonemap_bgra8(const float* accumulated, uint32_t width, uint32_t height,
                      uint8_t* bgra);

/* Sample-sharded combine: out = (sum_g shard[g]) * (1/G)
 * summed in rank order g = 0..G-1. Not in the reference. */
void qo_combine_shards(const float* const* shards, uint32_t n_shards,
                       uint64_t n_floats, float* out);

/* Correctly rounded binary32 sin/cos, exposed so the GPU implementation can be
 * checked against them directly. */
float qo_sin(float x);
float qo_cos(float x);
float qo_tan(float x);
void qo_sincos_array(const float* phi, uint64_t n, float* sin_out, float* cos_out);

int qo_online_cores(void);

#ifdef __cplusplus
}
#endif
#endif
