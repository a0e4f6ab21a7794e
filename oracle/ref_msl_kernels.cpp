// ref_msl_kernels.cpp - compiles the reference's own MSL shader text on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see oracle/quasi_oracle.h).
//
// quasi_msl_kernels.inc is NOT in this repository: oracle/Makefile extracts it
// verbatim from the string literal at
// /root/reference/backends/metal/plugin.mm:26-351 into oracle/_ref/ (which is
// git-ignored) and this file includes it unmodified inside a namespace, against
// the g++ Metal-stdlib shim in oracle/msl_shim/. No standard header may be
// included here: the shim's address-space macros (thread, constant, device...)
// and its unqualified sqrt/sin/abs would collide with them. ref_driver.cpp
// holds everything that needs the standard library.
#include <metal_stdlib>

namespace ref_msl {
#include "quasi_msl_kernels.inc"
}  // namespace ref_msl

using metal::access;
using metal::texture2d;

extern "C" {

__attribute__((visibility("default"))) unsigned ref_msl_sizeof_uniforms(void) {
    return sizeof(ref_msl::SceneUniforms);
}

__attribute__((visibility("default"))) unsigned ref_msl_pcg_hash(unsigned x) {
    return ref_msl::pcg_hash(x);
}

__attribute__((visibility("default"))) float ref_msl_random_float(unsigned* state) {
    return ref_msl::random_float(*state);
}

// One invocation of pathtrace_frag for the fragment at pixel (px, py), py = 0
// at the top. in.position is the pixel centre; in.uv is the closed form of the
// rasteriser's interpolation of fullscreen_vert.
__attribute__((visibility("default"))) void ref_msl_pathtrace_frag(
    const void* uniforms, unsigned px, unsigned py, unsigned width, unsigned height,
    float* out16) {
    const ref_msl::SceneUniforms& scene = *static_cast<const ref_msl::SceneUniforms*>(uniforms);
    ref_msl::VertexOut in;
    in.position = metal::float4(float(px) + 0.5f, float(py) + 0.5f, 0.0f, 1.0f);
    in.uv.x = (float(px) + 0.5f) / float(width);
    in.uv.y = 1.0f - (float(py) + 0.5f) / float(height);
    ref_msl::FragmentOutput o = ref_msl::pathtrace_frag(in, scene);
    const metal::float4* l[4] = {&o.beauty, &o.albedo, &o.normal, &o.depth};
    for (int i = 0; i < 4; i++) {
        out16[i * 4 + 0] = l[i]->x;
        out16[i * 4 + 1] = l[i]->y;
        out16[i * 4 + 2] = l[i]->z;
        out16[i * 4 + 3] = l[i]->w;
    }
}

// One row of threads of the accumulate kernel (gid.y = gy, gid.x = 0..width-1).
__attribute__((visibility("default"))) void ref_msl_accumulate_row(
    const float* current, const float* history, float* output, unsigned width,
    unsigned height, unsigned frame_count, unsigned gy) {
    texture2d<float, access::read> cur{const_cast<float*>(current), width, height};
    texture2d<float, access::read> hist{const_cast<float*>(history), width, height};
    texture2d<float, access::write> out{output, width, height};
    for (unsigned gx = 0; gx < width; gx++) {
        ref_msl::accumulate(cur, hist, out, frame_count, metal::uint2(gx, gy));
    }
}

// One invocation of tonemap_frag at pixel (px, py); returns the shader's float4.
__attribute__((visibility("default"))) void ref_msl_tonemap_frag(
    const float* accumulated, unsigned width, unsigned height, unsigned px, unsigned py,
    float* out4) {
    texture2d<float> tex{const_cast<float*>(accumulated), width, height};
    ref_msl::VertexOut in;
    in.position = metal::float4(float(px) + 0.5f, float(py) + 0.5f, 0.0f, 1.0f);
    in.uv.x = (float(px) + 0.5f) / float(width);
    in.uv.y = 1.0f - (float(py) + 0.5f) / float(height);
    metal::float4 c = ref_msl::tonemap_frag(in, tex);
    out4[0] = c.x;
    out4[1] = c.y;
    out4[2] = c.z;
    out4[3] = c.w;
}

}  // extern "C"
