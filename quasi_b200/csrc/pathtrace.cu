/// @file pathtrace.cu
/// @brief Fused camera-ray / intersection / shading / accumulate kernel for sm_100a.
///
/// One launch does what Q_plugin_render's PASS 1 (pathtrace_frag into four MRT
/// targets) and PASS 2 (four accumulate dispatches) do in the Metal backend
/// (/root/reference/backends/metal/plugin.mm:715-774): each thread traces one new
/// sample of one pixel and folds it straight into the four accumulated RGBA32F
/// layers in place. The 64 B/px MRT store, its 64 B/px re-read and the ping-pong
/// copies never exist: 128 B/px of HBM traffic per call instead of 276.
///
/// Arithmetic contract (DESIGN.md "Numerics"): compiled with -fmad=false,
/// IEEE division and square root; every expression below is written in the
/// operation order of the shader so results are bit-identical to oracle/.
/// sin/cos are evaluated in binary64 and rounded once (correctly rounded
/// binary32), which is also how the oracle defines them.

#include "kernels.hpp"

#include <cuda_runtime.h>

namespace Q::cuda {

namespace {

struct v3 {
    float x, y, z;
};

__device__ __forceinline__ v3 make_v3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ v3 load_v3(const float* p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ v3 operator+(v3 a, v3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ v3 operator-(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ v3 operator*(v3 a, v3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ v3 operator*(v3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ v3 operator/(v3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
/// (a.x*b.x + a.y*b.y) + a.z*b.z
__device__ __forceinline__ float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ v3 cross(v3 a, v3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ v3 normalize(v3 a) { return a / sqrtf(dot(a, a)); }

/// plugin.mm:98-102
__device__ __forceinline__ uint32_t pcg_hash(uint32_t input) {
    uint32_t state = input * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}

/// plugin.mm:104-107. float(0xFFFFFFFF) is 2^32; dividing by a power of two is
/// the same exactly rounded result as multiplying by 2^-32.
__device__ __forceinline__ float random_float(uint32_t& state) {
    state = pcg_hash(state);
    return __uint2float_rn(state) * 2.3283064365386963e-10f;
}

struct hit_record {
    float t;
    v3 point;
    int index;   ///< quad index, -1 = miss
    bool front;  ///< denom < 0: the stored normal faces the ray
};

/// trace_scene + intersect_quad, plugin.mm:157-212. Planes, w = n/|n|^2 and the
/// degenerate-quad test come pre-computed in scene_params.
__device__ __forceinline__ hit_record trace_scene(const scene_params& scene, v3 ro, v3 rd) {
    hit_record closest;
    closest.t = 1e30f;
    closest.index = -1;
    closest.front = true;
    closest.point = make_v3(0.0f, 0.0f, 0.0f);

    const uint32_t count = scene.quad_count;
#pragma unroll 1
    for (uint32_t i = 0; i < count; i++) {
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
    return closest;
}

struct path_result {
    v3 color;
    v3 albedo;
    v3 normal;
    float depth;
    uint32_t rays;
};

/// path_trace + cosine_sample_hemisphere, plugin.mm:111-125, 223-276.
__device__ __forceinline__ path_result path_trace(const scene_params& scene, v3 ro, v3 rd,
                                                  uint32_t& rng) {
    path_result r;
    r.color = make_v3(0.0f, 0.0f, 0.0f);
    r.albedo = make_v3(0.0f, 0.0f, 0.0f);
    r.normal = make_v3(0.0f, 0.0f, 0.0f);
    r.depth = 0.0f;
    r.rays = 0;
    v3 throughput = make_v3(1.0f, 1.0f, 1.0f);

    const uint32_t max_bounces = scene.max_bounces;
#pragma unroll 1
    for (uint32_t bounce = 0; bounce < max_bounces; bounce++) {
        hit_record hit = trace_scene(scene, ro, rd);
        r.rays++;
        if (hit.index < 0) break;

        const quad_params& q = scene.quads[hit.index];
        const float side = hit.front ? 1.0f : -1.0f;
        v3 normal = load_v3(q.normal) * side;  // exact sign flip
        v3 emission = load_v3(q.emission);
        v3 albedo = load_v3(q.albedo);
        const bool is_light = q.is_light != 0.0f;

        if (bounce == 0) {
            r.albedo = is_light ? emission : albedo;
            r.normal = make_v3(normal.x * 0.5f + 0.5f, normal.y * 0.5f + 0.5f,
                               normal.z * 0.5f + 0.5f);
            r.depth = hit.t;
        }

        r.color = r.color + throughput * emission;
        if (is_light) break;

        throughput = throughput * albedo;

        if (bounce > 2) {
            float p = fmaxf(0.05f, fmaxf(throughput.x, fmaxf(throughput.y, throughput.z)));
            if (random_float(rng) > p) break;
            throughput = throughput / p;
        }

        ro = hit.point + normal * 0.001f;

        // cosine_sample_hemisphere with the per-quad frame (scene_params.hpp).
        float r1 = random_float(rng);
        float r2 = random_float(rng);
        float phi = 6.28318548202514648f * r1;  // 2.0f * M_PI_F, exact product
        float cos_theta = sqrtf(1.0f - r2);
        float sin_theta = sqrtf(r2);
        double sd, cd;
        sincos(static_cast<double>(phi), &sd, &cd);
        float sin_phi = static_cast<float>(sd);
        float cos_phi = static_cast<float>(cd);

        v3 fu = load_v3(q.frame_u);
        v3 fv = load_v3(q.frame_v) * side;
        v3 fw = load_v3(q.frame_w) * side;
        v3 d = (fu * cos_phi) * sin_theta + (fv * sin_phi) * sin_theta + fw * cos_theta;
        rd = normalize(d);
    }
    return r;
}

__device__ __forceinline__ float4 mix4(float4 a, float4 b, float w) {
    return make_float4(a.x + (b.x - a.x) * w, a.y + (b.y - a.y) * w, a.z + (b.z - a.z) * w,
                       a.w + (b.w - a.w) * w);
}

constexpr int k_block_threads = 256;

/// Block = 16x16 pixels; each warp owns an 8x4 tile so that primary rays stay
/// coherent and every row of a tile is one 128-byte float4 segment.
__global__ void __launch_bounds__(k_block_threads)
pathtrace_accumulate_kernel(const __grid_constant__ scene_params scene,
                            const __grid_constant__ frame_params frame, float4* __restrict__ beauty,
                            float4* __restrict__ albedo, float4* __restrict__ normal,
                            float4* __restrict__ depth, unsigned long long* __restrict__ ray_counter) {
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t px = blockIdx.x * 16u + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t py = blockIdx.y * 16u + (warp >> 1) * 4u + (lane >> 3);
    const bool inside = px < frame.width && py < frame.height;

    uint32_t rays = 0;
    if (inside) {
        // pathtrace_frag, plugin.mm:287-309; uv is the closed form of the
        // rasteriser's interpolation, py = 0 at the top.
        float uv_x = (static_cast<float>(px) + 0.5f) / static_cast<float>(frame.width);
        float uv_y = 1.0f - (static_cast<float>(py) + 0.5f) / static_cast<float>(frame.height);

        uint32_t rng = pcg_hash(px + py * 1920u + frame.frame_index * 1920u * 1080u);

        // get_camera_ray, plugin.mm:129-153: jitter x is drawn before jitter y.
        float jx = random_float(rng);
        float jy = random_float(rng);
        uv_x += (jx - 0.5f) * 0.001f;
        uv_y += (jy - 0.5f) * 0.001f;
        v3 origin = load_v3(frame.position);
        v3 target = load_v3(frame.lower_left) + load_v3(frame.horizontal) * uv_x +
                    load_v3(frame.vertical) * uv_y;
        v3 direction = normalize(target - origin);

        path_result pt = path_trace(scene, origin, direction, rng);
        rays = pt.rays;

        float norm_depth = fminf(fmaxf((pt.depth - 1.0f) / 4.0f, 0.0f), 1.0f);
        float4 s_beauty = make_float4(pt.color.x, pt.color.y, pt.color.z, 1.0f);
        float4 s_albedo = make_float4(pt.albedo.x, pt.albedo.y, pt.albedo.z, 1.0f);
        float4 s_normal = make_float4(pt.normal.x, pt.normal.y, pt.normal.z, 1.0f);
        float4 s_depth = make_float4(norm_depth, norm_depth, norm_depth, 1.0f);

        // accumulate, plugin.mm:313-330, in place. History before the first
        // sample is defined as zero, so sample 0 is stored as is.
        const size_t idx = static_cast<size_t>(py) * frame.width + px;
        if (frame.sample_index != 0) {
            const float w = frame.weight;
            s_beauty = mix4(beauty[idx], s_beauty, w);
            s_albedo = mix4(albedo[idx], s_albedo, w);
            s_normal = mix4(normal[idx], s_normal, w);
            s_depth = mix4(depth[idx], s_depth, w);
        }
        beauty[idx] = s_beauty;
        albedo[idx] = s_albedo;
        normal[idx] = s_normal;
        depth[idx] = s_depth;
    }

    // One atomic per warp for the ray statistic (SURVEY.md 8d "Ray counting").
    rays = __reduce_add_sync(0xffffffffu, rays);
    if (lane == 0 && rays != 0) atomicAdd(ray_counter, static_cast<unsigned long long>(rays));
}

}  // namespace

cudaError_t launch_pathtrace_accumulate(const scene_params& scene, const frame_params& frame,
                                        float* const layers[4], unsigned long long* ray_counter,
                                        cudaStream_t stream) {
    dim3 grid((frame.width + 15u) / 16u, (frame.height + 15u) / 16u);
    pathtrace_accumulate_kernel<<<grid, k_block_threads, 0, stream>>>(
        scene, frame, reinterpret_cast<float4*>(layers[0]), reinterpret_cast<float4*>(layers[1]),
        reinterpret_cast<float4*>(layers[2]), reinterpret_cast<float4*>(layers[3]), ray_counter);
    return cudaGetLastError();
}

}  // namespace Q::cuda
