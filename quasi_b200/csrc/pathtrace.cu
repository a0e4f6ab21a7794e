/// @file pathtrace.cu
/// @brief Fused camera-ray / intersection / shading / accumulate kernel for sm_100a.
///
/// One launch does what Q_plugin_render's PASS 1 (pathtrace_frag into four MRT
/// targets) and PASS 2 (four accumulate dispatches) do in the Metal backend
/// (/root/reference/backends/metal/plugin.mm:715-774): every pixel gets one new
/// sample, folded straight into the four accumulated RGBA32F layers in place.
/// The 64 B/px MRT store, its re-read and the ping-pong copies never exist:
/// 128 B/px of HBM traffic per call instead of 276.
///
/// Execution model (DESIGN.md "Kernels"):
///  - Persistent warps with path regeneration. The grid is resident CTAs only
///    (4 per SM); each warp owns every n_warps-th 8x4 pixel tile. A lane whose
///    path ended writes its pixel and immediately starts the warp's next pixel
///    (ballot + popc compaction of the idle lanes), so every lane of the warp is
///    always inside the same bounce step. In the thread-per-pixel form the warp
///    ran at 18 of 32 lanes (ncu, profiles/r1_v1_*).
///  - The 16-quad scan is fully unrolled; plane, edge and margin constants are
///    immediate constant-bank operands. The scan DECIDES with cheap forms
///    (MUFU.RCP instead of IEEE division, one dot product per barycentric
///    instead of two cross products) plus explicit margins, then computes the
///    accepted hit with the literal IEEE sequence. Anything inside a margin
///    (an edge, a near tie between two quads, t next to t_min) goes to the
///    literal scan, so the result is bit-identical to the shader's loop.
///  - Shading constants of the hit quad are read from shared memory (the index
///    diverges across lanes; constant-bank loads would serialise).
///
/// Arithmetic contract (DESIGN.md "Numerics"): -fmad=false, IEEE division and
/// square root; every value that reaches a pixel is computed with the shader's
/// operation order. sin/cos are evaluated in binary64 and rounded once.

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

/// One MUFU.RCP, relative error <= 2^-23. Only ever feeds decisions with margins.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

struct hit_record {
    float t;
    v3 point;
    int index;   ///< quad index, -1 = miss
    bool front;  ///< denom < 0: the stored normal faces the ray
};

/// The shader's loop, literally: trace_scene + intersect_quad, plugin.mm:157-212.
/// Used for the rare rays the fast scan cannot decide with its margins.
__device__ __noinline__ void trace_scene_exact(const scene_params& scene, v3 ro, v3 rd,
                                               hit_record& closest) {
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
}

constexpr int k_shade_stride = 25;  // odd: lanes reading different quads hit different banks

// Offsets inside the shared-memory shading record (quad_params floats 28..51).
constexpr int k_sh_normal = 0, k_sh_is_light = 3, k_sh_frame_u = 4, k_sh_plane_d = 7,
              k_sh_frame_v = 8, k_sh_frame_w = 12, k_sh_albedo = 16, k_sh_emission = 20;

// Margins of the fast scan. t' = num * rcp(denom) is within 2^-21 relative of the
// IEEE quotient t; the bands below are ten times wider than that.
constexpr float k_inf = __builtin_huge_valf();
constexpr float k_t_min_lo = 0.001f * (1.0f - 4e-6f);  // below: surely t < t_min
constexpr float k_t_min_hi = 0.001f * (1.0f + 4e-6f);  // above: surely t >= t_min
constexpr float k_tie = 1.0f + 4e-6f;                  // two hits closer than this: literal scan

/// Closest hit of one ray against the NQ first quads. Same result, bit for bit,
/// as trace_scene_exact. `shade` is the shared-memory copy of the quads' shading
/// records; the winner's plane is re-read from it for the literal finalisation.
template <int NQ>
__device__ __forceinline__ void trace_scene(const scene_params& scene, const float* shade, v3 ro,
                                            v3 rd, hit_record& hit) {
    float best_t = k_inf;        // smallest t' among quads that are surely hit
    float challenger_t = k_inf;  // smallest t' among every other quad that may be hit
    int best_i = -1;

#pragma unroll
    for (int i = 0; i < NQ; i++) {
        const quad_params& q = scene.quads[i];
        if (q.valid == 0.0f) continue;  // launch-uniform
        const v3 normal = load_v3(q.normal);
        const float denom = dot(normal, rd);            // the shader's value
        const float num = q.plane_d - dot(normal, ro);  // the shader's value
        const float t_fast = num * rcp_approx(denom);
        // plugin.mm:174 exactly; plugin.mm:177 (t < t_min) with a margin
        if (fabsf(denom) >= 1e-8f && t_fast > k_t_min_lo) {
            const v3 planar = (ro + rd * t_fast) - load_v3(q.origin);
            const float alpha = dot(planar, load_v3(q.fast_a));
            const float beta = dot(planar, load_v3(q.fast_b));
            const float lo = fminf(alpha, beta);
            const float hi = fmaxf(alpha, beta);
            const bool maybe = lo >= -q.delta && hi <= q.one_plus_delta;  // else surely outside
            const bool sure = lo > q.delta && hi < q.one_minus_delta && t_fast > k_t_min_hi;
            const bool better = maybe && sure && t_fast < best_t;
            // Whoever does not become the winner challenges it.
            const float loser = better ? best_t : (maybe ? t_fast : k_inf);
            challenger_t = fminf(challenger_t, loser);
            best_t = better ? t_fast : best_t;
            best_i = better ? i : best_i;
        }
    }

    // A challenger matters only if it could be in front of, or tie with, the sure
    // winner. With no sure winner (best_t == inf) any challenger matters.
    if (challenger_t < k_inf && challenger_t <= best_t * k_tie) {
        trace_scene_exact(scene, ro, rd, hit);
        return;
    }
    hit.index = best_i;
    if (best_i < 0) return;
    // The accepted hit, with the shader's operations: plugin.mm:173, 176, 179.
    const float* sh = shade + best_i * k_shade_stride;
    const v3 normal = load_v3(sh + k_sh_normal);
    const float denom = dot(normal, rd);
    hit.front = denom < 0.0f;
    hit.t = (sh[k_sh_plane_d] - dot(normal, ro)) / denom;
    hit.point = ro + rd * hit.t;
}

__device__ __forceinline__ float4 mix4(float4 a, float4 b, float w) {
    return make_float4(a.x + (b.x - a.x) * w, a.y + (b.y - a.y) * w, a.z + (b.z - a.z) * w,
                       a.w + (b.w - a.w) * w);
}

/// accumulate, plugin.mm:313-330, in place on one pixel of one layer. History
/// before the first sample is defined as zero, so sample 0 is stored as is.
__device__ __forceinline__ void accumulate_pixel(float4* __restrict__ layer, size_t idx,
                                                 float4 sample, bool first, float weight) {
    if (!first) sample = mix4(layer[idx], sample, weight);
    layer[idx] = sample;
}

constexpr int k_block_threads = 256;
constexpr int k_blocks_per_sm = 4;
constexpr int k_warps_per_block = k_block_threads / 32;

template <int NQ>
__global__ void __launch_bounds__(k_block_threads, k_blocks_per_sm)
pathtrace_accumulate_kernel(const __grid_constant__ scene_params scene,
                            const __grid_constant__ frame_params frame, float4* __restrict__ beauty,
                            float4* __restrict__ albedo, float4* __restrict__ normal,
                            float4* __restrict__ depth, unsigned long long* __restrict__ ray_counter) {
    __shared__ float s_shade[NQ * k_shade_stride];
    for (int idx = threadIdx.x; idx < NQ * static_cast<int>(k_shade_floats); idx += k_block_threads) {
        const int qi = idx / static_cast<int>(k_shade_floats);
        const int k = idx - qi * static_cast<int>(k_shade_floats);
        s_shade[qi * k_shade_stride + k] =
            reinterpret_cast<const float*>(&scene.quads[qi])[k_shade_offset + k];
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lanes_below = (1u << lane) - 1u;
    const uint32_t warp_global = blockIdx.x * k_warps_per_block + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * k_warps_per_block;
    const uint32_t tiles_x = (frame.width + 7u) >> 3;
    const uint32_t n_tiles = tiles_x * ((frame.height + 3u) >> 2);
    const bool first_sample = frame.sample_index == 0;
    const float weight = frame.weight;
    const uint32_t max_bounces = scene.max_bounces;
    const v3 cam_origin = load_v3(frame.position);

    // Path state of this lane.
    bool active = false;
    v3 ro = cam_origin, rd = make_v3(0.0f, 0.0f, -1.0f);
    v3 throughput = make_v3(1.0f, 1.0f, 1.0f), color = make_v3(0.0f, 0.0f, 0.0f);
    uint32_t rng = 0, bounce = 0;
    size_t pixel = 0;
    uint32_t rays = 0;
    uint32_t next_item = 0;  // warp-uniform cursor into this warp's pixel sequence

    while (true) {
        // Regeneration: idle lanes take the warp's next pixels, compacted.
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u) {
            if (!active) {
                const uint32_t item = next_item + __popc(idle & lanes_below);
                const uint32_t tile = warp_global + (item >> 5) * n_warps;
                if (tile < n_tiles) {
                    const uint32_t tile_y = tile / tiles_x;
                    const uint32_t tile_x = tile - tile_y * tiles_x;
                    const uint32_t px = tile_x * 8u + (item & 7u);
                    const uint32_t py = tile_y * 4u + ((item >> 3) & 3u);
                    if (px < frame.width && py < frame.height) {
                        // pathtrace_frag, plugin.mm:287-309; uv is the closed form of
                        // the rasteriser's interpolation, py = 0 on top.
                        float uv_x = (static_cast<float>(px) + 0.5f) / static_cast<float>(frame.width);
                        float uv_y = 1.0f - (static_cast<float>(py) + 0.5f) /
                                                static_cast<float>(frame.height);
                        rng = pcg_hash(px + py * 1920u + frame.frame_index * 1920u * 1080u);
                        // get_camera_ray, plugin.mm:129-153: jitter x is drawn before y.
                        const float jx = random_float(rng);
                        const float jy = random_float(rng);
                        uv_x += (jx - 0.5f) * 0.001f;
                        uv_y += (jy - 0.5f) * 0.001f;
                        const v3 target = load_v3(frame.lower_left) +
                                          load_v3(frame.horizontal) * uv_x +
                                          load_v3(frame.vertical) * uv_y;
                        ro = cam_origin;
                        rd = normalize(target - cam_origin);
                        throughput = make_v3(1.0f, 1.0f, 1.0f);
                        color = make_v3(0.0f, 0.0f, 0.0f);
                        bounce = 0;
                        pixel = static_cast<size_t>(py) * frame.width + px;
                        active = true;
                        // The history of this pixel is needed ~1000 cycles from now
                        // (after the first trace); start pulling it out of HBM.
                        if (!first_sample) {
                            prefetch_l2(beauty + pixel);
                            prefetch_l2(albedo + pixel);
                            prefetch_l2(normal + pixel);
                            prefetch_l2(depth + pixel);
                        }
                    }
                }
            }
            next_item += __popc(idle);
        }
        if (!__any_sync(0xffffffffu, active)) {
            if (warp_global + (next_item >> 5) * n_warps >= n_tiles) break;
            continue;  // only off-image pixels of an edge tile were handed out
        }

        if (active) {
            // One bounce of path_trace, plugin.mm:232-273.
            hit_record hit;
            trace_scene<NQ>(scene, s_shade, ro, rd, hit);
            rays++;
            const bool is_hit = hit.index >= 0;
            const float* sh = s_shade + (is_hit ? hit.index : 0) * k_shade_stride;
            const float side = (is_hit && !hit.front) ? -1.0f : 1.0f;
            const v3 n = load_v3(sh + k_sh_normal) * side;  // exact sign flip
            const v3 q_emission = load_v3(sh + k_sh_emission);
            const v3 q_albedo = load_v3(sh + k_sh_albedo);
            const bool is_light = sh[k_sh_is_light] != 0.0f;

            if (bounce == 0) {
                // First-hit AOVs (plugin.mm:241-249, 300-307) go out now, so the
                // lane does not carry them to the end of the path.
                v3 a = make_v3(0.0f, 0.0f, 0.0f), nn = a;
                float first_t = 0.0f;
                if (is_hit) {
                    a = is_light ? q_emission : q_albedo;
                    nn = make_v3(n.x * 0.5f + 0.5f, n.y * 0.5f + 0.5f, n.z * 0.5f + 0.5f);
                    first_t = hit.t;
                }
                const float nd = fminf(fmaxf((first_t - 1.0f) / 4.0f, 0.0f), 1.0f);
                accumulate_pixel(albedo, pixel, make_float4(a.x, a.y, a.z, 1.0f), first_sample, weight);
                accumulate_pixel(normal, pixel, make_float4(nn.x, nn.y, nn.z, 1.0f), first_sample, weight);
                accumulate_pixel(depth, pixel, make_float4(nd, nd, nd, 1.0f), first_sample, weight);
            }

            bool done = !is_hit;
            if (is_hit) {
                color = color + throughput * q_emission;
                done = is_light;
                if (!done) {
                    throughput = throughput * q_albedo;
                    if (bounce > 2) {
                        const float p = fmaxf(0.05f, fmaxf(throughput.x, fmaxf(throughput.y, throughput.z)));
                        if (random_float(rng) > p) {
                            done = true;
                        } else {
                            throughput = throughput / p;
                        }
                    }
                }
            }
            bounce++;
            done = done || bounce >= max_bounces;  // the shader's last scatter is never observed

            if (!done) {
                ro = hit.point + n * 0.001f;
                // cosine_sample_hemisphere (plugin.mm:111-125) with the per-quad frame.
                const float r1 = random_float(rng);
                const float r2 = random_float(rng);
                const float phi = 6.28318548202514648f * r1;  // 2.0f * M_PI_F, exact product
                const float cos_theta = sqrtf(1.0f - r2);
                const float sin_theta = sqrtf(r2);
                double sd, cd;
                sincos(static_cast<double>(phi), &sd, &cd);
                const float sin_phi = static_cast<float>(sd);
                const float cos_phi = static_cast<float>(cd);
                const v3 fu = load_v3(sh + k_sh_frame_u);
                const v3 fv = load_v3(sh + k_sh_frame_v) * side;
                const v3 fw = load_v3(sh + k_sh_frame_w) * side;
                const v3 d = (fu * cos_phi) * sin_theta + (fv * sin_phi) * sin_theta + fw * cos_theta;
                rd = normalize(d);
            } else {
                accumulate_pixel(beauty, pixel, make_float4(color.x, color.y, color.z, 1.0f),
                                 first_sample, weight);
                active = false;
            }
        }
    }

    // One atomic per warp for the ray statistic (SURVEY.md 8d "Ray counting").
    rays = __reduce_add_sync(0xffffffffu, rays);
    if (lane == 0 && rays != 0) atomicAdd(ray_counter, static_cast<unsigned long long>(rays));
}

}  // namespace

cudaError_t launch_pathtrace_accumulate(const scene_params& scene, const frame_params& frame,
                                        float* const layers[4], unsigned long long* ray_counter,
                                        int sm_count, cudaStream_t stream) {
    const uint32_t n_tiles = ((frame.width + 7u) / 8u) * ((frame.height + 3u) / 4u);
    uint32_t blocks = static_cast<uint32_t>(sm_count) * k_blocks_per_sm;
    const uint32_t needed = (n_tiles + k_warps_per_block - 1u) / k_warps_per_block;
    if (needed < blocks) blocks = needed;
    if (blocks == 0) return cudaSuccess;
    auto* b = reinterpret_cast<float4*>(layers[0]);
    auto* a = reinterpret_cast<float4*>(layers[1]);
    auto* n = reinterpret_cast<float4*>(layers[2]);
    auto* d = reinterpret_cast<float4*>(layers[3]);
    if (scene.quad_count <= 16) {
        pathtrace_accumulate_kernel<16><<<blocks, k_block_threads, 0, stream>>>(scene, frame, b, a, n,
                                                                              d, ray_counter);
    } else {
        pathtrace_accumulate_kernel<32><<<blocks, k_block_threads, 0, stream>>>(scene, frame, b, a, n,
                                                                              d, ray_counter);
    }
    return cudaGetLastError();
}

}  // namespace Q::cuda
