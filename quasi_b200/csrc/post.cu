/// @file post.cu
/// @brief Bandwidth-bound float4 passes: accumulate, tonemap, shard combine.
///
/// These are the stand-alone forms of the reference's accumulate and tonemap
/// shaders (/root/reference/backends/metal/plugin.mm:313-330, 334-350) and the
/// post-reduce scale of the sample-sharded multi-GPU path. Every
/// thread moves whole float4s (16-byte LDG/STG), consecutive threads touch
/// consecutive addresses, and the grid is a fixed number of waves over the 148
/// SMs with a grid-stride loop. Nothing is reused, so shared memory and TMA
/// staging would only add latency: the roofline for all three is HBM bandwidth.

#include "kernels.hpp"

#include <cuda_runtime.h>

namespace Q::cuda {

namespace {

constexpr int k_threads = 256;
constexpr int k_sm_count = 148;     // B200
constexpr int k_blocks_per_sm = 8;  // 2048 threads/SM = full occupancy at <= 32 regs

inline unsigned grid_for(uint64_t items) {
    uint64_t want = (items + k_threads - 1) / k_threads;
    uint64_t cap = static_cast<uint64_t>(k_sm_count) * k_blocks_per_sm;
    return static_cast<unsigned>(want < cap ? (want ? want : 1) : cap);
}

__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }

/// plugin.mm:313-330: mix(accumulated, new_sample, weight).
__global__ void __launch_bounds__(k_threads)
accumulate_kernel(const float4* __restrict__ current, const float4* __restrict__ history,
                  float4* __restrict__ output, uint64_t n, float weight) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * k_threads;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * k_threads + threadIdx.x; i < n;
         i += stride) {
        float4 c = ld_stream(current + i);
        float4 h = ld_stream(history + i);
        float4 r = make_float4(h.x + (c.x - h.x) * weight, h.y + (c.y - h.y) * weight,
                               h.z + (c.z - h.z) * weight, h.w + (c.w - h.w) * weight);
        __stcs(output + i, r);
    }
}

/// plugin.mm:343-349 for one channel, then Metal's unorm8 conversion.
__device__ __forceinline__ unsigned tonemap_channel(float c) {
    c = c / (c + 1.0f);
    c = powf(c, 1.0f / 2.2f);
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return static_cast<unsigned>(__float2int_rn(c * 255.0f));
}

__device__ __forceinline__ uchar4 tonemap_pixel(float4 c) {
    return make_uchar4(static_cast<unsigned char>(tonemap_channel(c.z)),
                       static_cast<unsigned char>(tonemap_channel(c.y)),
                       static_cast<unsigned char>(tonemap_channel(c.x)), 255);
}

/// plugin.mm:334-350. Output texel (x, y) reads accumulated texel (x, y): the
/// shader's Y flip undoes the rasteriser's.
__global__ void __launch_bounds__(k_threads)
tonemap_kernel(const float4* __restrict__ accumulated, uint32_t width, uint32_t height,
               unsigned char* __restrict__ bgra, uint32_t pitch_bytes) {
    const uint64_t n = static_cast<uint64_t>(width) * height;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * k_threads;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * k_threads + threadIdx.x; i < n;
         i += stride) {
        uint32_t y = static_cast<uint32_t>(i / width);
        uint32_t x = static_cast<uint32_t>(i - static_cast<uint64_t>(y) * width);
        uchar4 o = tonemap_pixel(__ldg(accumulated + i));
        *reinterpret_cast<uchar4*>(bgra + static_cast<size_t>(y) * pitch_bytes + 4u * x) = o;
    }
}

struct shard_list {
    const float4* ptr[k_max_shards];
};

/// Rank-ordered sum of the per-GPU means, scaled by 1/n_shards. With peer-mapped
/// pointers this is the whole reduce: the root GPU pulls each shard over NVLink
/// with 16-byte loads while it writes the result (and optionally the presentable
/// BGRA8 image) locally.
__global__ void __launch_bounds__(k_threads)
combine_shards_kernel(const __grid_constant__ shard_list shards, uint32_t n_shards, uint64_t n,
                      float scale, float4* __restrict__ out, uchar4* __restrict__ bgra) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * k_threads;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * k_threads + threadIdx.x; i < n;
         i += stride) {
        float4 s = ld_stream(shards.ptr[0] + i);
        for (uint32_t g = 1; g < n_shards; g++) {
            float4 v = ld_stream(shards.ptr[g] + i);
            s.x = s.x + v.x;
            s.y = s.y + v.y;
            s.z = s.z + v.z;
            s.w = s.w + v.w;
        }
        s = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
        out[i] = s;
        if (bgra != nullptr) bgra[i] = tonemap_pixel(s);
    }
}

struct sync_list {
    uint32_t* flags[k_max_shards];  ///< rank g's flag block (k_sync_words uint32), peer-mapped
};

constexpr uint32_t k_sync_arrive = 0;               ///< [from rank]: that rank's layers are final for this epoch
constexpr uint32_t k_sync_done = k_max_shards;      ///< [from rank]: that rank has read and written all it will
constexpr uint32_t k_sync_blocks = 2 * k_max_shards;  ///< local: CTAs of this launch that have finished

__device__ __forceinline__ void store_release_system(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t load_acquire_system(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
/// Spins until the epoch stored at p has reached `epoch` (wrap-safe).
__device__ __forceinline__ void wait_epoch(const uint32_t* p, uint32_t epoch) {
    while (static_cast<int32_t>(load_acquire_system(p) - epoch) < 0) __nanosleep(64);
}

/// The whole multi-GPU exchange of one frame as ONE kernel per GPU, with no host
/// synchronisation: every rank runs this on its 1/G slice of the contiguous layer block.
///   arrive   the kernel starts after this rank's render (stream order): it stores
///            `epoch` into every peer's arrive[rank], then waits until every peer has
///            arrived in its own flag block. From here on all layers are final.
///   combine  rank-ordered sum of the G per-rank means over NVLink (16-byte peer
///            loads), 1/G scale, slice stored straight into the root GPU's result
///            buffer (a peer store for every rank but the root): reduce-scatter and
///            gather, no staging.
///   done     the last CTA of this launch stores `epoch` into every peer's done[rank] and
///            waits for every peer's done: when the kernel ends on the root all slices have
///            landed there, and when it ends anywhere nobody still reads this rank's
///            layers, so the next render may overwrite them.
/// Nothing here waits for another CTA of the same launch, and a peer's flags are written by
/// a kernel that itself waits only for `arrive`, so there is no cycle as long as every rank
/// launches the kernel for the same epoch.
__global__ void __launch_bounds__(k_threads)
combine_shards_sync_kernel(const __grid_constant__ shard_list shards, const __grid_constant__ sync_list sync,
                           uint32_t n_shards, uint32_t rank, uint32_t epoch, uint64_t n, float scale,
                           float4* __restrict__ out, uchar4* __restrict__ bgra) {
    uint32_t* mine = sync.flags[rank];
    if (blockIdx.x == 0 && threadIdx.x < n_shards) {
        store_release_system(sync.flags[threadIdx.x] + k_sync_arrive + rank, epoch);
    }
    if (threadIdx.x < n_shards) wait_epoch(mine + k_sync_arrive + threadIdx.x, epoch);
    __syncthreads();

    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * k_threads;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * k_threads + threadIdx.x; i < n; i += stride) {
        float4 s = ld_stream(shards.ptr[0] + i);
        for (uint32_t g = 1; g < n_shards; g++) {
            float4 v = ld_stream(shards.ptr[g] + i);
            s.x = s.x + v.x;
            s.y = s.y + v.y;
            s.z = s.z + v.z;
            s.w = s.w + v.w;
        }
        s = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
        out[i] = s;
        if (bgra != nullptr) bgra[i] = tonemap_pixel(s);
    }

    __threadfence_system();  // this thread's slice stores are visible before the done flag is
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        s_last = atomicAdd(mine + k_sync_blocks, 1u) == gridDim.x - 1u;
        if (s_last) mine[k_sync_blocks] = 0u;  // for the next launch (ordered by the stream)
    }
    __syncthreads();
    if (s_last && threadIdx.x < n_shards) {
        __threadfence_system();
        store_release_system(sync.flags[threadIdx.x] + k_sync_done + rank, epoch);
        wait_epoch(mine + k_sync_done + threadIdx.x, epoch);
    }
}

}  // namespace

cudaError_t launch_accumulate(const float* current, const float* history, float* output,
                              uint64_t n_pixels, uint32_t frame_count, cudaStream_t stream) {
    if (n_pixels == 0) return cudaSuccess;
    float weight = 1.0f / static_cast<float>(frame_count + 1u);
    accumulate_kernel<<<grid_for(n_pixels), k_threads, 0, stream>>>(
        reinterpret_cast<const float4*>(current), reinterpret_cast<const float4*>(history),
        reinterpret_cast<float4*>(output), n_pixels, weight);
    return cudaGetLastError();
}

cudaError_t launch_tonemap_bgra8(const float* accumulated, uint32_t width, uint32_t height,
                                 void* bgra, uint32_t pitch_bytes, cudaStream_t stream) {
    uint64_t n = static_cast<uint64_t>(width) * height;
    if (n == 0) return cudaSuccess;
    tonemap_kernel<<<grid_for(n), k_threads, 0, stream>>>(
        reinterpret_cast<const float4*>(accumulated), width, height,
        static_cast<unsigned char*>(bgra), pitch_bytes);
    return cudaGetLastError();
}

cudaError_t launch_combine_shards(const float* const* shards, uint32_t n_shards,
                                  uint64_t n_pixels, float scale, float* out, void* bgra,
                                  cudaStream_t stream) {
    if (n_shards == 0 || n_shards > k_max_shards) return cudaErrorInvalidValue;
    if (n_pixels == 0) return cudaSuccess;
    shard_list list{};
    for (uint32_t g = 0; g < n_shards; g++) list.ptr[g] = reinterpret_cast<const float4*>(shards[g]);
    if (scale == 0.0f) scale = 1.0f / static_cast<float>(n_shards);
    combine_shards_kernel<<<grid_for(n_pixels), k_threads, 0, stream>>>(
        list, n_shards, n_pixels, scale, reinterpret_cast<float4*>(out),
        static_cast<uchar4*>(bgra));
    return cudaGetLastError();
}

}  // namespace Q::cuda

namespace Q::cuda {

cudaError_t launch_combine_shards_sync(const float* const* shards, uint32_t* const* flags, uint32_t n_shards,
                                       uint32_t rank, uint32_t epoch, uint64_t n_pixels, float scale, float* out,
                                       void* bgra, cudaStream_t stream) {
    if (n_shards == 0 || n_shards > k_max_shards || rank >= n_shards) return cudaErrorInvalidValue;
    shard_list list{};
    sync_list sync{};
    for (uint32_t g = 0; g < n_shards; g++) {
        list.ptr[g] = reinterpret_cast<const float4*>(shards[g]);
        sync.flags[g] = flags[g];
        if (flags[g] == nullptr) return cudaErrorInvalidValue;
    }
    if (scale == 0.0f) scale = 1.0f / static_cast<float>(n_shards);
    combine_shards_sync_kernel<<<grid_for(n_pixels), k_threads, 0, stream>>>(
        list, sync, n_shards, rank, epoch, n_pixels, scale, reinterpret_cast<float4*>(out),
        static_cast<uchar4*>(bgra));
    return cudaGetLastError();
}

}  // namespace Q::cuda
