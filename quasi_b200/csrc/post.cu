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

/// plugin.mm:313-330: mix(accumulated, new_sample, weight). Two float4 per thread and iteration,
/// all four loads issued before the first use (the pass is pure HBM streaming: what limits it is
/// bytes in flight per SM).
__global__ void __launch_bounds__(k_threads)
accumulate_kernel(const float4* __restrict__ current, const float4* __restrict__ history,
                  float4* __restrict__ output, uint64_t n, float weight) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * k_threads;
    auto mix = [weight](float4 c, float4 h) {
        return make_float4(h.x + (c.x - h.x) * weight, h.y + (c.y - h.y) * weight, h.z + (c.z - h.z) * weight,
                           h.w + (c.w - h.w) * weight);
    };
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * k_threads + threadIdx.x;
    for (; i + stride < n; i += 2 * stride) {
        const float4 c0 = ld_stream(current + i), h0 = ld_stream(history + i);
        const float4 c1 = ld_stream(current + i + stride), h1 = ld_stream(history + i + stride);
        __stcs(output + i, mix(c0, h0));
        __stcs(output + i + stride, mix(c1, h1));
    }
    if (i < n) __stcs(output + i, mix(ld_stream(current + i), ld_stream(history + i)));
}

__device__ __forceinline__ float mufu_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

/// plugin.mm:343-349 for one channel, then Metal's unorm8 conversion: the literal form.
__device__ __noinline__ unsigned tonemap_channel_precise(float c) {
    c = c / (c + 1.0f);
    c = powf(c, 1.0f / 2.2f);
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return static_cast<unsigned>(__float2int_rn(c * 255.0f));
}

/// The same 8-bit value, cheaply: three MUFU operations give 255 * (c / (c + 1))^(1/2.2) to within
/// 6e-4 (rcp 2^-23, lg2 2^-22 absolute on an exponent of at most 22 where the result is not yet
/// zero, ex2 2^-22), which rounds to the same integer as the literal form unless it lies within
/// 7e-4 of a rounding boundary (6e-4, the literal form's own 6e-5 and the product's 1.5e-5); those
/// channels (0.14 %), and anything not finite, take the literal form. The width of that band is what
/// the pass costs: the literal form is ~150 instructions for the whole warp whenever ONE lane needs it,
/// and with a band of 2e-3 (0.4 % of the channels, 12 % of the warp-level calls) the pass ran at 0.65
/// of HBM bandwidth at 4K, with 1e-3 at 0.77, with 7e-4 at 0.79 (profiles/r2b_tonemap_band.log). Rounding and conversion are done by the adder (y + 2^23 is y rounded to nearest even, its
/// low mantissa bits are the integer): F2I and FRND share the MUFU pipe, which is what bounds this
/// pass once powf is gone (ncu, round 2: powf for every channel ran at 0.24 of HBM bandwidth, the
/// MUFU form with floorf and F2I at 0.43 with the XU pipe 71 % busy).
__device__ __forceinline__ unsigned tonemap_channel(float c) {
    const float r = c * mufu_rcp(c + 1.0f);
    const float y = mufu_ex2(mufu_lg2(r) * (1.0f / 2.2f)) * 255.0f;
    const float rounded = y + 8388608.0f;
    const float off = y - (rounded - 8388608.0f);  // in [-1/2, 1/2]
    if (!(fabsf(off) <= 0.5f - 7e-4f)) return tonemap_channel_precise(c);  // near a boundary, or NaN
    return __float_as_uint(rounded) & 0xffu;
}

__device__ __forceinline__ uchar4 tonemap_pixel(float4 c) {
    return make_uchar4(static_cast<unsigned char>(tonemap_channel(c.z)),
                       static_cast<unsigned char>(tonemap_channel(c.y)),
                       static_cast<unsigned char>(tonemap_channel(c.x)), 255);
}

/// plugin.mm:334-350. Output texel (x, y) reads accumulated texel (x, y): the
/// shader's Y flip undoes the rasteriser's. `Packed`: rows follow each other without padding,
/// so the texel index is the pixel index (no division per pixel).
template <bool Packed>
__global__ void __launch_bounds__(k_threads)
tonemap_kernel(const float4* __restrict__ accumulated, uint32_t width, uint32_t height,
               unsigned char* __restrict__ bgra, uint32_t pitch_bytes) {
    const uint64_t n = static_cast<uint64_t>(width) * height;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * k_threads;
    auto store = [&](uint64_t i, uchar4 o) {
        if constexpr (Packed) {
            __stcs(reinterpret_cast<uchar4*>(bgra) + i, o);
        } else {
            const uint32_t y = static_cast<uint32_t>(i / width);
            const uint32_t x = static_cast<uint32_t>(i - static_cast<uint64_t>(y) * width);
            *reinterpret_cast<uchar4*>(bgra + static_cast<size_t>(y) * pitch_bytes + 4u * x) = o;
        }
    };
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * k_threads + threadIdx.x;
    for (; i + stride < n; i += 2 * stride) {
        const float4 a = ld_stream(accumulated + i), b = ld_stream(accumulated + i + stride);
        store(i, tonemap_pixel(a));
        store(i + stride, tonemap_pixel(b));
    }
    if (i < n) store(i, tonemap_pixel(ld_stream(accumulated + i)));
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

struct export_list {
    float4* base[k_max_export_slices];
    uint32_t slice_lo[k_max_export_slices + 1];
    uint32_t slices;
};

/// The layer block of this GPU, slice by slice, to where batch_params' export would have sent it
/// (scene_params.hpp): for frames whose last launch was not a batched one (a launch of one sample
/// exports nothing). Reads HBM once, writes over NVLink (or PCIe) with 16-byte stores.
__global__ void __launch_bounds__(k_threads)
push_slices_kernel(const float4* __restrict__ block, const __grid_constant__ export_list to, uint64_t n) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * k_threads;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * k_threads + threadIdx.x; i < n; i += stride) {
        const uint32_t index = static_cast<uint32_t>(i);
        uint32_t slice = 0;
        while (slice + 1u < to.slices && index >= to.slice_lo[slice + 1u]) slice++;
        to.base[slice][index - to.slice_lo[slice]] = ld_stream(block + i);
    }
}

/// The multi-GPU exchange of one frame, on the GPU that owns a slice: every rank has PUSHED its part of this
/// slice into this GPU's staging memory while it rendered (the export of the batched launches, or
/// push_slices_kernel), so what is left is a local pass: wait until every rank has said that its frame is
/// complete, sum the G staged copies in rank order, scale by 1/G, store the slice where the frame is wanted
/// (pinned host memory: each GPU sends its own slice over its own PCIe link; or device memory).
///   arrive   this kernel follows this rank's render in stream order: block 0 stores `epoch` into every
///            rank's arrive[rank] (release, system scope), then every block waits until all G ranks have
///            arrived in this GPU's own flag block. From here on all G staged copies are complete.
/// Nothing waits for another block of this launch or for a later kernel of a peer, so there is no cycle as
/// long as every rank launches the kernel for the same epoch. The staging memory is double-buffered by the
/// caller (epoch parity): a rank can push frame e + 2 only after its reduce of frame e + 1 has seen every
/// rank arrive, i.e. after every rank's reduce of frame e (which precedes its render of e + 1 in stream
/// order) has finished reading the buffer of frame e.
/// Pull (the first form: every rank loads its slice of all G layer blocks over NVLink, round trips) took
/// 0.75 ms for a 1080p frame on 8 GPUs against 0.65 ms for ncclReduce + scale.
__global__ void __launch_bounds__(k_threads)
reduce_slices_sync_kernel(const float4* __restrict__ staging, uint32_t per, const __grid_constant__ sync_list sync,
                          uint32_t n_shards, uint32_t rank, uint32_t epoch, uint64_t n, float scale,
                          float4* __restrict__ out) {
    uint32_t* mine = sync.flags[rank];
    if (blockIdx.x == 0 && threadIdx.x < n_shards) {
        __threadfence_system();
        store_release_system(sync.flags[threadIdx.x] + rank, epoch);
    }
    if (threadIdx.x < n_shards) wait_epoch(mine + threadIdx.x, epoch);
    __syncthreads();

    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * k_threads;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * k_threads + threadIdx.x; i < n; i += stride) {
        float4 s = ld_stream(staging + i);
        for (uint32_t g = 1; g < n_shards; g++) {
            const float4 v = ld_stream(staging + static_cast<uint64_t>(g) * per + i);
            s.x = s.x + v.x;
            s.y = s.y + v.y;
            s.z = s.z + v.z;
            s.w = s.w + v.w;
        }
        __stcs(out + i, make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale));
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
    if (pitch_bytes == 4u * width) {
        tonemap_kernel<true><<<grid_for(n), k_threads, 0, stream>>>(
            reinterpret_cast<const float4*>(accumulated), width, height, static_cast<unsigned char*>(bgra), pitch_bytes);
    } else {
        tonemap_kernel<false><<<grid_for(n), k_threads, 0, stream>>>(
            reinterpret_cast<const float4*>(accumulated), width, height, static_cast<unsigned char*>(bgra), pitch_bytes);
    }
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

cudaError_t launch_push_slices(const float* block, uint64_t n_float4, void* const* base, const uint32_t* slice_lo,
                               uint32_t slices, cudaStream_t stream) {
    if (slices == 0 || slices > k_max_export_slices || n_float4 >= (uint64_t{1} << 32) || slice_lo[0] != 0 ||
        slice_lo[slices] < n_float4) {
        return cudaErrorInvalidValue;
    }
    if (n_float4 == 0) return cudaSuccess;
    export_list to{};
    to.slices = slices;
    for (uint32_t g = 0; g < slices; g++) {
        if (base[g] == nullptr || slice_lo[g + 1] < slice_lo[g]) return cudaErrorInvalidValue;
        to.base[g] = static_cast<float4*>(base[g]);
        to.slice_lo[g] = slice_lo[g];
    }
    to.slice_lo[slices] = slice_lo[slices];
    push_slices_kernel<<<grid_for(n_float4), k_threads, 0, stream>>>(reinterpret_cast<const float4*>(block), to, n_float4);
    return cudaGetLastError();
}

cudaError_t launch_reduce_slices_sync(const float* staging, uint32_t float4_per_slice, uint32_t* const* flags,
                                      uint32_t n_shards, uint32_t rank, uint32_t epoch, uint64_t n_float4, float scale,
                                      float* out, cudaStream_t stream) {
    if (n_shards == 0 || n_shards > k_max_shards || rank >= n_shards || n_float4 > float4_per_slice) {
        return cudaErrorInvalidValue;
    }
    sync_list sync{};
    for (uint32_t g = 0; g < n_shards; g++) {
        if (flags[g] == nullptr) return cudaErrorInvalidValue;
        sync.flags[g] = flags[g];
    }
    if (scale == 0.0f) scale = 1.0f / static_cast<float>(n_shards);
    // at least one block even for an empty slice: the arrive flags are part of every rank's protocol
    reduce_slices_sync_kernel<<<grid_for(n_float4 ? n_float4 : 1), k_threads, 0, stream>>>(
        reinterpret_cast<const float4*>(staging), float4_per_slice, sync, n_shards, rank, epoch, n_float4, scale,
        reinterpret_cast<float4*>(out));
    return cudaGetLastError();
}

}  // namespace Q::cuda
