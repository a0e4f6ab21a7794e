// A warp/CTA emulator, just large enough to run quasi_b200/csrc/pathtrace.cu's kernel SOURCE on
// the host: every CUDA thread is a std::thread, __syncthreads and the warp collectives are
// barriers over the 256 threads of a CTA / the 32 of a warp, atomics are std::atomic_ref.
// Test infrastructure only (tests/test_kernel_source_on_host_cpu.py): it lets the control flow of
// the persistent kernel (tile queue, path regeneration, the batched form) and its arithmetic be
// checked bit for bit against the oracle without a GPU. It says nothing about what nvcc and
// ptxas make of the same source; the GPU parity tests do.
#pragma once

#include <cuda_runtime.h>  // vector types and make_float4 for the host compiler; __device__ etc. become no-ops

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace emu {

inline constexpr unsigned k_warp = 32;

struct warp_shared {
    std::barrier<> gate{k_warp};
    uint32_t slot[k_warp];
};

struct block_shared {
    explicit block_shared(unsigned threads, size_t dynamic_bytes = 0)
        : gate(threads), warps((threads + k_warp - 1) / k_warp), dynamic((dynamic_bytes + 15) / 16 * 2, 0) {}
    std::barrier<> gate;
    std::vector<warp_shared> warps;
    std::vector<uint64_t> dynamic;  ///< the CTA's dynamic shared memory
};

struct thread_context {
    uint3 thread_idx{0, 0, 0}, block_idx{0, 0, 0};
    dim3 block_dim{1, 1, 1}, grid_dim{1, 1, 1};
    block_shared* block = nullptr;
};
inline thread_local thread_context tls;

inline void* dynamic_shared() { return tls.block->dynamic.data(); }
inline warp_shared& my_warp() { return tls.block->warps[tls.thread_idx.x / k_warp]; }
inline unsigned my_lane() { return tls.thread_idx.x % k_warp; }

/// Every lane deposits a word, then reads all 32: the base of every collective.
template <class F>
inline auto exchange(uint32_t mine, F&& reduce) {
    warp_shared& w = my_warp();
    w.slot[my_lane()] = mine;
    w.gate.arrive_and_wait();
    auto result = reduce(w.slot);
    w.gate.arrive_and_wait();
    return result;
}

/// Runs `kernel(args...)` as grid x block threads, `concurrent_blocks` CTAs at a time.
template <class Kernel, class... Args>
void launch_with_shared(Kernel kernel, unsigned grid, unsigned block, unsigned concurrent_blocks, size_t dynamic_bytes,
                        Args... args) {
    for (unsigned first = 0; first < grid; first += concurrent_blocks) {
        const unsigned n_blocks = std::min(concurrent_blocks, grid - first);
        std::vector<std::unique_ptr<block_shared>> shared;
        std::vector<std::thread> threads;
        for (unsigned b = 0; b < n_blocks; b++) shared.push_back(std::make_unique<block_shared>(block, dynamic_bytes));
        for (unsigned b = 0; b < n_blocks; b++) {
            for (unsigned t = 0; t < block; t++) {
                threads.emplace_back([=, &shared] {
                    tls.thread_idx = {t, 0, 0};
                    tls.block_idx = {first + b, 0, 0};
                    tls.block_dim = dim3(block);
                    tls.grid_dim = dim3(grid);
                    tls.block = shared[b].get();
                    kernel(args...);
                });
            }
        }
        for (auto& th : threads) th.join();
    }
}

template <class Kernel, class... Args>
void launch(Kernel kernel, unsigned grid, unsigned block, unsigned concurrent_blocks, Args... args) {
    launch_with_shared(kernel, grid, block, concurrent_blocks, size_t{0}, args...);
}

/// Phase statistics (QUASI_EMU_STATS): warp-level passes per phase and the lanes working in them.
inline constexpr int k_phases = 8;
inline std::atomic<uint64_t> g_phase_passes[k_phases], g_phase_lanes[k_phases];
inline void phase_pass(int phase, unsigned mask) {
    if (my_lane() == 0 && mask != 0u) {
        g_phase_passes[phase] += 1;
        g_phase_lanes[phase] += static_cast<uint64_t>(__builtin_popcount(mask));
    }
}

}  // namespace emu

// ---- what the kernel source sees -------------------------------------------------------------
#define threadIdx (emu::tls.thread_idx)
#define blockIdx (emu::tls.block_idx)
#define blockDim (emu::tls.block_dim)
#define gridDim (emu::tls.grid_dim)
#undef __shared__
#define __shared__ static  // one CTA's worth: the harness runs CTAs that use it one at a time
#undef __grid_constant__
#define __grid_constant__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __noinline__
#define __noinline__ __attribute__((noinline))

inline void __syncthreads() { emu::tls.block->gate.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::my_warp().gate.arrive_and_wait(); }

inline unsigned __ballot_sync(unsigned, bool predicate) {
    return emu::exchange(predicate ? 1u : 0u, [](const uint32_t* s) {
        unsigned m = 0;
        for (unsigned i = 0; i < emu::k_warp; i++) m |= (s[i] & 1u) << i;
        return m;
    });
}
inline bool __any_sync(unsigned mask, bool predicate) { return __ballot_sync(mask, predicate) != 0u; }
inline uint32_t __shfl_sync(unsigned, uint32_t value, int source_lane) {
    return emu::exchange(value, [source_lane](const uint32_t* s) { return s[source_lane & 31]; });
}
inline uint32_t __reduce_add_sync(unsigned, uint32_t value) {
    return emu::exchange(value, [](const uint32_t* s) {
        uint32_t sum = 0;
        for (unsigned i = 0; i < emu::k_warp; i++) sum += s[i];
        return sum;
    });
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline unsigned __umulhi(unsigned a, unsigned b) { return static_cast<unsigned>((static_cast<uint64_t>(a) * b) >> 32); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __uint2float_rn(unsigned x) { return static_cast<float>(x); }
inline unsigned __float_as_uint(float x) {
    unsigned u;
    std::memcpy(&u, &x, sizeof u);
    return u;
}
inline float __uint_as_float(unsigned u) {
    float x;
    std::memcpy(&x, &u, sizeof x);
    return x;
}
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T __ldcg(const T* p) { return *p; }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
inline unsigned atomicSub(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_sub(v); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return std::atomic_ref<unsigned long long>(*p).fetch_add(v);
}
