/// @file context.hpp
/// @brief Host-side CUDA context: device, streams, frames in flight, presentable buffers.
///
/// The CUDA counterpart of Q::gpu::metal::context
/// (/root/reference/src/quasi/gpu/metal/context.hpp:54-103, context.mm:76-188):
/// same shape (create / gpu() / begin_frame / end_frame / resize) and the same
/// policy of at most three frames in flight (context.mm:16), where a Metal
/// command buffer becomes a CUDA stream and the drawable-available semaphore
/// becomes a per-slot CUDA event. It owns everything the plugin borrows through
/// Q_gpu_context / Q_render_frame (quasi_cuda_backend.h).
///
/// Nothing here needs nvcc: the implementation uses the CUDA runtime C API.

#pragma once

#include <quasi/async/async.hpp>
#include <quasi/gpu/types.hpp>

#include <array>
#include <cstdint>
#include <expected>

namespace Q::gpu::cuda {

inline constexpr int k_max_frames_in_flight = 3;  ///< context.mm:16

enum class context_error {
    device_not_found,   ///< No CUDA device with that ordinal.
    stream_failed,      ///< Could not create streams or events.
    allocation_failed,  ///< Could not allocate the presentable buffers.
    no_drawable,        ///< begin_frame() could not get a frame slot.
};

[[nodiscard]] constexpr const char* to_string(context_error e) noexcept {
    switch (e) {
        case context_error::device_not_found: return "CUDA device not found";
        case context_error::stream_failed: return "CUDA stream setup failed";
        case context_error::allocation_failed: return "CUDA framebuffer allocation failed";
        case context_error::no_drawable: return "Failed to get a frame slot";
    }
    return "Unknown error";
}

struct context_options {
    int device_ordinal = 0;
    uint32_t shard_index = 0;  ///< this GPU's sample shard
    uint32_t shard_count = 1;
    bool present = false;      ///< allocate BGRA8 presentable buffers (pinned host memory)
    /// Offer the backend a Q_cuda_persistent_state through Q_gpu_context::layer, so the
    /// accumulation outlives a plugin instance (a hot reload continues instead of restarting).
    bool keep_accumulation = false;
};

class context {
public:
    template <typename T>
    using result = std::expected<T, context_error>;

    [[nodiscard]] static result<context> create(uint32_t width, uint32_t height,
                                                context_options options = {});

    context() = default;
    ~context();
    context(context&& other) noexcept;
    context& operator=(context&& other) noexcept;
    context(const context&) = delete;
    context& operator=(const context&) = delete;

    [[nodiscard]] gpu_context* gpu() noexcept { return &gpu_ctx_; }
    [[nodiscard]] const gpu_context* gpu() const noexcept { return &gpu_ctx_; }

    /// Hands out the next frame slot; blocks only if that slot's previous frame
    /// (three frames ago) is still running. Coroutines call slot_available() first.
    [[nodiscard]] result<render_frame> begin_frame();

    /// Marks the end of the frame's GPU work (the presentDrawable + commit of Metal).
    void end_frame(render_frame& frame);

    void resize(uint32_t width, uint32_t height);

    [[nodiscard]] uint32_t width() const noexcept { return width_; }
    [[nodiscard]] uint32_t height() const noexcept { return height_; }
    [[nodiscard]] int device_ordinal() const noexcept { return device_desc_.device_ordinal; }
    [[nodiscard]] void* queue() const noexcept { return queue_; }  ///< cudaStream_t

    /// True when begin_frame() would not block.
    [[nodiscard]] bool slot_available() const;
    /// True when no frame is in flight.
    [[nodiscard]] bool idle() const;
    /// Coroutine forms for the scheduler: poll, never block.
    [[nodiscard]] async::task<void> wait_slot();
    [[nodiscard]] async::task<void> wait_idle();
    /// Blocking drain (shutdown paths only).
    void synchronize();

    /// GPU time of the most recently COMPLETED frame in milliseconds (begin_frame
    /// to end_frame on its stream), or a negative value if none has completed.
    [[nodiscard]] float last_frame_gpu_ms();

    /// Presentable BGRA8 pixels of the most recently completed frame (pinned host
    /// memory, row 0 = top), or nullptr when created without `present`.
    [[nodiscard]] const uint8_t* presented_pixels();

private:
    struct slot {
        void* stream = nullptr;       // cudaStream_t
        void* begin_event = nullptr;  // cudaEvent_t
        void* end_event = nullptr;    // cudaEvent_t
        bool in_flight = false;
        Q_cuda_drawable drawable{};
        uint8_t* pixels = nullptr;    // pinned host, 4 * width * height: what the host "presents"
        void* device_pixels = nullptr;  // the drawable the plugin tonemaps into, same size, in HBM
    };

    void release();
    bool allocate_drawables();
    void free_drawables();

    gpu_context gpu_ctx_{};
    Q_cuda_device_desc device_desc_{};
    void* queue_ = nullptr;
    std::array<slot, k_max_frames_in_flight> slots_{};
    int next_slot_ = 0;
    int last_completed_slot_ = -1;
    uint32_t width_ = 0;
    uint32_t height_ = 0;
    bool present_ = false;
    bool keep_accumulation_ = false;
    Q_cuda_persistent_state persistent_{};  ///< owned here; the backend parks its layer block in it
    bool valid_ = false;
};

/// Completes when the CUDA event has (polls cudaEventQuery once per tick).
[[nodiscard]] async::task<void> wait_event(void* cuda_event);

/// Number of CUDA devices visible to the process (0 on error).
[[nodiscard]] int device_count();

}  // namespace Q::gpu::cuda
