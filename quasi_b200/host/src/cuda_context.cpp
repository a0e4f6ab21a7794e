/// @file cuda_context.cpp
/// @brief Q::gpu::cuda::context implementation (CUDA runtime C API, no nvcc).

#include <quasi/gpu/cuda/context.hpp>

#include <cuda_runtime_api.h>

#include <utility>

namespace Q::gpu::cuda {

namespace {
inline cudaStream_t as_stream(void* p) { return static_cast<cudaStream_t>(p); }
inline cudaEvent_t as_event(void* p) { return static_cast<cudaEvent_t>(p); }
}  // namespace

int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

context::result<context> context::create(uint32_t width, uint32_t height, context_options options) {
    if (options.device_ordinal < 0 || options.device_ordinal >= device_count()) {
        return std::unexpected{context_error::device_not_found};
    }
    if (cudaSetDevice(options.device_ordinal) != cudaSuccess) {
        return std::unexpected{context_error::device_not_found};
    }

    context ctx;
    ctx.valid_ = true;
    ctx.width_ = width;
    ctx.height_ = height;
    ctx.present_ = options.present;
    ctx.keep_accumulation_ = options.keep_accumulation;
    ctx.persistent_.struct_size = sizeof(Q_cuda_persistent_state);
    ctx.device_desc_.struct_size = sizeof(Q_cuda_device_desc);
    ctx.device_desc_.device_ordinal = options.device_ordinal;
    ctx.device_desc_.shard_index = options.shard_index;
    ctx.device_desc_.shard_count = options.shard_count ? options.shard_count : 1u;

    cudaStream_t q = nullptr;
    if (cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking) != cudaSuccess) {
        return std::unexpected{context_error::stream_failed};
    }
    ctx.queue_ = q;
    for (auto& s : ctx.slots_) {
        cudaStream_t st = nullptr;
        cudaEvent_t b = nullptr, e = nullptr;
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreate(&b) != cudaSuccess || cudaEventCreate(&e) != cudaSuccess) {
            return std::unexpected{context_error::stream_failed};  // ctx's destructor cleans up
        }
        s.stream = st;
        s.begin_event = b;
        s.end_event = e;
    }
    if (ctx.present_ && !ctx.allocate_drawables()) {
        return std::unexpected{context_error::allocation_failed};
    }
    return ctx;
}

context::~context() { release(); }

context::context(context&& other) noexcept { *this = std::move(other); }

context& context::operator=(context&& other) noexcept {
    if (this != &other) {
        release();
        device_desc_ = other.device_desc_;
        queue_ = std::exchange(other.queue_, nullptr);
        slots_ = std::exchange(other.slots_, {});
        next_slot_ = other.next_slot_;
        last_completed_slot_ = other.last_completed_slot_;
        width_ = other.width_;
        height_ = other.height_;
        present_ = other.present_;
        keep_accumulation_ = other.keep_accumulation_;
        persistent_ = std::exchange(other.persistent_, {});
        valid_ = std::exchange(other.valid_, false);
    }
    // The plugin keeps the address of gpu_ctx_ and reads device_desc_ through it,
    // so both must point into THIS object. A context must not be moved after a
    // plugin has been created against it.
    gpu_ctx_.backend = valid_ ? k_backend_cuda : k_backend_none;
    gpu_ctx_.device = valid_ ? &device_desc_ : nullptr;
    gpu_ctx_.queue = queue_;
    gpu_ctx_.layer = (valid_ && keep_accumulation_) ? &persistent_ : nullptr;
    return *this;
}

void context::release() {
    if (!valid_) return;
    cudaSetDevice(device_desc_.device_ordinal);
    synchronize();
    free_drawables();
    if (persistent_.valid && persistent_.layers) cudaFree(persistent_.layers);  // parked by a destroyed plugin
    persistent_ = {};
    for (auto& s : slots_) {
        if (s.begin_event) cudaEventDestroy(as_event(s.begin_event));
        if (s.end_event) cudaEventDestroy(as_event(s.end_event));
        if (s.stream) cudaStreamDestroy(as_stream(s.stream));
        s = slot{};
    }
    if (queue_) cudaStreamDestroy(as_stream(queue_));
    queue_ = nullptr;
    valid_ = false;
}

bool context::allocate_drawables() {
    const size_t bytes = static_cast<size_t>(width_) * height_ * 4;
    for (auto& s : slots_) {
        void* p = nullptr;
        if (bytes != 0 && cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
            (void)cudaGetLastError();
            return false;
        }
        s.pixels = static_cast<uint8_t*>(p);
        // The plugin writes the drawable in HBM; end_frame brings it to the pinned
        // buffer with the copy engine, which overlaps the next frame's kernels. (A
        // kernel storing 8-bit pixels straight into pinned host memory crosses
        // PCIe at a third of the DMA rate and made the frame PCIe-bound.)
        void* d = nullptr;
        if (bytes != 0 && cudaMalloc(&d, bytes) != cudaSuccess) {
            (void)cudaGetLastError();
            return false;
        }
        s.device_pixels = d;
        s.drawable.pixels = d;
        s.drawable.pitch_bytes = 4u * width_;
        s.drawable.reserved = 0;
    }
    return true;
}

void context::free_drawables() {
    for (auto& s : slots_) {
        if (s.pixels) cudaFreeHost(s.pixels);
        if (s.device_pixels) cudaFree(s.device_pixels);
        s.pixels = nullptr;
        s.device_pixels = nullptr;
        s.drawable = Q_cuda_drawable{};
    }
}

void context::resize(uint32_t width, uint32_t height) {
    if (width == width_ && height == height_) return;
    synchronize();
    width_ = width;
    height_ = height;
    if (present_) {
        free_drawables();
        (void)allocate_drawables();
    }
}

bool context::slot_available() const {
    const slot& s = slots_[next_slot_];
    return !s.in_flight || cudaEventQuery(as_event(s.end_event)) == cudaSuccess;
}

bool context::idle() const {
    for (const auto& s : slots_) {
        if (s.in_flight && cudaEventQuery(as_event(s.end_event)) != cudaSuccess) return false;
    }
    return true;
}

async::task<void> context::wait_slot() {
    while (!slot_available()) co_await async::yield();
}

async::task<void> context::wait_idle() {
    while (!idle()) co_await async::yield();
}

void context::synchronize() {
    if (!valid_) return;
    cudaSetDevice(device_desc_.device_ordinal);
    for (auto& s : slots_) {
        if (s.stream) cudaStreamSynchronize(as_stream(s.stream));
        s.in_flight = false;
    }
    if (queue_) cudaStreamSynchronize(as_stream(queue_));
}

context::result<render_frame> context::begin_frame() {
    if (!valid_) return std::unexpected{context_error::no_drawable};
    cudaSetDevice(device_desc_.device_ordinal);
    slot& s = slots_[next_slot_];
    if (s.in_flight) {
        // The semaphore wait of context.mm:138: at most three frames in flight.
        if (cudaEventSynchronize(as_event(s.end_event)) != cudaSuccess) {
            return std::unexpected{context_error::no_drawable};
        }
        s.in_flight = false;
        last_completed_slot_ = next_slot_;
    }
    cudaEventRecord(as_event(s.begin_event), as_stream(s.stream));

    render_frame frame{};
    frame.drawable = (present_ && s.pixels) ? &s.drawable : nullptr;
    frame.command_buffer = s.stream;
    frame.width = width_;
    frame.height = height_;
    frame.camera_dirty = 0;
    return frame;
}

void context::end_frame(render_frame& frame) {
    if (!valid_) return;
    for (int i = 0; i < k_max_frames_in_flight; i++) {
        slot& s = slots_[i];
        if (s.stream == frame.command_buffer) {
            if (frame.drawable && s.device_pixels && s.pixels) {
                // "presentDrawable" (context.mm:166-188): the finished image goes to the host.
                cudaMemcpyAsync(s.pixels, s.device_pixels, static_cast<size_t>(width_) * height_ * 4,
                                cudaMemcpyDeviceToHost, as_stream(s.stream));
            }
            cudaEventRecord(as_event(s.end_event), as_stream(s.stream));
            s.in_flight = true;
            next_slot_ = (i + 1) % k_max_frames_in_flight;
            return;
        }
    }
}

float context::last_frame_gpu_ms() {
    // Newest completed slot = walk back from the slot before next_slot_.
    for (int back = 1; back <= k_max_frames_in_flight; back++) {
        int i = (next_slot_ - back + 2 * k_max_frames_in_flight) % k_max_frames_in_flight;
        slot& s = slots_[i];
        if (!s.end_event) continue;
        if (s.in_flight && cudaEventQuery(as_event(s.end_event)) != cudaSuccess) continue;
        float ms = -1.0f;
        if (cudaEventElapsedTime(&ms, as_event(s.begin_event), as_event(s.end_event)) == cudaSuccess) {
            last_completed_slot_ = i;
            return ms;
        }
        (void)cudaGetLastError();
    }
    return -1.0f;
}

const uint8_t* context::presented_pixels() {
    if (!present_) return nullptr;
    for (int back = 1; back <= k_max_frames_in_flight; back++) {
        int i = (next_slot_ - back + 2 * k_max_frames_in_flight) % k_max_frames_in_flight;
        slot& s = slots_[i];
        if (!s.in_flight || cudaEventQuery(as_event(s.end_event)) == cudaSuccess) {
            if (s.pixels) return s.pixels;
        }
    }
    return nullptr;
}

async::task<void> wait_event(void* cuda_event) {
    while (cudaEventQuery(as_event(cuda_event)) == cudaErrorNotReady) co_await async::yield();
    (void)cudaGetLastError();
}

}  // namespace Q::gpu::cuda
