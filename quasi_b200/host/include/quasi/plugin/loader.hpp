/// @file loader.hpp
/// @brief Binds the ten plugin entry points of a loaded library and owns one instance.
///
/// API-compatible with /root/reference/src/quasi/plugin/loader.hpp:35-297: six
/// required symbols, ABI range check, four optional readback symbols, create;
/// destroy on destruction. The symbol table is data-driven here instead of one
/// block per symbol. The optional Q_cuda_* extension symbols of
/// quasi_cuda_backend.h are bound too when present (cuda_stats()).

#pragma once

#include <quasi/plugin/dynamic_library.hpp>
#include <quasi/plugin/plugin_interface.hpp>

#include <expected>
#include <utility>

namespace Q::plugin {

class loader {
public:
    enum class error {
        abi_mismatch,
        symbol_not_found,
        create_failed,
    };

    template <typename T>
    using result = std::expected<T, error>;

    using abi_version_fn = uint32_t (*)();
    using get_info_fn = plugin_info (*)();
    using create_fn = plugin_handle* (*)(plugin_context*);
    using destroy_fn = void (*)(plugin_handle*);
    using update_fn = void (*)(plugin_handle*, float);
    using render_fn = void (*)(plugin_handle*, Q::gpu::render_frame*);
    using readback_fn = readback_result (*)(plugin_handle*);
    using readback_free_fn = void (*)(readback_result*);
    using readback_aov_fn = readback_aov_result (*)(plugin_handle*);
    using readback_aov_free_fn = void (*)(readback_aov_result*);
    using cuda_stats_fn = int (*)(plugin_handle*, Q_cuda_stats*);

    [[nodiscard]] static result<loader> load(dynamic_library& library, plugin_context* context) {
        loader p;
        symbols& s = p.sym_;
        const bool required_ok = bind(library, k_symbol_abi_version, s.abi_version) &&
                                 bind(library, k_symbol_get_info, s.get_info) &&
                                 bind(library, k_symbol_create, s.create) &&
                                 bind(library, k_symbol_destroy, s.destroy) &&
                                 bind(library, k_symbol_update, s.update) &&
                                 bind(library, k_symbol_render, s.render);
        if (!required_ok) return std::unexpected{error::symbol_not_found};

        const uint32_t abi = s.abi_version();
        if (abi < 1 || abi > k_plugin_abi_version) return std::unexpected{error::abi_mismatch};

        (void)bind(library, k_symbol_readback, s.readback);
        (void)bind(library, k_symbol_readback_free, s.readback_free);
        (void)bind(library, k_symbol_readback_aov, s.readback_aov);
        (void)bind(library, k_symbol_readback_aov_free, s.readback_aov_free);
        (void)bind(library, "Q_cuda_plugin_stats", s.cuda_stats);

        p.handle_ = s.create(context);
        if (!p.handle_) return std::unexpected{error::create_failed};
        return p;
    }

    loader() = default;
    ~loader() { destroy(); }

    loader(loader&& other) noexcept
        : handle_{std::exchange(other.handle_, nullptr)}, sym_{std::exchange(other.sym_, symbols{})} {}
    loader& operator=(loader&& other) noexcept {
        if (this != &other) {
            destroy();
            handle_ = std::exchange(other.handle_, nullptr);
            sym_ = std::exchange(other.sym_, symbols{});
        }
        return *this;
    }
    loader(const loader&) = delete;
    loader& operator=(const loader&) = delete;

    void update(float delta_time) {
        if (handle_ && sym_.update) sym_.update(handle_, delta_time);
    }

    void render(Q::gpu::render_frame* frame) {
        if (handle_ && sym_.render) sym_.render(handle_, frame);
    }

    [[nodiscard]] bool supports_readback() const noexcept {
        return sym_.readback && sym_.readback_free;
    }

    [[nodiscard]] readback_result readback() {
        if (handle_ && sym_.readback) return sym_.readback(handle_);
        return readback_result{nullptr, 0, 0, 0};
    }

    void readback_free(readback_result* result) {
        if (sym_.readback_free && result) sym_.readback_free(result);
    }

    [[nodiscard]] bool supports_readback_aov() const noexcept {
        return sym_.readback_aov && sym_.readback_aov_free;
    }

    [[nodiscard]] readback_aov_result readback_aov() {
        if (handle_ && sym_.readback_aov) return sym_.readback_aov(handle_);
        return readback_aov_result{};
    }

    void readback_aov_free(readback_aov_result* result) {
        if (sym_.readback_aov_free && result) sym_.readback_aov_free(result);
    }

    /// CUDA backend only: ray/path/launch counters. False when the backend has none.
    [[nodiscard]] bool cuda_stats(Q_cuda_stats* out) {
        return handle_ && sym_.cuda_stats && sym_.cuda_stats(handle_, out) == 0;
    }

    [[nodiscard]] plugin_info info() const { return sym_.get_info ? sym_.get_info() : plugin_info{}; }
    [[nodiscard]] uint32_t abi_version() const { return sym_.abi_version ? sym_.abi_version() : 0; }
    [[nodiscard]] bool is_valid() const noexcept { return handle_ != nullptr; }
    /// The opaque instance handle, for backend-specific extension entry points.
    [[nodiscard]] plugin_handle* native_handle() const noexcept { return handle_; }
    [[nodiscard]] explicit operator bool() const noexcept { return is_valid(); }

    void destroy() {
        if (handle_ && sym_.destroy) {
            sym_.destroy(handle_);
            handle_ = nullptr;
        }
    }

private:
    struct symbols {
        abi_version_fn abi_version = nullptr;
        get_info_fn get_info = nullptr;
        create_fn create = nullptr;
        destroy_fn destroy = nullptr;
        update_fn update = nullptr;
        render_fn render = nullptr;
        readback_fn readback = nullptr;
        readback_free_fn readback_free = nullptr;
        readback_aov_fn readback_aov = nullptr;
        readback_aov_free_fn readback_aov_free = nullptr;
        cuda_stats_fn cuda_stats = nullptr;
    };

    template <typename Fn>
    static bool bind(dynamic_library& library, const char* name, Fn& slot) {
        auto sym = library.get_symbol<Fn>(name);
        if (!sym || !*sym) return false;
        slot = *sym;
        return true;
    }

    plugin_handle* handle_ = nullptr;
    symbols sym_{};
};

[[nodiscard]] inline constexpr const char* to_string(loader::error e) noexcept {
    switch (e) {
        case loader::error::abi_mismatch: return "ABI version mismatch";
        case loader::error::symbol_not_found: return "Required symbol not found";
        case loader::error::create_failed: return "Plugin creation failed";
    }
    return "Unknown error";
}

}  // namespace Q::plugin
