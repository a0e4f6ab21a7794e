/// @file manager.hpp
/// @brief Plugin lifecycle with hot reload driven by the coroutine scheduler.
///
/// API-compatible with /root/reference/src/quasi/plugin/manager.hpp:30-385: same
/// hooks, stats, error enum, load_sync / watch_and_reload_loop / reload_async,
/// same reload sequence (pre_unload hook, destroy + dlclose, settle, copy the
/// library to a unique temp path, dlopen, create, post_load hook).
///
/// Two deliberate differences, both required to drive GPUs from the scheduler:
///  - The settle wait is `co_await wait_ms(...)`, not a blocking sleep
///    (reference manager.hpp:230 stalls the whole scheduler thread for 300 ms,
///    which would also stall every other GPU's frame coroutine).
///  - Hooks are awaited, so a pre_unload hook may itself wait on GPU events.
/// The reference host never instantiates this class;
/// quasi_b200/host/src/main.cpp does.

#pragma once

#include <quasi/async/async.hpp>
#include <quasi/plugin/dynamic_library.hpp>
#include <quasi/plugin/loader.hpp>
#include <quasi/plugin/plugin_interface.hpp>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <expected>
#include <filesystem>
#include <functional>
#include <iostream>
#include <optional>
#include <string>
#include <string_view>
#include <system_error>
#include <utility>

namespace Q::plugin {

struct reload_hooks {
    /// Runs before the old plugin is destroyed (e.g. drain the GPU).
    std::function<async::task<void>()> pre_unload = []() -> async::task<void> { co_return; };
    /// Runs after the new plugin has been created.
    std::function<async::task<void>()> post_load = []() -> async::task<void> { co_return; };
    std::function<void(const std::string&)> on_error = [](const std::string&) {};
};

struct reload_stats {
    uint64_t reload_count = 0;
    uint64_t success_count = 0;
    uint64_t failure_count = 0;
    float last_reload_time = 0.0f;  ///< seconds
};

class manager {
public:
    using path_type = std::filesystem::path;

    enum class error {
        file_not_found,
        load_failed,
        abi_mismatch,
        symbol_missing,
        create_failed,
        already_loading,
    };

    template <typename T>
    using result = std::expected<T, error>;

    explicit manager(path_type library_path, reload_hooks hooks = {})
        : library_path_{std::move(library_path)}, watcher_{library_path_}, hooks_{std::move(hooks)} {}

    ~manager() {
        unload_current();
        remove_temp_copy();
    }

    manager(const manager&) = delete;
    manager& operator=(const manager&) = delete;
    manager(manager&&) = delete;
    manager& operator=(manager&&) = delete;

    /// Loads the plugin now. The context setters must have been called first.
    [[nodiscard]] result<void> load_sync() { return load_from_disk(); }

    /// Spawn this on the scheduler: polls the library file every poll interval
    /// and reloads when its modification time moves. Never returns.
    [[nodiscard]] async::task<void> watch_and_reload_loop() {
        log() << "Watching: " << library_path_ << "\n";
        for (;;) {
            co_await async::wait_ms(poll_interval_ms_);
            if (!watcher_.has_changed()) continue;
            log() << "File changed: " << library_path_ << "\n";
            (void)co_await reload();
        }
    }

    [[nodiscard]] async::task<result<void>> reload_async() { co_return co_await reload(); }

    void update(float delta_time) {
        if (plugin_) plugin_->update(delta_time);
    }

    void render(Q::gpu::render_frame* frame) {
        if (plugin_) plugin_->render(frame);
    }

    void set_viewport(uint32_t width, uint32_t height) {
        context_.viewport_width = width;
        context_.viewport_height = height;
    }
    void set_host_data(void* data) { context_.host_data = data; }
    void set_gpu_context(Q::gpu::gpu_context* gpu) { context_.gpu = gpu; }
    void set_log_callback(void (*fn)(void*, const char*)) { context_.log = fn; }
    void set_shutdown_callback(void (*fn)(void*)) { context_.request_shutdown = fn; }

    /// Time to let the linker finish writing before the new file is opened
    /// (reference: 300 ms, manager.hpp:230) and mtime poll period (100 ms, :126).
    void set_timing(int64_t settle_ms, int64_t poll_interval_ms) {
        settle_ms_ = settle_ms;
        poll_interval_ms_ = poll_interval_ms;
    }

    [[nodiscard]] bool is_loaded() const noexcept { return plugin_.has_value() && plugin_->is_valid(); }
    [[nodiscard]] bool is_reloading() const noexcept { return reloading_; }

    [[nodiscard]] std::optional<plugin_info> info() const {
        if (plugin_) return plugin_->info();
        return std::nullopt;
    }

    /// The live plugin (readback, stats); nullptr while unloaded.
    [[nodiscard]] loader* plugin() noexcept { return plugin_ ? &*plugin_ : nullptr; }

    /// A backend-specific entry point beyond the ten ABI symbols (e.g. the
    /// Q_cuda_* functions of quasi_cuda_backend.h); nullptr if absent or unloaded.
    template <typename Fn>
    [[nodiscard]] Fn extension(std::string_view name) const {
        if (!library_.is_loaded()) return nullptr;
        auto sym = library_.get_symbol<Fn>(name);
        return sym ? *sym : nullptr;
    }

    [[nodiscard]] const reload_stats& stats() const noexcept { return stats_; }
    [[nodiscard]] const path_type& library_path() const noexcept { return library_path_; }

private:
    static std::ostream& log() { return std::cout << "[plugin::manager] "; }

    async::task<result<void>> reload() {
        if (reloading_) co_return std::unexpected{error::already_loading};
        reloading_ = true;
        const auto started = std::chrono::steady_clock::now();
        ++stats_.reload_count;

        log() << "Reload " << stats_.reload_count << ": pre-unload hook\n";
        co_await hooks_.pre_unload();
        unload_current();

        co_await async::wait_ms(settle_ms_);  // other coroutines keep running meanwhile

        result<void> loaded = load_from_disk();
        if (!loaded) {
            ++stats_.failure_count;
            hooks_.on_error(std::string{"Load failed: "} + std::string{to_string_impl(loaded.error())});
            log() << "Reload failed: " << to_string_impl(loaded.error()) << "\n";
            watcher_.refresh_timestamp();  // wait for the NEXT write instead of spinning on this one
            reloading_ = false;
            co_return loaded;
        }

        co_await hooks_.post_load();
        ++stats_.success_count;
        stats_.last_reload_time =
            std::chrono::duration<float>(std::chrono::steady_clock::now() - started).count();
        log() << "Reload complete in " << stats_.last_reload_time * 1000.0f << " ms\n";
        watcher_.refresh_timestamp();
        reloading_ = false;
        co_return result<void>{};
    }

    /// Copies the library to a unique temp path and loads THAT, so the linker may
    /// overwrite the original while it is in use, and so that dlopen sees a new
    /// inode (glibc would otherwise hand back the cached image).
    result<void> load_from_disk() {
        std::error_code ec;
        if (!std::filesystem::exists(library_path_, ec)) return std::unexpected{error::file_not_found};

        remove_temp_copy();
        static std::atomic<uint64_t> serial{0};
        path_type copy = std::filesystem::temp_directory_path(ec) /
                         ("quasi_plugin_" + std::to_string(serial++) + "_" + library_path_.filename().string());
        std::filesystem::copy_file(library_path_, copy, std::filesystem::copy_options::overwrite_existing, ec);
        if (ec) {
            log() << "Copy failed: " << ec.message() << "\n";
            return std::unexpected{error::load_failed};
        }
        temp_copy_ = copy;

        auto lib = dynamic_library::open(copy);
        if (!lib) {
            log() << "dlopen failed: " << dynamic_library::last_error() << "\n";
            return std::unexpected{error::load_failed};
        }
        library_ = std::move(*lib);

        auto loaded = loader::load(library_, &context_);
        if (!loaded) {
            log() << "Plugin load failed: " << Q::plugin::to_string(loaded.error()) << "\n";
            library_.close();
            switch (loaded.error()) {
                case loader::error::abi_mismatch: return std::unexpected{error::abi_mismatch};
                case loader::error::symbol_not_found: return std::unexpected{error::symbol_missing};
                case loader::error::create_failed: return std::unexpected{error::create_failed};
            }
            return std::unexpected{error::load_failed};
        }
        plugin_.emplace(std::move(*loaded));

        if (auto i = info()) {
            log() << "Loaded: " << i->name << " v" << i->version.major << "." << i->version.minor << "."
                  << i->version.patch << "\n";
        }
        watcher_.refresh_timestamp();
        return {};
    }

    void unload_current() {
        if (plugin_) {
            plugin_->destroy();
            plugin_.reset();
        }
        if (library_.is_loaded()) library_.close();
    }

    void remove_temp_copy() {
        if (!temp_copy_.empty()) {
            std::error_code ec;
            std::filesystem::remove(temp_copy_, ec);
            temp_copy_.clear();
        }
    }

public:
    static constexpr std::string_view to_string_impl(error e) noexcept {
        switch (e) {
            case error::file_not_found: return "file not found";
            case error::load_failed: return "load failed";
            case error::abi_mismatch: return "ABI version mismatch";
            case error::symbol_missing: return "missing symbols";
            case error::create_failed: return "plugin creation failed";
            case error::already_loading: return "already loading";
        }
        return "unknown error";
    }

private:
    path_type library_path_;
    path_type temp_copy_;
    dynamic_library library_;
    async::file_watcher watcher_;
    reload_hooks hooks_;
    reload_stats stats_;
    plugin_context context_{};
    std::optional<loader> plugin_;
    int64_t settle_ms_ = 300;
    int64_t poll_interval_ms_ = 100;
    bool reloading_ = false;
};

[[nodiscard]] inline constexpr std::string_view to_string(manager::error e) noexcept {
    return manager::to_string_impl(e);
}

}  // namespace Q::plugin
