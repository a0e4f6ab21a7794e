/// @file dynamic_library.hpp
/// @brief Move-only owner of a dlopen handle.
///
/// API-compatible with /root/reference/src/quasi/plugin/dynamic_library.hpp:19-178
/// (same names, same error enumerators and strings, RTLD_NOW | RTLD_LOCAL).

#pragma once

#include <quasi/platform.hpp>

#include <dlfcn.h>

#include <expected>
#include <filesystem>
#include <string>
#include <string_view>
#include <system_error>
#include <utility>

namespace Q::plugin {

enum class library_error {
    file_not_found,
    load_failed,
    symbol_not_found,
    not_loaded,
};

[[nodiscard]] constexpr std::string_view to_string(library_error err) noexcept {
    switch (err) {
        case library_error::file_not_found: return "file not found";
        case library_error::load_failed: return "failed to load library";
        case library_error::symbol_not_found: return "symbol not found";
        case library_error::not_loaded: return "library not loaded";
    }
    return "unknown error";
}

[[nodiscard]] constexpr std::string_view shared_library_extension() noexcept {
    return Q_SHARED_LIB_EXTENSION;
}

/// Appends the platform suffix unless the path already names a shared library.
[[nodiscard]] inline std::filesystem::path ensure_library_extension(std::filesystem::path path) {
    const auto ext = path.extension();
    if (ext == ".so" || ext == ".dylib" || ext == ".dll") return path;
    path += Q_SHARED_LIB_EXTENSION;
    return path;
}

class dynamic_library {
public:
    using handle_type = void*;
    using path_type = std::filesystem::path;
    template <typename T>
    using result = std::expected<T, library_error>;

    dynamic_library() noexcept = default;
    ~dynamic_library() { close(); }

    dynamic_library(const dynamic_library&) = delete;
    dynamic_library& operator=(const dynamic_library&) = delete;

    dynamic_library(dynamic_library&& other) noexcept { swap(other); }
    dynamic_library& operator=(dynamic_library&& other) noexcept {
        if (this != &other) {
            close();
            swap(other);
        }
        return *this;
    }

    [[nodiscard]] static result<dynamic_library> open(const path_type& path) {
        std::error_code ec;
        if (!std::filesystem::exists(path, ec)) return std::unexpected{library_error::file_not_found};
        (void)::dlerror();
        handle_type handle = ::dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!handle) return std::unexpected{library_error::load_failed};
        dynamic_library lib;
        lib.handle_ = handle;
        lib.path_ = path;
        return lib;
    }

    template <typename FuncPtr>
    [[nodiscard]] result<FuncPtr> get_symbol(std::string_view name) const {
        if (!handle_) return std::unexpected{library_error::not_loaded};
        (void)::dlerror();
        void* sym = ::dlsym(handle_, std::string{name}.c_str());
        if (::dlerror() != nullptr) return std::unexpected{library_error::symbol_not_found};
        return reinterpret_cast<FuncPtr>(sym);
    }

    void close() noexcept {
        if (handle_) {
            ::dlclose(handle_);
            handle_ = nullptr;
            path_.clear();
        }
    }

    [[nodiscard]] bool is_loaded() const noexcept { return handle_ != nullptr; }
    [[nodiscard]] explicit operator bool() const noexcept { return is_loaded(); }
    [[nodiscard]] const path_type& path() const noexcept { return path_; }
    [[nodiscard]] handle_type native_handle() const noexcept { return handle_; }

    [[nodiscard]] static const char* last_error() noexcept {
        const char* e = ::dlerror();
        return e ? e : "no error";
    }

private:
    void swap(dynamic_library& other) noexcept {
        std::swap(handle_, other.handle_);
        std::swap(path_, other.path_);
    }

    handle_type handle_ = nullptr;
    path_type path_;
};

}  // namespace Q::plugin
