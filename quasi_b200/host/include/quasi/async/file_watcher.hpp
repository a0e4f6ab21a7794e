/// @file file_watcher.hpp
/// @brief Modification-time polling for hot reload.
///
/// API-compatible with /root/reference/src/quasi/async/file_watcher.hpp:21-216.
/// A watcher remembers the mtime it last saw; has_changed() is true while the
/// file exists with a different mtime, until refresh_timestamp() is called.

#pragma once

#include <quasi/async/awaitables.hpp>
#include <quasi/async/task.hpp>

#include <filesystem>
#include <initializer_list>
#include <optional>
#include <system_error>
#include <utility>
#include <vector>

namespace Q::async {

struct file_change {
    std::filesystem::path path;
    std::filesystem::file_time_type old_time;
    std::filesystem::file_time_type new_time;
};

class file_watcher {
public:
    using path_type = std::filesystem::path;
    using time_type = std::filesystem::file_time_type;

    explicit file_watcher(path_type path) : path_{std::move(path)} { refresh_timestamp(); }

    [[nodiscard]] bool has_changed() const {
        const auto now = current_mtime();
        return now.has_value() && seen_.has_value() && *now != *seen_;
    }

    void refresh_timestamp() { seen_ = current_mtime(); }

    /// Awaitable form: resumes with the change once the mtime moves.
    [[nodiscard]] task<file_change> next_change() {
        co_await wait_until([this] { return has_changed(); });
        co_return take_change();
    }

    /// Awaitable that completes immediately if a change is pending, else suspends
    /// the caller (who is expected to be resumed by its own polling loop).
    [[nodiscard]] auto wait_for_change() {
        struct awaiter {
            file_watcher* self;
            [[nodiscard]] bool await_ready() const { return self->has_changed(); }
            [[nodiscard]] bool await_suspend(std::coroutine_handle<>) const { return !self->has_changed(); }
            [[nodiscard]] file_change await_resume() const { return self->take_change(); }
        };
        return awaiter{this};
    }

    [[nodiscard]] const path_type& path() const noexcept { return path_; }
    [[nodiscard]] std::optional<time_type> last_write_time() const noexcept { return seen_; }
    [[nodiscard]] bool exists() const {
        std::error_code ec;
        return std::filesystem::exists(path_, ec);
    }

    /// Reports the pending change and starts watching from the new mtime.
    [[nodiscard]] file_change take_change() {
        file_change change{path_, seen_.value_or(time_type{}), current_mtime().value_or(time_type{})};
        refresh_timestamp();
        return change;
    }

private:
    [[nodiscard]] std::optional<time_type> current_mtime() const {
        std::error_code ec;
        if (!std::filesystem::exists(path_, ec)) return std::nullopt;
        const auto t = std::filesystem::last_write_time(path_, ec);
        if (ec) return std::nullopt;
        return t;
    }

    path_type path_;
    std::optional<time_type> seen_;
};

class multi_file_watcher {
public:
    using path_type = std::filesystem::path;

    void add(path_type path) { watchers_.emplace_back(std::move(path)); }
    void add(std::initializer_list<path_type> paths) {
        for (const auto& p : paths) add(p);
    }

    /// First watched file whose mtime moved, if any; that file is re-armed.
    [[nodiscard]] std::optional<file_change> poll_change() {
        for (auto& w : watchers_) {
            if (w.has_changed()) return w.take_change();
        }
        return std::nullopt;
    }

    [[nodiscard]] task<file_change> next_change() {
        for (;;) {
            if (auto c = poll_change()) co_return *c;
            co_await yield();
        }
    }

    void refresh_all() {
        for (auto& w : watchers_) w.refresh_timestamp();
    }

    [[nodiscard]] std::size_t size() const noexcept { return watchers_.size(); }
    [[nodiscard]] bool empty() const noexcept { return watchers_.empty(); }

private:
    std::vector<file_watcher> watchers_;
};

}  // namespace Q::async
