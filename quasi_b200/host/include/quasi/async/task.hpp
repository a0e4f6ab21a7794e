/// @file task.hpp
/// @brief Lazy, move-only coroutine task.
///
/// API-compatible with /root/reference/src/quasi/async/task.hpp:33-261: a task
/// starts suspended, runs when resumed or awaited, hands control back to its
/// awaiter by symmetric transfer, and rethrows a stored exception from result().

#pragma once

#include <concepts>
#include <coroutine>
#include <exception>
#include <optional>
#include <utility>

namespace Q::async {

template <typename T = void>
class task;

namespace detail {

/// What every task promise shares: lazy start, continuation hand-off, exception slot.
struct promise_base {
    std::coroutine_handle<> continuation_;
    std::exception_ptr exception_;

    std::suspend_always initial_suspend() noexcept { return {}; }

    struct final_awaiter {
        bool await_ready() const noexcept { return false; }
        template <typename Promise>
        std::coroutine_handle<> await_suspend(std::coroutine_handle<Promise> self) const noexcept {
            std::coroutine_handle<> next = self.promise().continuation_;
            return next ? next : std::noop_coroutine();
        }
        void await_resume() const noexcept {}
    };
    final_awaiter final_suspend() noexcept { return {}; }

    void unhandled_exception() noexcept { exception_ = std::current_exception(); }

    void rethrow_if_failed() const {
        if (exception_) std::rethrow_exception(exception_);
    }
};

template <typename T>
struct task_promise : promise_base {
    std::optional<T> value_;

    task<T> get_return_object() noexcept;

    template <typename U>
        requires std::convertible_to<U, T>
    void return_value(U&& value) {
        value_.emplace(std::forward<U>(value));
    }

    [[nodiscard]] T get_result() {
        rethrow_if_failed();
        return std::move(*value_);
    }
    [[nodiscard]] bool has_result() const noexcept { return value_.has_value() || exception_; }
};

template <>
struct task_promise<void> : promise_base {
    bool returned_ = false;

    task<void> get_return_object() noexcept;
    void return_void() noexcept { returned_ = true; }
    void get_result() { rethrow_if_failed(); }
    [[nodiscard]] bool has_result() const noexcept { return returned_ || exception_; }
};

}  // namespace detail

template <typename T>
class [[nodiscard]] task {
public:
    using promise_type = detail::task_promise<T>;
    using handle_type = std::coroutine_handle<promise_type>;

    task() noexcept = default;
    explicit task(handle_type h) noexcept : handle_{h} {}
    ~task() { reset(); }

    task(task&& other) noexcept : handle_{std::exchange(other.handle_, nullptr)} {}
    task& operator=(task&& other) noexcept {
        if (this != &other) {
            reset();
            handle_ = std::exchange(other.handle_, nullptr);
        }
        return *this;
    }
    task(const task&) = delete;
    task& operator=(const task&) = delete;

    // awaitable: co_await task starts it and resumes the awaiter when it finishes
    [[nodiscard]] bool await_ready() const noexcept { return done(); }
    std::coroutine_handle<> await_suspend(std::coroutine_handle<> awaiter) noexcept {
        handle_.promise().continuation_ = awaiter;
        return handle_;
    }
    T await_resume() { return handle_.promise().get_result(); }

    // manual control
    [[nodiscard]] bool valid() const noexcept { return handle_ != nullptr; }
    [[nodiscard]] bool done() const noexcept { return !handle_ || handle_.done(); }
    void resume() {
        if (!done()) handle_.resume();
    }
    T result() { return handle_.promise().get_result(); }
    [[nodiscard]] handle_type release() noexcept { return std::exchange(handle_, nullptr); }
    [[nodiscard]] handle_type handle() const noexcept { return handle_; }

private:
    void reset() noexcept {
        if (handle_) handle_.destroy();
        handle_ = nullptr;
    }

    handle_type handle_ = nullptr;
};

template <typename T>
task<T> detail::task_promise<T>::get_return_object() noexcept {
    return task<T>{std::coroutine_handle<task_promise<T>>::from_promise(*this)};
}

inline task<void> detail::task_promise<void>::get_return_object() noexcept {
    return task<void>{std::coroutine_handle<task_promise<void>>::from_promise(*this)};
}

template <typename T>
task<T> make_ready_task(T value) {
    co_return std::move(value);
}

inline task<void> make_ready_task() { co_return; }

}  // namespace Q::async
