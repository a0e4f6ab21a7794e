/// @file awaitables.hpp
/// @brief yield, predicate and timer waits, when_all / when_any.
///
/// API-compatible with /root/reference/src/quasi/async/awaitables.hpp:20-137.
/// Every wait is a yield loop: a waiting coroutine costs one predicate call per
/// scheduler tick and never blocks the thread.

#pragma once

#include <quasi/async/scheduler.hpp>
#include <quasi/async/task.hpp>

#include <chrono>
#include <concepts>
#include <coroutine>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <tuple>
#include <utility>

namespace Q::async {

struct yield_awaitable {
    bool await_ready() const noexcept { return false; }
    void await_suspend(std::coroutine_handle<> h) const noexcept {
        if (scheduler* s = current_scheduler()) s->enqueue(h);
    }
    void await_resume() const noexcept {}
};

/// Suspends until the next scheduler tick.
[[nodiscard]] inline yield_awaitable yield() noexcept { return {}; }

using std::suspend_always;
using std::suspend_never;

template <std::invocable<> Pred>
    requires std::convertible_to<std::invoke_result_t<Pred>, bool>
task<void> wait_until(Pred pred) {
    while (!std::invoke(pred)) co_await yield();
}

template <typename Rep, typename Period>
task<void> wait_for(std::chrono::duration<Rep, Period> duration) {
    const auto deadline = std::chrono::steady_clock::now() + duration;
    while (std::chrono::steady_clock::now() < deadline) co_await yield();
}

inline task<void> wait_ms(int64_t ms) { return wait_for(std::chrono::milliseconds{ms}); }
inline task<void> wait_us(int64_t us) { return wait_for(std::chrono::microseconds{us}); }

/// Starts every task and completes when all of them have. The tasks' own yields
/// keep them on the scheduler's queue; this coroutine only polls.
template <typename... Tasks>
task<void> when_all(Tasks... tasks) {
    (tasks.resume(), ...);
    while (!(tasks.done() && ...)) co_await yield();
}

/// Starts every task and completes with the position of the first one to finish.
/// The caller keeps the tasks alive until they are done (unfinished ones are
/// still queued on the scheduler).
template <typename... Tasks>
task<std::size_t> when_any(Tasks&... tasks) {
    (tasks.resume(), ...);
    for (;;) {
        std::size_t index = 0, winner = sizeof...(Tasks);
        ((winner == sizeof...(Tasks) && tasks.done() ? winner = index : 0, ++index), ...);
        if (winner != sizeof...(Tasks)) co_return winner;
        co_await yield();
    }
}

}  // namespace Q::async
