/// @file scheduler.hpp
/// @brief Single-threaded cooperative scheduler: one FIFO pass per tick.
///
/// API-compatible with /root/reference/src/quasi/async/scheduler.hpp:20-156.
/// tick() resumes exactly the coroutines that were ready when it was called, in
/// spawn order; anything they enqueue runs on the next tick. The host's frame
/// loop calls tick() once per frame, so nothing spawned here may block: GPU
/// waits are expressed as polling awaitables (quasi/gpu/cuda/context.hpp).

#pragma once

#include <quasi/async/task.hpp>

#include <coroutine>
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

namespace Q::async {

class scheduler;

namespace detail {
inline thread_local scheduler* t_current_scheduler = nullptr;
}

[[nodiscard]] inline scheduler* current_scheduler() noexcept { return detail::t_current_scheduler; }

class scheduler {
public:
    scheduler() = default;
    ~scheduler() {
        // Only spawned (root) coroutines belong to the scheduler. Anything else in
        // the ready queue is a nested task whose frame is owned by a task object
        // inside one of those roots and dies with it.
        for (auto h : roots_) {
            if (h) h.destroy();
        }
    }

    scheduler(const scheduler&) = delete;
    scheduler& operator=(const scheduler&) = delete;
    scheduler(scheduler&&) = delete;
    scheduler& operator=(scheduler&&) = delete;

    /// Takes ownership of a task; it first runs on the next tick().
    void spawn(task<void> t) {
        if (t.valid() && !t.done()) {
            std::coroutine_handle<> h = t.release();
            roots_.push_back(h);
            ready_.push_back(h);
        }
    }

    /// Re-queues a suspended coroutine (used by yield()). Ownership stays put.
    void enqueue(std::coroutine_handle<> h) {
        if (h && !h.done()) ready_.push_back(h);
    }

    void tick() {
        ++tick_count_;
        struct scope {
            scheduler* previous;
            explicit scope(scheduler* self) : previous{std::exchange(detail::t_current_scheduler, self)} {}
            ~scope() { detail::t_current_scheduler = previous; }
        } guard{this};

        std::vector<std::coroutine_handle<>> batch;
        batch.swap(ready_);
        for (auto h : batch) {
            // A queued handle is always suspended at a yield, so it is alive here.
            // After resume() it may have finished AND been freed by the task object
            // that owns it (a nested wait_until, say), so it is not touched again.
            h.resume();
            reap_finished_roots();
        }
    }

    void run_until_empty() {
        while (!ready_.empty()) tick();
    }

    [[nodiscard]] bool empty() const noexcept { return ready_.empty(); }
    [[nodiscard]] std::size_t size() const noexcept { return ready_.size(); }
    [[nodiscard]] uint64_t tick_count() const noexcept { return tick_count_; }

private:
    void reap_finished_roots() {
        for (std::size_t i = 0; i < roots_.size();) {
            if (roots_[i].done()) {
                roots_[i].destroy();
                roots_[i] = roots_.back();
                roots_.pop_back();
            } else {
                ++i;
            }
        }
    }

    std::vector<std::coroutine_handle<>> ready_;  ///< resumed, in order, by the next tick
    std::vector<std::coroutine_handle<>> roots_;  ///< spawned coroutines the scheduler owns
    uint64_t tick_count_ = 0;
};

inline scheduler& default_scheduler() {
    static scheduler instance;
    return instance;
}

inline void spawn(task<void> t) { default_scheduler().spawn(std::move(t)); }

}  // namespace Q::async
