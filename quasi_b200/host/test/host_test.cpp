/// @file host_test.cpp
/// @brief Tests of what the reference leaves untested: loader and manager against
/// a real plugin library, hot reload through the scheduler, nested coroutines.
///
/// Needs FAKE_PLUGIN_V1 / FAKE_PLUGIN_V2 / FAKE_PLUGIN_BAD_ABI / FAKE_PLUGIN_MINIMAL
/// in the environment (paths of builds of fake_plugin.cpp; see the Makefile).

#include <quasi/io/exr_writer.hpp>
#include <quasi/plugin/plugin.hpp>

#include <catch2/catch_test_macros.hpp>

#include <chrono>
#include <cstdlib>
#include <filesystem>
#include <string>
#include <thread>
#include <vector>

using namespace Q::plugin;
namespace async = Q::async;

namespace {

std::vector<std::string> g_log;
void log_cb(void*, const char* m) { g_log.emplace_back(m); }

std::filesystem::path env_path(const char* name) {
    const char* v = std::getenv(name);
    return v ? std::filesystem::path{v} : std::filesystem::path{};
}

plugin_context make_context(Q::gpu::gpu_context* gpu) {
    plugin_context ctx{};
    ctx.viewport_width = 8;
    ctx.viewport_height = 6;
    ctx.gpu = gpu;
    ctx.log = log_cb;
    return ctx;
}

}  // namespace

TEST_CASE("loader binds a real library and drives the ten entry points", "[plugin][loader]") {
    g_log.clear();
    auto lib = dynamic_library::open(env_path("FAKE_PLUGIN_V1"));
    REQUIRE(lib.has_value());
    Q::gpu::gpu_context gpu{};
    auto ctx = make_context(&gpu);
    {
        auto p = loader::load(*lib, &ctx);
        REQUIRE(p.has_value());
        REQUIRE(p->is_valid());
        REQUIRE(p->abi_version() == k_plugin_abi_version);
        REQUIRE(std::string{p->info().name} == "Fake Plugin");
        REQUIRE(p->supports_readback());
        REQUIRE_FALSE(p->supports_readback_aov());  // the fixture omits the AOV pair: optional symbols
        REQUIRE(p->readback().data == nullptr);      // nothing rendered yet
        Q::gpu::render_frame frame{};
        frame.width = 8;
        frame.height = 6;
        p->update(0.016f);
        p->render(&frame);
        p->render(nullptr);  // ignored by the plugin
        auto rb = p->readback();
        REQUIRE(rb.data != nullptr);
        REQUIRE(rb.width == 8);
        REQUIRE(rb.channels == 4);
        REQUIRE(rb.data[0] == 1.0f);
        REQUIRE(Q::io::write_exr(std::filesystem::temp_directory_path() / "quasi_host_test.exr", rb).has_value());
        p->readback_free(&rb);
        REQUIRE(rb.data == nullptr);
        auto moved = std::move(*p);  // exactly one destroy, from the moved-to object
        REQUIRE_FALSE(p->is_valid());
        REQUIRE(moved.is_valid());
    }
    REQUIRE(g_log == std::vector<std::string>{"fake1:create", "fake1:update", "fake1:render", "fake1:destroy"});
}

TEST_CASE("loader error cases", "[plugin][loader]") {
    Q::gpu::gpu_context gpu{};
    auto ctx = make_context(&gpu);
    SECTION("ABI outside [1, k_plugin_abi_version] is rejected before create") {
        g_log.clear();
        auto lib = dynamic_library::open(env_path("FAKE_PLUGIN_BAD_ABI"));
        REQUIRE(lib.has_value());
        auto p = loader::load(*lib, &ctx);
        REQUIRE_FALSE(p.has_value());
        REQUIRE(p.error() == loader::error::abi_mismatch);
        REQUIRE(g_log.empty());
    }
    SECTION("a library without the required symbols") {
        auto lib = dynamic_library::open("/lib/x86_64-linux-gnu/libm.so.6");
        REQUIRE(lib.has_value());
        auto p = loader::load(*lib, &ctx);
        REQUIRE_FALSE(p.has_value());
        REQUIRE(p.error() == loader::error::symbol_not_found);
    }
    SECTION("create returning nullptr") {
        auto lib = dynamic_library::open(env_path("FAKE_PLUGIN_V1"));
        REQUIRE(lib.has_value());
        ctx.viewport_width = 0xdead;
        auto p = loader::load(*lib, &ctx);
        REQUIRE_FALSE(p.has_value());
        REQUIRE(p.error() == loader::error::create_failed);
        REQUIRE(std::string{to_string(p.error())} == "Plugin creation failed");
    }
    SECTION("optional readback symbols may be absent") {
        auto lib = dynamic_library::open(env_path("FAKE_PLUGIN_MINIMAL"));
        REQUIRE(lib.has_value());
        auto p = loader::load(*lib, &ctx);
        REQUIRE(p.has_value());
        REQUIRE_FALSE(p->supports_readback());
        REQUIRE(p->readback().data == nullptr);
    }
}

TEST_CASE("manager loads from a temp copy and hot-reloads through the scheduler", "[plugin][manager]") {
    g_log.clear();
    const auto dir = std::filesystem::temp_directory_path() / "quasi_manager_test";
    std::filesystem::create_directories(dir);
    const auto lib_path = dir / "libplugin_under_test.so";
    std::filesystem::copy_file(env_path("FAKE_PLUGIN_V1"), lib_path,
                               std::filesystem::copy_options::overwrite_existing);

    int pre = 0, post = 0;
    reload_hooks hooks;
    hooks.pre_unload = [&]() -> async::task<void> {
        ++pre;
        co_await async::yield();  // hooks may suspend; the manager awaits them
        co_return;
    };
    hooks.post_load = [&]() -> async::task<void> {
        ++post;
        co_return;
    };

    Q::gpu::gpu_context gpu{};
    manager mgr{lib_path, std::move(hooks)};
    mgr.set_viewport(8, 6);
    mgr.set_gpu_context(&gpu);
    mgr.set_log_callback(log_cb);
    mgr.set_timing(/*settle_ms=*/20, /*poll_interval_ms=*/5);
    REQUIRE_FALSE(mgr.is_loaded());
    REQUIRE(mgr.load_sync().has_value());
    REQUIRE(mgr.is_loaded());
    REQUIRE(mgr.info()->version.minor == 1);

    async::scheduler sched;
    sched.spawn(mgr.watch_and_reload_loop());
    int other_coroutine_ticks = 0;
    sched.spawn([](int& n) -> async::task<void> {
        for (;;) {
            ++n;
            co_await async::yield();
        }
    }(other_coroutine_ticks));

    Q::gpu::render_frame frame{};
    frame.width = 8;
    frame.height = 6;
    for (int i = 0; i < 3; i++) {
        sched.tick();
        mgr.update(0.0f);
        mgr.render(&frame);
    }
    REQUIRE(mgr.stats().reload_count == 0);

    // "Rebuild" the plugin: replace the file (new inode, new mtime).
    std::this_thread::sleep_for(std::chrono::milliseconds{20});
    std::filesystem::remove(lib_path);
    std::filesystem::copy_file(env_path("FAKE_PLUGIN_V2"), lib_path);
    std::filesystem::last_write_time(lib_path, std::filesystem::file_time_type::clock::now());

    const int ticks_before = other_coroutine_ticks;
    const auto give_up = std::chrono::steady_clock::now() + std::chrono::seconds{10};
    bool saw_unloaded = false;
    while (mgr.stats().success_count == 0 && std::chrono::steady_clock::now() < give_up) {
        sched.tick();
        saw_unloaded = saw_unloaded || !mgr.is_loaded();
        mgr.render(&frame);  // a no-op while no plugin is loaded (manager.hpp:153-157)
    }
    REQUIRE(mgr.stats().reload_count == 1);
    REQUIRE(mgr.stats().success_count == 1);
    REQUIRE(mgr.stats().failure_count == 0);
    REQUIRE(mgr.stats().last_reload_time > 0.0f);
    REQUIRE(saw_unloaded);
    REQUIRE(pre == 1);
    REQUIRE(post == 1);
    REQUIRE(mgr.is_loaded());
    REQUIRE(mgr.info()->version.minor == 2);                   // the new build is the one running
    REQUIRE(other_coroutine_ticks - ticks_before > 3);          // the settle wait did not block the scheduler

    // destroy of generation 1 happened before create of generation 2
    auto find = [&](const std::string& s) { return std::find(g_log.begin(), g_log.end(), s) - g_log.begin(); };
    REQUIRE(find("fake1:destroy") < find("fake2:create"));
    REQUIRE(find("fake2:create") < static_cast<long>(g_log.size()));

    SECTION("a broken rebuild is counted and leaves no plugin loaded") {
        std::this_thread::sleep_for(std::chrono::milliseconds{20});
        std::filesystem::remove(lib_path);
        std::filesystem::copy_file(env_path("FAKE_PLUGIN_BAD_ABI"), lib_path);
        std::filesystem::last_write_time(lib_path, std::filesystem::file_time_type::clock::now());
        const auto limit = std::chrono::steady_clock::now() + std::chrono::seconds{10};
        while (mgr.stats().failure_count == 0 && std::chrono::steady_clock::now() < limit) sched.tick();
        REQUIRE(mgr.stats().failure_count == 1);
        REQUIRE_FALSE(mgr.is_loaded());
        mgr.render(&frame);  // harmless
    }
    std::filesystem::remove_all(dir);
}

TEST_CASE("manager reports a missing library", "[plugin][manager]") {
    manager mgr{"/nonexistent/libnothing.so"};
    auto r = mgr.load_sync();
    REQUIRE_FALSE(r.has_value());
    REQUIRE(r.error() == manager::error::file_not_found);
    REQUIRE(to_string(r.error()) == "file not found");
}

TEST_CASE("nested tasks finish without the scheduler touching freed frames", "[async]") {
    async::scheduler sched;
    int done = 0;
    bool flag = false;
    auto inner = [&]() -> async::task<int> {
        co_await async::wait_until([&] { return flag; });
        co_return 7;
    };
    auto outer = [&]() -> async::task<void> {
        int v = co_await inner();
        co_await async::wait_ms(1);
        done = v;
    };
    sched.spawn(outer());
    for (int i = 0; i < 5; i++) sched.tick();
    REQUIRE(done == 0);
    flag = true;
    sched.run_until_empty();
    REQUIRE(done == 7);
    REQUIRE(sched.empty());
}

TEST_CASE("when_all and when_any", "[async]") {
    async::scheduler sched;
    int a = 0, b = 0;
    std::size_t first = 99;
    auto count_to = [](int& n, int limit) -> async::task<void> {
        while (n < limit) {
            ++n;
            co_await async::yield();
        }
    };
    // A named closure: a coroutine lambda must outlive the coroutine it creates.
    auto driver = [&]() -> async::task<void> {
        co_await async::when_all(count_to(a, 3), count_to(b, 6));
        auto fast = count_to(a, 5);
        auto slow = count_to(b, 40);
        first = co_await async::when_any(fast, slow);
        while (!slow.done()) co_await async::yield();  // keep it alive until it finishes
    };
    sched.spawn(driver());
    sched.run_until_empty();
    REQUIRE(a == 5);
    REQUIRE(b == 40);
    REQUIRE(first == 0);
}
