/// @file catch_test_macros.hpp
/// @brief A minimal stand-in for the Catch2 v3 macros the Quasi tests use.
///
/// Catch2 is fetched from the Bazel Central Registry by the reference
/// (/root/reference/MODULE.bazel:26) and is not available offline. This header
/// implements just TEST_CASE, SECTION (one level), REQUIRE and REQUIRE_FALSE
/// (+ CHECK variants) so that /root/reference/test/*.cpp compile UNMODIFIED, from
/// where they lie, against this repository's headers, and so that this
/// repository's own C++ tests read like the reference's.

#pragma once

#include <cstdio>
#include <exception>
#include <string>
#include <vector>

namespace mini_catch {

struct test_case {
    const char* name;
    const char* tags;
    void (*body)();
};

inline std::vector<test_case>& registry() {
    static std::vector<test_case> r;
    return r;
}

struct registrar {
    registrar(const char* name, const char* tags, void (*body)()) { registry().push_back({name, tags, body}); }
};

struct assertion_failed : std::exception {};

struct run_state {
    int section_target = 0;   ///< which top-level SECTION this pass enters
    int section_seen = 0;     ///< SECTIONs encountered so far in this pass
    int assertions = 0;
    int failures = 0;
};

inline run_state& state() {
    static run_state s;
    return s;
}

/// SECTION: each pass through the test body enters exactly one section.
struct section_guard {
    bool entered;
    explicit section_guard(const char*) : entered{state().section_seen++ == state().section_target} {}
    explicit operator bool() const { return entered; }
};

inline void report(bool ok, bool fatal, const char* macro, const char* expr, const char* file, int line) {
    ++state().assertions;
    if (ok) return;
    ++state().failures;
    std::fprintf(stderr, "%s:%d: FAILED: %s( %s )\n", file, line, macro, expr);
    if (fatal) throw assertion_failed{};
}

inline int run_all(const char* filter = nullptr) {
    int failed_cases = 0, ran = 0;
    for (const auto& tc : registry()) {
        if (filter && std::string{tc.name}.find(filter) == std::string::npos) continue;
        ++ran;
        const int failures_before = state().failures;
        int target = 0;
        for (;;) {
            state().section_target = target;
            state().section_seen = 0;
            try {
                tc.body();
            } catch (const assertion_failed&) {
            } catch (const std::exception& e) {
                ++state().failures;
                std::fprintf(stderr, "%s: unexpected exception: %s\n", tc.name, e.what());
            }
            if (state().section_seen <= target + 1) break;  // no further section to visit
            ++target;
        }
        if (state().failures != failures_before) {
            ++failed_cases;
            std::fprintf(stderr, "FAILED test case: %s\n", tc.name);
        }
    }
    std::printf("%s: %d test cases, %d assertions, %d failed\n", failed_cases ? "FAILED" : "All tests passed",
                ran, state().assertions, state().failures);
    return failed_cases ? 1 : 0;
}

}  // namespace mini_catch

#define MINI_CATCH_CAT2(a, b) a##b
#define MINI_CATCH_CAT(a, b) MINI_CATCH_CAT2(a, b)

#define TEST_CASE(...) MINI_CATCH_TEST_CASE(MINI_CATCH_CAT(mini_catch_test_, __LINE__), __VA_ARGS__, "")
#define MINI_CATCH_TEST_CASE(fn, name, tags, ...)                                   \
    static void fn();                                                                \
    static ::mini_catch::registrar MINI_CATCH_CAT(fn, _reg){name, tags, &fn};        \
    static void fn()

#define SECTION(name) if (::mini_catch::section_guard MINI_CATCH_CAT(mini_catch_section_, __LINE__){name})

#define REQUIRE(...) ::mini_catch::report(static_cast<bool>(__VA_ARGS__), true, "REQUIRE", #__VA_ARGS__, __FILE__, __LINE__)
#define REQUIRE_FALSE(...) ::mini_catch::report(!static_cast<bool>(__VA_ARGS__), true, "REQUIRE_FALSE", #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK(...) ::mini_catch::report(static_cast<bool>(__VA_ARGS__), false, "CHECK", #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK_FALSE(...) ::mini_catch::report(!static_cast<bool>(__VA_ARGS__), false, "CHECK_FALSE", #__VA_ARGS__, __FILE__, __LINE__)
