/// @file fake_plugin.cpp
/// @brief A do-nothing backend that implements the plugin C ABI, for host tests.
///
/// The reference has no such fixture (no plugin .so is ever
/// loaded in a test), which is why its loader and manager are untested. Every
/// entry point reports itself through the host's log callback so a test can
/// assert the exact call sequence. FAKE_PLUGIN_GENERATION is baked in at build
/// time so that a hot-reload test can tell two builds apart.

#include <quasi_plugin_abi.h>

#include <cstdio>
#include <cstdlib>
#include <new>

#ifndef FAKE_PLUGIN_GENERATION
#define FAKE_PLUGIN_GENERATION 1
#endif

#ifndef FAKE_PLUGIN_ABI
#define FAKE_PLUGIN_ABI Q_PLUGIN_ABI_VERSION
#endif

namespace {

struct state {
    Q_plugin_context* ctx;
    unsigned renders = 0;
    unsigned width = 0, height = 0;
};

void say(Q_plugin_context* ctx, const char* what) {
    if (ctx && ctx->log) {
        char buf[96];
        std::snprintf(buf, sizeof buf, "fake%d:%s", FAKE_PLUGIN_GENERATION, what);
        ctx->log(ctx->host_data, buf);
    }
}

}  // namespace

extern "C" {

Q_PLUGIN_EXPORT uint32_t Q_plugin_abi_version(void) { return FAKE_PLUGIN_ABI; }

Q_PLUGIN_EXPORT Q_plugin_info Q_plugin_get_info(void) {
    return Q_plugin_info{"Fake Plugin", {0, FAKE_PLUGIN_GENERATION, 0}, "test fixture", "tests"};
}

Q_PLUGIN_EXPORT Q_plugin_handle* Q_plugin_create(Q_plugin_context* ctx) {
    if (!ctx) return nullptr;
    if (ctx->viewport_width == 0xdead) {  // lets a test provoke create_failed
        say(ctx, "create-refused");
        return nullptr;
    }
    auto* s = new (std::nothrow) state{ctx};
    say(ctx, "create");
    return reinterpret_cast<Q_plugin_handle*>(s);
}

Q_PLUGIN_EXPORT void Q_plugin_destroy(Q_plugin_handle* handle) {
    if (!handle) return;
    auto* s = reinterpret_cast<state*>(handle);
    say(s->ctx, "destroy");
    delete s;
}

Q_PLUGIN_EXPORT void Q_plugin_update(Q_plugin_handle* handle, float) {
    if (handle) say(reinterpret_cast<state*>(handle)->ctx, "update");
}

Q_PLUGIN_EXPORT void Q_plugin_render(Q_plugin_handle* handle, Q_render_frame* frame) {
    if (!handle || !frame) return;
    auto* s = reinterpret_cast<state*>(handle);
    s->renders++;
    s->width = frame->width;
    s->height = frame->height;
    say(s->ctx, "render");
}

#ifndef FAKE_PLUGIN_NO_READBACK
Q_PLUGIN_EXPORT Q_readback_result Q_plugin_readback(Q_plugin_handle* handle) {
    Q_readback_result r{};
    if (!handle) return r;
    auto* s = reinterpret_cast<state*>(handle);
    if (s->renders == 0) return r;
    size_t n = static_cast<size_t>(s->width) * s->height * 4;
    r.data = static_cast<float*>(std::malloc(n * sizeof(float)));
    for (size_t i = 0; i < n; i++) r.data[i] = static_cast<float>(s->renders);
    r.width = s->width;
    r.height = s->height;
    r.channels = 4;
    return r;
}

Q_PLUGIN_EXPORT void Q_plugin_readback_free(Q_readback_result* result) {
    if (result && result->data) {
        std::free(result->data);
        result->data = nullptr;
    }
}
#endif

}  // extern "C"
