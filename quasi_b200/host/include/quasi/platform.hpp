/// @file platform.hpp
/// @brief Platform facts the host needs: shared-library suffix, export attribute.
///
/// Same macro names as /root/reference/src/quasi/platform.hpp so that code written
/// against the reference compiles here; only Linux and macOS are recognised.

#pragma once

#if defined(__linux__)
#define Q_PLATFORM_LINUX 1
#define Q_PLATFORM_POSIX 1
#define Q_PLATFORM_NAME "Linux"
#define Q_SHARED_LIB_EXTENSION ".so"
#elif defined(__APPLE__) && defined(__MACH__)
#define Q_PLATFORM_MACOS 1
#define Q_PLATFORM_POSIX 1
#define Q_PLATFORM_NAME "macOS"
#define Q_SHARED_LIB_EXTENSION ".dylib"
#else
#error "quasi host: unsupported platform (needs dlopen)"
#endif

#if __cplusplus < 202002L
#error "quasi host: C++20 or later required (coroutines)"
#endif

#define Q_HAS_COROUTINES 1
#define Q_HAS_STD_EXPECTED 1

#define Q_EXPORT_API __attribute__((visibility("default")))
#define Q_IMPORT_API
