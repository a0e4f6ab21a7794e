/// @file catch2_main.cpp
/// @brief main() for test binaries built against the mini Catch2 stand-in.

#include <catch2/catch_test_macros.hpp>

int main(int argc, char** argv) { return mini_catch::run_all(argc > 1 ? argv[1] : nullptr); }
