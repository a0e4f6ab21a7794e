/// @file gpu.hpp
/// @brief Umbrella header for the GPU layer.

#pragma once

#include <quasi/gpu/types.hpp>
