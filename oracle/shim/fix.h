// Force-included (-include) when building oracle/_ref from the reference sources.
// (1) rasterizer_impl.h:40,59 use std::uintptr_t / uint32_t without <cstdint>.
// (2) gcc-13's M_1_PIf32 / M_2_PIf32 are _Float32 literals, which nvcc 12.9 cannot
//     code-generate ("unsupported float variant"); re-define them as plain float literals
//     with the same value (auxiliary.h:230-231,244-245; forward.cu:154-155; backward.cu:352-353).
#pragma once
#include <cstdint>
#include <math.h>
#undef M_1_PIf32
#undef M_2_PIf32
#define M_1_PIf32 0.318309886183790671537767526745028724f
#define M_2_PIf32 0.636619772367581343075535053490057448f
