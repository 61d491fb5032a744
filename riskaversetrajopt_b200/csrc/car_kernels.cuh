// placeholder, replaced below
#pragma once
#include "saa_common.cuh"
namespace saa {
template <int S> struct CarRed { static constexpr int N = 4 * (S - 1) + 2 * S + 4; };
}
