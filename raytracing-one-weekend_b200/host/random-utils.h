// random-utils.h -- host random numbers used by the scene builders.
// Same generator and distributions as the reference (random-utils.cpp:6-22: one process-wide std::mt19937 with the
// default seed, libstdc++ uniform_real_distribution<double> / uniform_int_distribution<int>) because the cover scene's
// sphere positions and materials ARE this stream.  The per-sample stream inside the kernels is Philox;
// random_in_unit_sphere/disk/unit_vector have no host counterpart here (they only ever fed the ray loop).
#pragma once
#include <cstdint>

#include "vec3.h"

namespace rtweekend::detail {
void seed_host_rng(std::uint32_t seed);  // extension: the reference never reseeds
double random_double(double a = 0, double b = 1.0);
int random_int(int a = 0, int b = 1);
color random_vec3(double min = 0, double max = 1.0);
}  // namespace rtweekend::detail

namespace rtweekend {
using detail::random_double;
using detail::random_int;
using detail::random_vec3;
using detail::seed_host_rng;
}  // namespace rtweekend
