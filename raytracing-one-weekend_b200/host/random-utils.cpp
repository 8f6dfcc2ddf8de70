#include "random-utils.h"

#include <random>

namespace rtweekend::detail {
namespace {
std::mt19937& engine() {
  static std::mt19937 e;  // default seed 5489, as the reference's gen()
  return e;
}
}  // namespace

void seed_host_rng(std::uint32_t seed) { engine().seed(seed); }

double random_double(double a, double b) {
  std::uniform_real_distribution<double> d(a, b);
  return d(engine());
}
int random_int(int a, int b) {
  std::uniform_int_distribution<int> d(a, b);
  return d(engine());
}
color random_vec3(double min, double max) {
  const double x = random_double(min, max);
  const double y = random_double(min, max);
  const double z = random_double(min, max);
  return {x, y, z};
}
}  // namespace rtweekend::detail
