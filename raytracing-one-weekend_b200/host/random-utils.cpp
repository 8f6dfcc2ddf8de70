// random-utils.cpp -- host-side random numbers, used ONLY while scenes are built (the kernels draw from Philox).
// The cover scene's layout is defined by the sequence the reference's generator produces: a default-seeded
// std::mt19937 (seed 5489) read through libstdc++'s uniform_real_distribution<double> (two 32-bit draws per double) and
// uniform_int_distribution<int>.  Keeping that sequence keeps every sphere position, material and albedo identical.
#include "random-utils.h"

#include <random>

namespace rtweekend::detail {
namespace {

// One process-wide engine, like the reference, but behind a type so that the seed hook and the draws share it explicitly.
class HostRng {
 public:
  static HostRng& instance() {
    static HostRng rng;
    return rng;
  }
  void reseed(std::uint32_t seed) { engine_.seed(seed); }
  double real(double lo, double hi) {
    std::uniform_real_distribution<double> dist(lo, hi);
    return dist(engine_);
  }
  int integer(int lo, int hi) {
    std::uniform_int_distribution<int> dist(lo, hi);
    return dist(engine_);
  }

 private:
  std::mt19937 engine_;  // default-constructed: seed 5489
};

}  // namespace

void seed_host_rng(std::uint32_t seed) { HostRng::instance().reseed(seed); }

double random_double(double a, double b) { return HostRng::instance().real(a, b); }

int random_int(int a, int b) { return HostRng::instance().integer(a, b); }

color random_vec3(double min, double max) {
  // three draws in x, y, z order (the reference's braced initialiser evaluates left to right)
  HostRng& rng = HostRng::instance();
  color v;
  v.x = rng.real(min, max);
  v.y = rng.real(min, max);
  v.z = rng.real(min, max);
  return v;
}

}  // namespace rtweekend::detail
