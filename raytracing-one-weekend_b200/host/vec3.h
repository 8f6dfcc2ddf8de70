// vec3.h -- value type behind rtweekend::vec3 / point / color.
// The reference aliases glm::dvec3 (src/vec3.h:6-8); GLM is a Conan dependency that is not available here, and the
// host side of this renderer only builds scenes and cameras (all ray math lives in the CUDA kernels), so a small
// self-contained double-precision 3-vector is enough.  normalize/length/dot/cross use GLM's operation order so
// that the camera block and the scene layout come out bit-identical to the reference's.
#pragma once
#include <cmath>
#include <cstddef>

namespace rtweekend::detail {

struct dvec3 {
  double x = 0.0, y = 0.0, z = 0.0;
  constexpr dvec3() = default;
  constexpr dvec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
  constexpr double& operator[](std::size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
  constexpr const double& operator[](std::size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
  constexpr dvec3& operator+=(const dvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
};

constexpr dvec3 operator+(const dvec3& a, const dvec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
constexpr dvec3 operator-(const dvec3& a, const dvec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
constexpr dvec3 operator-(const dvec3& a) { return {-a.x, -a.y, -a.z}; }
constexpr dvec3 operator*(const dvec3& a, const dvec3& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
constexpr dvec3 operator*(const dvec3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
constexpr dvec3 operator*(double s, const dvec3& a) { return {s * a.x, s * a.y, s * a.z}; }
constexpr dvec3 operator/(const dvec3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
constexpr double dot(const dvec3& a, const dvec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
constexpr dvec3 cross(const dvec3& a, const dvec3& b) {
  return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline double length(const dvec3& a) { return std::sqrt(dot(a, a)); }
inline dvec3 normalize(const dvec3& a) { return a * (1.0 / std::sqrt(dot(a, a))); }

using vec3 = dvec3;
using color = dvec3;
using point = dvec3;

}  // namespace rtweekend::detail

namespace rtweekend {
using detail::color;
using detail::point;
using detail::vec3;
}  // namespace rtweekend

// Callers written against the reference spell a few vector functions through glm:: (main.cpp:49).
namespace glm {
using rtweekend::detail::cross;
using rtweekend::detail::dot;
using rtweekend::detail::length;
using rtweekend::detail::normalize;
}  // namespace glm
