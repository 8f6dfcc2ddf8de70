// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the subset of GLM that the
// reference's hot-path translation units use, so that /root/reference/src/*.cpp can be
// compiled UNMODIFIED into oracle/_ref/ (glm/cci.20220420 is a Conan dependency of the
// reference, conanfile.txt:2, and is not vendored in the mount; there is no network).
//
// Formulas follow GLM's published generic implementations:
//   dot(a,b)        = (a*b).x + (a*b).y + (a*b).z            (detail/func_geometric.inl compute_dot<3>)
//   length(v)       = sqrt(dot(v,v))
//   inversesqrt(x)  = 1 / sqrt(x)
//   normalize(v)    = v * inversesqrt(dot(v,v))
//   reflect(I,N)    = I - N * dot(N,I) * 2
//   refract(I,N,e)  : k = 1 - e*e*(1 - dot(N,I)^2);  k >= 0 ? e*I - (e*dot(N,I) + sqrt(k))*N : 0
//   epsilonEqual    = |a-b| < eps (component-wise),  all(bvec3) = x && y && z
//   length2(v)      = dot(v,v)                                (gtx/norm)
// Call sites served: vec3.h:3-8, render.cpp:14,125, random-utils.cpp:26,37,
// common-model.cpp:16,26,43,45,55,57,71-73,86,88,108-115,128-131,139-141,148, main.cpp:49.
#pragma once
#include <cmath>
#include <cstddef>

namespace glm {

template <typename T>
struct tvec3 {
  union { T x, r; };
  union { T y, g; };
  union { T z, b; };

  constexpr tvec3() : x(0), y(0), z(0) {}
  template <typename A, typename B, typename C>
  constexpr tvec3(A a, B b_, C c) : x(static_cast<T>(a)), y(static_cast<T>(b_)), z(static_cast<T>(c)) {}
  // GLM's cross-type converting constructor is implicit unless GLM_FORCE_EXPLICIT_CTOR.
  template <typename U>
  constexpr tvec3(const tvec3<U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)) {}

  constexpr T& operator[](std::size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
  constexpr const T& operator[](std::size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }

  constexpr tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  constexpr tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  constexpr tvec3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
  constexpr tvec3& operator/=(T s) { x /= s; y /= s; z /= s; return *this; }
};

struct bvec3 { bool x, y, z; };

using dvec3 = tvec3<double>;
using vec3 = tvec3<float>;

template <typename T> constexpr tvec3<T> operator+(const tvec3<T>& a, const tvec3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> constexpr tvec3<T> operator-(const tvec3<T>& a, const tvec3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> constexpr tvec3<T> operator*(const tvec3<T>& a, const tvec3<T>& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <typename T> constexpr tvec3<T> operator/(const tvec3<T>& a, const tvec3<T>& b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
template <typename T> constexpr tvec3<T> operator-(const tvec3<T>& a) { return {-a.x, -a.y, -a.z}; }
template <typename T> constexpr tvec3<T> operator*(const tvec3<T>& a, T s) { return {a.x * s, a.y * s, a.z * s}; }
template <typename T> constexpr tvec3<T> operator*(T s, const tvec3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename T> constexpr tvec3<T> operator/(const tvec3<T>& a, T s) { return {a.x / s, a.y / s, a.z / s}; }
template <typename T> constexpr tvec3<T> operator+(const tvec3<T>& a, T s) { return {a.x + s, a.y + s, a.z + s}; }
template <typename T> constexpr tvec3<T> operator-(const tvec3<T>& a, T s) { return {a.x - s, a.y - s, a.z - s}; }

template <typename T> constexpr T dot(const tvec3<T>& a, const tvec3<T>& b) {
  tvec3<T> tmp = a * b;
  return tmp.x + tmp.y + tmp.z;
}
template <typename T> constexpr tvec3<T> cross(const tvec3<T>& a, const tvec3<T>& b) {
  return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
template <typename T> inline T length(const tvec3<T>& v) { return std::sqrt(dot(v, v)); }
template <typename T> inline T length2(const tvec3<T>& v) { return dot(v, v); }
template <typename T> inline T inversesqrt(T x) { return static_cast<T>(1) / std::sqrt(x); }
template <typename T> inline tvec3<T> normalize(const tvec3<T>& v) { return v * inversesqrt(dot(v, v)); }
template <typename T> inline tvec3<T> sqrt(const tvec3<T>& v) { return {std::sqrt(v.x), std::sqrt(v.y), std::sqrt(v.z)}; }
template <typename T> inline tvec3<T> reflect(const tvec3<T>& I, const tvec3<T>& N) {
  return I - N * dot(N, I) * static_cast<T>(2);
}
template <typename T> inline tvec3<T> refract(const tvec3<T>& I, const tvec3<T>& N, T eta) {
  T const d = dot(N, I);
  T const k = static_cast<T>(1) - eta * eta * (static_cast<T>(1) - d * d);
  return (k >= static_cast<T>(0)) ? (eta * I - (eta * d + std::sqrt(k)) * N) : tvec3<T>{};
}
template <typename T> inline bvec3 epsilonEqual(const tvec3<T>& a, const tvec3<T>& b, T eps) {
  return {std::abs(a.x - b.x) < eps, std::abs(a.y - b.y) < eps, std::abs(a.z - b.z) < eps};
}
inline bool all(const bvec3& v) { return v.x && v.y && v.z; }

}  // namespace glm
