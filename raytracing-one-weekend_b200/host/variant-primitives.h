// variant-primitives.h -- the non-virtual primitive model: Sphere / MovingSphere / Triangle as plain value types held in a
// VariantStore<Sphere, MovingSphere, Triangle> (a vector of std::variant), enabled by -DRTWEEKEND_USE_VARIANT_PRIMITIVES
// (reference: variant-primitives.h:14-113, selector primitive-model.h:1-2).  Same constructor signatures, accessors and
// add<T>(args...) -> T& as the virtual model in oo-primitives.h; the dispatch shims go through std::visit where the virtual
// model makes a virtual call.  As in the reference, references returned by add() are only valid until the next add()
// (the store is a vector); materials stay in the pointer-stable OOStore.
#pragma once
#include <cstddef>
#include <utility>
#include <variant>
#include <vector>

#include "common-model.h"

namespace rtweekend::detail {

// Common part of the three value types: the flat C-ABI record and the material it points at.
class Primitive {
 public:
  [[nodiscard]] const Material& material() const { return *material_; }
  [[nodiscard]] const rtw_primitive& flat() const { return record_; }

 protected:
  Primitive(rtw_prim_kind kind, const Material& m) : material_{&m} { record_.kind = kind; }
  static void put(double (&dst)[3], const point& p) { dst[0] = p.x; dst[1] = p.y; dst[2] = p.z; }
  static point get(const double (&src)[3]) { return point{src[0], src[1], src[2]}; }
  rtw_primitive record_{};

 private:
  const Material* material_;
};

class Sphere : public Primitive {
 public:
  Sphere(point center, double radius, const Material& material) : Primitive{RTW_SPHERE, material} {
    put(record_.a, center); put(record_.b, center); record_.radius = radius;
  }
  [[nodiscard]] point center() const { return get(record_.a); }
  [[nodiscard]] double radius() const { return record_.radius; }
};

class MovingSphere : public Primitive {
 public:
  MovingSphere(point c0, point c1, double radius, const Material& material) : Primitive{RTW_MOVING_SPHERE, material} {
    put(record_.a, c0); put(record_.b, c1); record_.radius = radius;
  }
  [[nodiscard]] point center() const { return get(record_.a); }
  [[nodiscard]] point center(time_t time) const {  // shutter interval [0, 1] (variant-primitives.h:46-47,59-61)
    const point from = get(record_.a), to = get(record_.b);
    return from + time * (to - from);
  }
  [[nodiscard]] double radius() const { return record_.radius; }
};

class Triangle : public Primitive {
 public:
  Triangle(point a, point b, point c, const Material& material) : Primitive{RTW_TRIANGLE, material} {
    put(record_.a, a); put(record_.b, b); put(record_.c, c);
  }
  [[nodiscard]] point a() const { return get(record_.a); }
  [[nodiscard]] point b() const { return get(record_.b); }
  [[nodiscard]] point c() const { return get(record_.c); }
};

template <typename... T>
class VariantStore {
 public:
  using value_type = std::variant<T...>;

  template <typename U, typename... Args>
  U& add(Args&&... args) {
    items_.emplace_back(std::in_place_type<U>, std::forward<Args>(args)...);
    return std::get<U>(items_.back());
  }
  [[nodiscard]] std::size_t size() const { return items_.size(); }
  [[nodiscard]] auto begin() const { return items_.begin(); }
  [[nodiscard]] auto end() const { return items_.end(); }
  [[nodiscard]] auto cbegin() const { return items_.cbegin(); }
  [[nodiscard]] auto cend() const { return items_.cend(); }
  [[nodiscard]] const value_type* data() const { return items_.data(); }

 private:
  std::vector<value_type> items_;
};

using PrimitiveStore_t = VariantStore<Sphere, MovingSphere, Triangle>;

// dispatch shims over a store element: std::visit where oo-primitives.h makes a virtual call (variant-primitives.h:107-113)
inline const rtw_primitive& flat(const PrimitiveStore_t::value_type& p) {
  return std::visit([](const auto& q) -> const rtw_primitive& { return q.flat(); }, p);
}
inline const Material& material_of(const PrimitiveStore_t::value_type& p) {
  return std::visit([](const auto& q) -> const Material& { return q.material(); }, p);
}

}  // namespace rtweekend::detail

namespace rtweekend {
using detail::PrimitiveStore_t;
using MaterialStore_t = detail::OOStore<detail::Material>;
using detail::MovingSphere;
using detail::Sphere;
using detail::Triangle;
}  // namespace rtweekend
