// oo-primitives.h -- Sphere / MovingSphere / Triangle behind a virtual base, held by OOStore<Primitive> (the default model,
// selected by primitive-model.h like the reference's oo-primitives.h; the other one is variant-primitives.h).
// Constructor signatures and accessor names follow the reference (oo-primitives.h:28,49,76,37-43,60-67) and the store offers
// the same add<T>(args...) -> T&.  A primitive here is nothing but its flat C-ABI record (rtw_primitive) plus the material it
// points at: intersection and bounding boxes are computed on the device and in the BVH builder, so the host classes only have
// to remember their constructor arguments in the form Scene::flatten() hands to rtw_render.  Where the reference dispatches
// hit()/bounding_box() through detail::hit / detail::bounding_box (oo-primitives.h:92-98), this model dispatches the one thing
// the host still needs per primitive -- its flat record and its material -- through detail::flat / detail::material_of.
#pragma once
#include <cstdint>
#include <vector>

#include "common-model.h"

namespace rtweekend::detail {

class Primitive {
 public:
  virtual ~Primitive() = default;
  [[nodiscard]] const Material& material() const { return *material_; }
  // kind + geometry as the C ABI wants them; the material index is filled in by Scene::flatten
  [[nodiscard]] const rtw_primitive& flat() const { return record_; }

 protected:
  Primitive(rtw_prim_kind kind, const Material& m) : material_{&m} { record_.kind = kind; }
  static void store(double (&dst)[3], const point& p) { dst[0] = p.x; dst[1] = p.y; dst[2] = p.z; }
  static point load(const double (&src)[3]) { return point{src[0], src[1], src[2]}; }
  rtw_primitive record_{};

 private:
  const Material* material_;
};

// record_.a = record_.b = centre
class Sphere final : public Primitive {
 public:
  Sphere(point center, double radius, const Material& material) : Primitive{RTW_SPHERE, material} {
    store(record_.a, center); store(record_.b, center); record_.radius = radius;
  }
  [[nodiscard]] point center() const { return load(record_.a); }
  [[nodiscard]] double radius() const { return record_.radius; }
};

// record_.a = centre when the shutter opens (time 0), record_.b = centre when it closes (time 1), oo-primitives.h:51-52
class MovingSphere final : public Primitive {
 public:
  MovingSphere(point c0, point c1, double radius, const Material& material) : Primitive{RTW_MOVING_SPHERE, material} {
    store(record_.a, c0); store(record_.b, c1); record_.radius = radius;
  }
  [[nodiscard]] point center() const { return load(record_.a); }
  [[nodiscard]] point center(time_t time) const {
    const point from = load(record_.a), to = load(record_.b);
    return from + time * (to - from);
  }
  [[nodiscard]] double radius() const { return record_.radius; }
};

// record_.a/b/c = the three vertices in the order given (the winding decides the culled side)
class Triangle final : public Primitive {
 public:
  Triangle(point a, point b, point c, const Material& material) : Primitive{RTW_TRIANGLE, material} {
    store(record_.a, a); store(record_.b, b); store(record_.c, c);
  }
  [[nodiscard]] point a() const { return load(record_.a); }
  [[nodiscard]] point b() const { return load(record_.b); }
  [[nodiscard]] point c() const { return load(record_.c); }
};

// dispatch shims over a store element (a unique_ptr here, a std::variant in variant-primitives.h)
inline const rtw_primitive& flat(const std::unique_ptr<Primitive>& p) { return p->flat(); }
inline const Material& material_of(const std::unique_ptr<Primitive>& p) { return p->material(); }

}  // namespace rtweekend::detail

namespace rtweekend {
using PrimitiveStore_t = detail::OOStore<detail::Primitive>;
using MaterialStore_t = detail::OOStore<detail::Material>;
using detail::MovingSphere;
using detail::Sphere;
using detail::Triangle;
}  // namespace rtweekend
