// primitive-model.h -- Sphere / MovingSphere / Triangle and their store.
// One model instead of the reference's two interchangeable ones (oo-primitives.h, variant-primitives.h selected by
// primitive-model.h:1-4): constructors and accessors match both (oo-primitives.h:28,49,76,37-43,60-67), the
// store offers the same add<T>(args...) -> T&.  Primitives are plain records that know how to flatten themselves;
// intersection and bounding boxes are computed on the device / in the BVH builder.
#pragma once
#include <cstdint>
#include <vector>

#include "common-model.h"

namespace rtweekend::detail {

class Primitive {
 public:
  explicit Primitive(const Material& m) : material_{&m} {}
  virtual ~Primitive() = default;
  [[nodiscard]] const Material& material() const { return *material_; }
  // kind, geometry; the material index is filled in by Scene::flatten
  [[nodiscard]] virtual rtw_primitive flat() const = 0;

 private:
  const Material* material_;
};

class Sphere final : public Primitive {
 public:
  Sphere(point center, double radius, const Material& material) : Primitive{material}, center_{center}, radius_{radius} {}
  [[nodiscard]] const point& center() const { return center_; }
  [[nodiscard]] const double& radius() const { return radius_; }
  [[nodiscard]] rtw_primitive flat() const override {
    return {RTW_SPHERE, 0, {center_.x, center_.y, center_.z}, {center_.x, center_.y, center_.z}, {0, 0, 0}, radius_};
  }

 private:
  point center_;
  double radius_;
};

class MovingSphere final : public Primitive {
 public:
  MovingSphere(point c0, point c1, double radius, const Material& material)
      : Primitive{material}, center0_{c0}, center1_{c1}, radius_{radius} {}
  [[nodiscard]] const point& center() const { return center0_; }
  // shutter runs over [0,1] (oo-primitives.h:51-52)
  [[nodiscard]] point center(time_t time) const { return center0_ + time * (center1_ - center0_); }
  [[nodiscard]] const double& radius() const { return radius_; }
  [[nodiscard]] rtw_primitive flat() const override {
    return {RTW_MOVING_SPHERE, 0, {center0_.x, center0_.y, center0_.z}, {center1_.x, center1_.y, center1_.z}, {0, 0, 0}, radius_};
  }

 private:
  point center0_, center1_;
  double radius_;
};

class Triangle final : public Primitive {
 public:
  Triangle(point a, point b, point c, const Material& material) : Primitive{material}, a_{a}, b_{b}, c_{c} {}
  [[nodiscard]] const point& a() const { return a_; }
  [[nodiscard]] const point& b() const { return b_; }
  [[nodiscard]] const point& c() const { return c_; }
  [[nodiscard]] rtw_primitive flat() const override {
    return {RTW_TRIANGLE, 0, {a_.x, a_.y, a_.z}, {b_.x, b_.y, b_.z}, {c_.x, c_.y, c_.z}, 0.0};
  }

 private:
  point a_, b_, c_;
};

}  // namespace rtweekend::detail

namespace rtweekend {
using PrimitiveStore_t = detail::OOStore<detail::Primitive>;
using MaterialStore_t = detail::OOStore<detail::Material>;
using detail::MovingSphere;
using detail::Sphere;
using detail::Triangle;
}  // namespace rtweekend
