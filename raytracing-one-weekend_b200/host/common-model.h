// common-model.h -- camera, materials and the pointer-stable store of the host scene model.
// API-compatible with the reference's common-model.h for everything a caller of render() touches
// (Camera ctor common-model.h:95-98, Lambertian/Metal/Dielectric ctors :124,132,143, OOStore::add :158-162).
// Unlike the reference these are plain data: hit()/scatter()/get_ray() live in the CUDA kernels, so the classes
// expose the fields the flattener needs instead of virtual behaviour.
#pragma once
#include <algorithm>
#include <deque>
#include <memory>
#include <optional>
#include <type_traits>

#include "../../include/rtw_b200.h"
#include "vec3.h"

namespace rtweekend::detail {

using time_t = double;

class Camera {
 public:
  Camera(point lookfrom, point lookat, vec3 vup, double fov, double aspect_ratio, double aperture,
         std::optional<double> focus_dist = std::nullopt, time_t t0 = 0, time_t t1 = 0);
  // the ten derived fields of the reference camera (common-model.h:104-112), ready for the device
  [[nodiscard]] const rtw_camera& block() const { return block_; }

 private:
  rtw_camera block_{};
};

enum class MaterialKind : int { lambertian = RTW_LAMBERTIAN, metal = RTW_METAL, dielectric = RTW_DIELECTRIC };

class Material {
 public:
  virtual ~Material() = default;
  [[nodiscard]] virtual rtw_material flat() const = 0;
};

struct Lambertian final : Material {
  explicit Lambertian(const color& a) : albedo{a} {}
  [[nodiscard]] rtw_material flat() const override { return {RTW_LAMBERTIAN, 0, {albedo.x, albedo.y, albedo.z}, 0.0, 0.0}; }
  color albedo;
};

struct Metal final : Material {
  explicit Metal(const color& a, double f = 0) : albedo{a}, fuzz{std::clamp(f, 0.0, 1.0)} {}
  [[nodiscard]] rtw_material flat() const override { return {RTW_METAL, 0, {albedo.x, albedo.y, albedo.z}, fuzz, 0.0}; }
  color albedo;
  double fuzz;
};

struct Dielectric final : Material {
  explicit Dielectric(double index_of_refraction, double f = 0) : ir{index_of_refraction}, fuzz{std::clamp(f, 0.0, 1.0)} {}
  [[nodiscard]] rtw_material flat() const override { return {RTW_DIELECTRIC, 0, {1.0, 1.0, 1.0}, fuzz, ir}; }
  double ir;
  double fuzz;
};

// Append-only store whose elements never move: callers keep Material& / Material* across later add() calls
// (main.cpp:38-39,48-69).
template <typename Base>
class OOStore {
  std::deque<std::unique_ptr<Base>> items_;

 public:
  template <typename Derived, typename... Args>
  Derived& add(Args&&... args) {
    static_assert(std::is_base_of_v<Base, Derived>);
    auto p = std::make_unique<Derived>(std::forward<Args>(args)...);
    Derived& ref = *p;
    items_.push_back(std::move(p));
    return ref;
  }
  [[nodiscard]] auto size() const { return items_.size(); }
  [[nodiscard]] auto begin() const { return items_.begin(); }
  [[nodiscard]] auto end() const { return items_.end(); }
  [[nodiscard]] auto cbegin() const { return items_.cbegin(); }
  [[nodiscard]] auto cend() const { return items_.cend(); }
  [[nodiscard]] const Base& operator[](std::size_t i) const { return *items_[i]; }
};

}  // namespace rtweekend::detail

namespace rtweekend {
using detail::Camera;
using detail::Dielectric;
using detail::Lambertian;
using detail::Material;
using detail::Metal;
using detail::OOStore;
}  // namespace rtweekend
