#include "common-model.h"

#include <numbers>

namespace rtweekend::detail {

// Same construction as the reference camera (common-model.cpp:136-154): orthonormal basis from lookfrom/lookat/vup,
// viewport from the vertical field of view, everything scaled by the focus distance, lens radius = aperture / 2.
Camera::Camera(point lookfrom, point lookat, vec3 vup, double fov, double aspect_ratio, double aperture,
               std::optional<double> focus_dist, time_t t0, time_t t1) {
  const vec3 w = normalize(lookfrom - lookat);
  const vec3 u = normalize(cross(vup, w));
  const vec3 v = normalize(cross(w, u));
  const double viewport_height = 2.0 * std::tan(fov * std::numbers::pi / 180 / 2);
  const double viewport_width = aspect_ratio * viewport_height;
  const double fd = focus_dist ? *focus_dist : length(lookfrom - lookat);
  const vec3 horizontal = fd * viewport_width * u;
  const vec3 vertical = fd * viewport_height * v;
  const point lower_left = lookfrom - horizontal / 2.0 - vertical / 2.0 - fd * w;
  for (std::size_t k = 0; k < 3; ++k) {
    block_.origin[k] = lookfrom[k];
    block_.lower_left[k] = lower_left[k];
    block_.horizontal[k] = horizontal[k];
    block_.vertical[k] = vertical[k];
    block_.u[k] = u[k];
    block_.v[k] = v[k];
  }
  block_.lens_radius = aperture / 2;
  block_.t0 = t0;
  block_.t1 = t1;
}

}  // namespace rtweekend::detail
