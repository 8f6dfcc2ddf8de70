#include "scenes.h"

#include <cstdio>
#include <stdexcept>

#include "obj-loader.h"
#include "random-utils.h"

namespace rtweekend {

// Cover scene.  The layout IS the host random stream, so the draws happen in the order the reference's binary makes
// them: choose_mat, then the two position draws, of which g++ evaluates the z one first (the reference writes them as
// arguments of one constructor call, main.cpp:46, whose evaluation order the compiler picks right to left); then the
// material's draws.  tests/test_host.py::test_cover_scene_matches_oracle_scene pins every sphere and material against the reference build.
Scene lots_of_balls(const Config& cfg) {
  Scene world{Camera{point(13, 2, 3), point(0, 0, 0), vec3(0, 1, 0), 20.0, cfg.aspect_ratio, 0.1, 10.0, 0, 1}};
  auto& shop = world.boutique();
  auto& things = world.primitives();

  things.add<Sphere>(point{0, -1000, 0}, 1000.0, shop.add<Lambertian>(color{0.5, 0.5, 0.5}));

  const int n = cfg.number_of_balls_sqrt;
  for (int a = -n; a < n; ++a) {
    for (int b = -n; b < n; ++b) {
      const double pick = random_double();
      const double jz = random_double();
      const double jx = random_double();
      const point center(a + 0.9 * jx, 0.2, b + 0.9 * jz);
      if (!(length(center - point{4, 0.2, 0}) > 0.9)) continue;
      if (pick < 0.8) {  // diffuse, moving upwards during the shutter interval if requested
        const color c1 = random_vec3();
        const color c2 = random_vec3();
        const Material& m = shop.add<Lambertian>(c1 * c2);
        if (cfg.moving_spheres) {
          const point center2 = center + point(0, random_double(0, .5), 0);
          things.add<MovingSphere>(center, center2, 0.2, m);
        } else {
          things.add<Sphere>(center, 0.2, m);
        }
      } else if (pick < 0.95) {  // metal
        const color albedo = random_vec3(0.5, 1);
        const double fuzz = random_double(0, 0.5);
        things.add<Sphere>(center, 0.2, shop.add<Metal>(albedo, fuzz));
      } else {  // glass
        things.add<Sphere>(center, 0.2, shop.add<Dielectric>(1.5));
      }
    }
  }
  const Material& glass = shop.add<Dielectric>(1.5);
  const Material& reddish = shop.add<Lambertian>(color{0.4, 0.2, 0.1});
  const Material& reddish_metal = shop.add<Metal>(color{0.7, 0.6, 0.5});
  things.add<Sphere>(point(0, 1, 0), 1.0, glass);
  things.add<Sphere>(point(-4, 1, 0), 1.0, reddish);
  things.add<Sphere>(point(4, 1, 0), 1.0, reddish_metal);
  return world;
}

// OBJ scene: grey Lambertian mesh and nothing else, camera (1,0,-1) -> origin, 35 degrees, aperture 0.01, focus on the
// origin, shutter [0,1].  One host random_int() is consumed first, as in the reference (main.cpp:86).
Scene foo(const Config& cfg) {
  random_int();
  Scene world{Camera{point(1, 0, -1), point(0, 0, 0), vec3(0, 1, 0), 35.0, cfg.aspect_ratio, 0.01, std::nullopt, 0, 1}};
  const Material& grey = world.boutique().add<Lambertian>(color{0.5, 0.5, 0.5});
  if (!cfg.model) throw std::runtime_error("foo(): Config::model is not set");
  const detail::ObjMesh mesh = detail::load_obj(*cfg.model, device_options().obj_all_shapes);
  for (const auto& f : mesh.faces)
    world.primitives().add<Triangle>(mesh.vertices[static_cast<std::size_t>(f[0])], mesh.vertices[static_cast<std::size_t>(f[1])],
                                     mesh.vertices[static_cast<std::size_t>(f[2])], grey);
  std::fprintf(stderr, "Scene has %zu triangles\n", world.primitives().size());
  return world;
}

// Mesh standing on the r=1000 ground sphere of the cover scene, seen from a three-quarter view.
Scene mesh_on_ground(const Config& cfg) {
  if (!cfg.model) throw std::runtime_error("mesh_on_ground(): Config::model is not set");
  const detail::ObjMesh mesh = detail::load_obj(*cfg.model, device_options().obj_all_shapes);
  double ymin = 1e300, ymax = -1e300;
  for (const auto& v : mesh.vertices) { ymin = std::min(ymin, v.y); ymax = std::max(ymax, v.y); }
  const double lift = -ymin;
  const point target(0, 0.5 * (ymax - ymin), 0);
  Scene world{Camera{point(2.6, 1.7, 4.2), target, vec3(0, 1, 0), 30.0, cfg.aspect_ratio, 0.02, std::nullopt, 0, 1}};
  auto& shop = world.boutique();
  world.primitives().add<Sphere>(point{0, -1000, 0}, 1000.0, shop.add<Lambertian>(color{0.5, 0.5, 0.5}));
  const Material& clay = shop.add<Lambertian>(color{0.7, 0.45, 0.3});
  const vec3 up(0, lift, 0);
  for (const auto& f : mesh.faces)
    world.primitives().add<Triangle>(mesh.vertices[static_cast<std::size_t>(f[0])] + up, mesh.vertices[static_cast<std::size_t>(f[1])] + up,
                                     mesh.vertices[static_cast<std::size_t>(f[2])] + up, clay);
  std::fprintf(stderr, "Scene has %zu primitives\n", world.primitives().size());
  return world;
}

}  // namespace rtweekend
