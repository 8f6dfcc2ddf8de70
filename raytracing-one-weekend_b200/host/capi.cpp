// capi.cpp -- C hooks onto the host model for the Python tests and bench.py (ctypes): build the product's own scenes,
// flatten them into the C-ABI arrays, run the drop-in render() into a file.  Test plumbing only; the renderer's ABI is
// include/rtw_b200.h.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "obj-loader.h"
#include "random-utils.h"
#include "render.h"
#include "scenes.h"

namespace rt = rtweekend;
#define RTWH_API extern "C" __attribute__((visibility("default")))

namespace {
thread_local std::string g_err;
struct Handle {
  rt::Scene scene;
  rt::Scene::Flat flat;
  explicit Handle(rt::Scene&& s) : scene(std::move(s)) { flat = scene.flatten(); }
};
template <typename F>
Handle* guarded(F&& f) {
  try {
    return new Handle(f());
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
}  // namespace

RTWH_API const char* rtwh_last_error() { return g_err.c_str(); }
RTWH_API const char* rtwh_primitive_model() {   // which of the two primitive models this library was compiled with (primitive-model.h)
#ifdef RTWEEKEND_USE_VARIANT_PRIMITIVES
  return "variant";
#else
  return "oo";
#endif
}
RTWH_API void rtwh_seed(unsigned seed) { rt::seed_host_rng(seed); }
RTWH_API double rtwh_random_double() { return rt::random_double(); }
RTWH_API void rtwh_set_obj_all_shapes(int on) { rt::device_options().obj_all_shapes = on != 0; }

RTWH_API void* rtwh_scene_cover(int nsqrt, double aspect, int moving) {
  rt::Config cfg{};
  cfg.number_of_balls_sqrt = nsqrt; cfg.aspect_ratio = aspect; cfg.moving_spheres = moving != 0;
  return guarded([&] { return rt::lots_of_balls(cfg); });
}
RTWH_API void* rtwh_scene_obj(const char* path, double aspect) {
  rt::Config cfg{};
  cfg.aspect_ratio = aspect; cfg.model = std::string(path);
  return guarded([&] { return rt::foo(cfg); });
}
RTWH_API void* rtwh_scene_mesh_on_ground(const char* path, double aspect) {
  rt::Config cfg{};
  cfg.aspect_ratio = aspect; cfg.model = std::string(path);
  return guarded([&] { return rt::mesh_on_ground(cfg); });
}
RTWH_API void rtwh_scene_free(void* h) { delete static_cast<Handle*>(h); }
RTWH_API long long rtwh_scene_nprims(void* h) { return static_cast<Handle*>(h)->flat.desc.nprims; }
RTWH_API long long rtwh_scene_nmats(void* h) { return static_cast<Handle*>(h)->flat.desc.nmats; }
// copies the flattened arrays out; `desc` receives pointers into the handle's own storage (valid until rtwh_scene_free)
RTWH_API void rtwh_scene_flatten(void* h, rtw_primitive* prims, rtw_material* mats, rtw_scene_desc* desc) {
  auto* s = static_cast<Handle*>(h);
  if (prims) std::memcpy(prims, s->flat.prims.data(), s->flat.prims.size() * sizeof(rtw_primitive));
  if (mats) std::memcpy(mats, s->flat.mats.data(), s->flat.mats.size() * sizeof(rtw_material));
  if (desc) *desc = s->flat.desc;
}
// Camera constructor alone (common-model.cpp:136-154); focus_dist <= 0 means "distance to lookat"
RTWH_API void rtwh_camera(const double from[3], const double at[3], const double vup[3], double fov, double aspect, double aperture,
                          double focus_dist, double t0, double t1, rtw_camera* out) {
  std::optional<double> fd;
  if (focus_dist > 0) fd = focus_dist;
  rt::Camera c{rt::point(from[0], from[1], from[2]), rt::point(at[0], at[1], at[2]), rt::vec3(vup[0], vup[1], vup[2]), fov, aspect, aperture, fd, t0, t1};
  *out = c.block();
}
// the drop-in render() with stdout captured into `ppm_path`; returns 0 on success
RTWH_API int rtwh_render_to_file(void* h, int width, double aspect, int spp, int max_child_rays, int nthreads, const char* ppm_path,
                                 int ngpus, unsigned long long seed, int kernel) {
  auto* s = static_cast<Handle*>(h);
  rt::Config cfg{};
  cfg.image_width = width; cfg.aspect_ratio = aspect; cfg.samples_per_pixel = spp; cfg.max_child_rays = max_child_rays; cfg.nthreads = nthreads;
  auto& dev = rt::device_options();
  dev.ngpus = ngpus; dev.seed = seed; dev.kernel = kernel;
  try {
    std::ofstream out(ppm_path, std::ios::binary);
    if (!out) throw std::runtime_error(std::string("cannot write ") + ppm_path);
    const rt::Accum img = rt::render_accum(s->scene, cfg);
    rt::write_ppm(out, img);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}
RTWH_API int rtwh_config_string(int nsqrt, double aspect, int width, int spp, int moving, int max_child_rays, int nthreads, char* buf, int cap) {
  rt::Config cfg{};
  cfg.number_of_balls_sqrt = nsqrt; cfg.aspect_ratio = aspect; cfg.image_width = width; cfg.samples_per_pixel = spp;
  cfg.moving_spheres = moving != 0; cfg.max_child_rays = max_child_rays; cfg.nthreads = nthreads;
  std::ostringstream os;
  os << cfg;
  const std::string s = os.str();
  if (static_cast<int>(s.size()) + 1 > cap) return -1;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return static_cast<int>(s.size());
}
RTWH_API int rtwh_image_height(int width, double aspect) { rt::Config c{}; c.image_width = width; c.aspect_ratio = aspect; return rt::detail::image_height(c); }
RTWH_API int rtwh_effective_spp(int spp, int nthreads) { rt::Config c{}; c.samples_per_pixel = spp; c.nthreads = nthreads; try { return rt::detail::effective_spp(c); } catch (...) { return -1; } }
RTWH_API int rtwh_make_mesh(const char* base_obj, const char* out_obj, int rounds, unsigned seed, double amplitude, long long* ntris) {
  try {
    const auto base = rt::detail::load_obj(base_obj);
    const auto fine = rt::detail::subdivide_displace(base, rounds, seed, amplitude);
    rt::detail::save_obj(out_obj, fine);
    if (ntris) *ntris = static_cast<long long>(fine.faces.size());
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
// write_ppm on a caller-provided accumulation buffer (format pin against the reference's write_color)
RTWH_API int rtwh_write_ppm(const float* rgba, int width, int height, int spp, const char* path) {
  rt::Accum img; img.width = width; img.height = height; img.spp = spp;
  img.rgba.assign(rgba, rgba + static_cast<std::size_t>(width) * height * 4);
  std::ofstream out(path, std::ios::binary);
  if (!out) return 1;
  rt::write_ppm(out, img);
  return 0;
}
