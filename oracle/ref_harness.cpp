// TEST INFRASTRUCTURE ONLY (oracle/). Never linked, imported or executed by the product path.
//
// Harness around the UNMODIFIED reference sources. It #includes /root/reference/src/render.cpp
// (to reach the file-private BVHNode class, render.cpp:22-35, and ray_color, render.cpp:112-129)
// and /root/reference/src/main.cpp (for the scene builders lots_of_balls, main.cpp:23-83, and
// foo, main.cpp:85-136; its main() is renamed).  common-model.cpp and random-utils.cpp are
// compiled as their own translation units by oracle/Makefile.  Third-party headers come from
// oracle/shim/.  Output goes to oracle/_ref/ only.
//
// Exposed as a C ABI so Python tests can drive it with ctypes:
//   ref_scene_cover / ref_scene_obj / ref_scene_custom / ref_scene_free
//   ref_scene_dump          primitives + materials of a scene in insertion order
//   ref_primary_hits        deterministic primary-ray closest hit (BVH and brute force)
//   ref_render_linear       the three nested loops of render.cpp:152-163, linear-domain sums
//   ref_render_ppm          the reference's own render() (P3 text to stdout)
//   ref_seed / ref_random_double
#include <cstdint>
#include <cstring>
#include <map>

#include "render.cpp"  // NOLINT: reference TU, found via -I/root/reference/src

#define main rtweekend_reference_main
#include "main.cpp"  // NOLINT: reference TU
#undef main

namespace rtweekend::detail {
std::mt19937& gen();  // random-utils.cpp:6 (external linkage, not declared in the header)
}

namespace rd = rtweekend::detail;

extern "C" {

struct ref_prim {
  int32_t kind;      // 0 sphere, 1 moving sphere, 2 triangle
  int32_t material;  // index into the material list
  double a[3];       // sphere: centre (t=0) ; triangle: vertex a
  double b[3];       // moving sphere: centre at t=1 ; triangle: vertex b
  double c[3];       // triangle: vertex c
  double radius;
};
struct ref_mat {
  int32_t kind;  // 0 lambertian, 1 metal, 2 dielectric
  int32_t pad;
  double albedo[3];
  double fuzz;
  double ior;
};
struct ref_camera {
  double lookfrom[3], lookat[3], vup[3];
  double vfov, aspect, aperture, focus_dist /* <= 0: |from-at| */, t0, t1;
};

struct ref_scene {
  rt::Scene scene;
  ref_camera cam;
  std::vector<const rd::Primitive*> order;  // insertion order, captured before any BVH build (sorts in place)
  std::vector<const rd::Material*> mat_order;
  explicit ref_scene(rt::Scene&& s, const ref_camera& c) : scene(std::move(s)), cam(c) {
    for (auto& p : scene.primitives()) order.push_back(p.get());
    for (auto& m : scene.boutique()) mat_order.push_back(m.get());
  }
};

static rt::Camera make_camera(const ref_camera& c) {
  std::optional<double> fd;
  if (c.focus_dist > 0) fd = c.focus_dist;
  return rt::Camera{rt::point(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]),
                    rt::point(c.lookat[0], c.lookat[1], c.lookat[2]),
                    rt::vec3(c.vup[0], c.vup[1], c.vup[2]),
                    c.vfov, c.aspect, c.aperture, fd, c.t0, c.t1};
}

void ref_seed(uint32_t seed) { rd::gen().seed(seed); }
double ref_random_double(void) { return rd::random_double(); }

// Cover scene exactly as main.cpp:23-83 builds it (consumes the global mt19937 from its current state).
ref_scene* ref_scene_cover(int nsqrt, double aspect, int moving) {
  rt::Config cfg{};
  cfg.number_of_balls_sqrt = nsqrt;
  cfg.aspect_ratio = aspect;
  cfg.moving_spheres = moving != 0;
  ref_camera c{{13, 2, 3}, {0, 0, 0}, {0, 1, 0}, 20.0, aspect, 0.1, 10.0, 0, 1};  // main.cpp:25-33
  return new ref_scene(lots_of_balls(cfg), c);
}

// OBJ scene exactly as main.cpp:85-136 builds it.  Returns nullptr when the loader throws.
ref_scene* ref_scene_obj(const char* path, double aspect) {
  rt::Config cfg{};
  cfg.aspect_ratio = aspect;
  cfg.model = std::string(path);
  ref_camera c{{1, 0, -1}, {0, 0, 0}, {0, 1, 0}, 35.0, aspect, 0.01, -1.0, 0, 1};  // main.cpp:89-97
  try {
    return new ref_scene(foo(cfg), c);
  } catch (const std::exception& e) {
    std::cerr << "ref_scene_obj: " << e.what() << "\n";
    return nullptr;
  }
}

// Arbitrary scene through the reference's public Scene API (render.h:22-33).
ref_scene* ref_scene_custom(const ref_prim* prims, int nprims, const ref_mat* mats, int nmats,
                            const ref_camera* cam) {
  rt::Scene world{make_camera(*cam)};
  std::vector<rt::Material*> mp;
  for (int i = 0; i < nmats; ++i) {
    const auto& m = mats[i];
    rt::color a{m.albedo[0], m.albedo[1], m.albedo[2]};
    if (m.kind == 0) mp.push_back(&world.boutique().add<rt::Lambertian>(a));
    else if (m.kind == 1) mp.push_back(&world.boutique().add<rt::Metal>(a, m.fuzz));
    else mp.push_back(&world.boutique().add<rt::Dielectric>(m.ior, m.fuzz));
  }
  for (int i = 0; i < nprims; ++i) {
    const auto& p = prims[i];
    rt::point a{p.a[0], p.a[1], p.a[2]}, b{p.b[0], p.b[1], p.b[2]}, c{p.c[0], p.c[1], p.c[2]};
    if (p.kind == 0) world.primitives().add<rt::Sphere>(a, p.radius, *mp[p.material]);
    else if (p.kind == 1) world.primitives().add<rt::MovingSphere>(a, b, p.radius, *mp[p.material]);
    else world.primitives().add<rt::Triangle>(a, b, c, *mp[p.material]);
  }
  return new ref_scene(std::move(world), *cam);
}

void ref_scene_free(ref_scene* s) { delete s; }
int ref_scene_nprims(const ref_scene* s) { return static_cast<int>(s->order.size()); }
int ref_scene_nmats(const ref_scene* s) { return static_cast<int>(s->mat_order.size()); }
void ref_scene_camera(const ref_scene* s, ref_camera* out) { *out = s->cam; }

// Dump the scene (insertion order) so tests can compare it with the new host's scene builders.
// Triangle vertices are private in the reference (oo-primitives.h:85): triangles are reported with
// kind=2 and their bounding box corners in a/b (float-rounded, common-model.cpp:127-134).
void ref_scene_dump(const ref_scene* s, ref_prim* prims, ref_mat* mats) {
  std::map<const rd::Material*, int> mi;
  for (size_t i = 0; i < s->mat_order.size(); ++i) {
    const rd::Material* m = s->mat_order[i];
    mi[m] = static_cast<int>(i);
    ref_mat o{};
    if (auto* l = dynamic_cast<const rd::Lambertian*>(m)) { o.kind = 0; o.albedo[0] = l->albedo.x; o.albedo[1] = l->albedo.y; o.albedo[2] = l->albedo.z; }
    else if (auto* me = dynamic_cast<const rd::Metal*>(m)) { o.kind = 1; o.albedo[0] = me->albedo.x; o.albedo[1] = me->albedo.y; o.albedo[2] = me->albedo.z; o.fuzz = me->fuzz; }
    else if (auto* d = dynamic_cast<const rd::Dielectric*>(m)) { o.kind = 2; o.ior = d->ir; o.fuzz = d->fuzz; o.albedo[0] = o.albedo[1] = o.albedo[2] = 1.0; }
    mats[i] = o;
  }
  for (size_t i = 0; i < s->order.size(); ++i) {
    const rd::Primitive* p = s->order[i];
    ref_prim o{};
    o.material = mi.at(&p->material());
    if (auto* sp = dynamic_cast<const rd::Sphere*>(p)) {
      o.kind = 0; o.radius = sp->radius();
      for (int k = 0; k < 3; ++k) { o.a[k] = sp->center()[k]; o.b[k] = sp->center()[k]; }
    } else if (auto* ms = dynamic_cast<const rd::MovingSphere*>(p)) {
      o.kind = 1; o.radius = ms->radius();
      auto c0 = ms->center(0.0), c1 = ms->center(1.0);
      for (int k = 0; k < 3; ++k) { o.a[k] = c0[k]; o.b[k] = c1[k]; }
    } else {
      o.kind = 2;
      auto bb = p->bounding_box();
      for (int k = 0; k < 3; ++k) { o.a[k] = bb.min()[k]; o.b[k] = bb.max()[k]; }
    }
    prims[i] = o;
  }
}

// Deterministic primary-ray mode (BASELINE.json north_star): aperture 0 and shutter [time,time] make
// Camera::get_ray (common-model.cpp:156-167) independent of the RNG; rays go through pixel centres with the
// pixel mapping of render.cpp:152-160 (jitter replaced by 0.5).
//   id:     insertion-order primitive index of the closest hit (BVHNode::hit, render.cpp:52-71), -1 = miss
//   t:      Hit::at();  nrm: Hit::normal() (3 doubles/pixel);  front: Hit::front_facing()
// Returns the number of pixels where brute force over the primitive list (detail::hit with a shrinking upper
// bound, same accept rule as render.cpp:57-64) disagrees with the BVH on the primitive.
int ref_primary_hits(ref_scene* s, int width, int height, double time, int32_t* id, double* t, double* nrm,
                     uint8_t* front) {
  ref_camera c = s->cam;
  c.aperture = 0.0; c.t0 = time; c.t1 = time;
  rt::Camera cam = make_camera(c);
  std::map<const rd::Primitive*, int> idx;
  for (size_t i = 0; i < s->order.size(); ++i) idx[s->order[i]] = static_cast<int>(i);
  auto root = s->scene.get_root_bvh();
  int disagreements = 0;
  for (int i = 0; i < height; ++i) {
    int from_top_i = height - i - 1;
    for (int j = 0; j < width; ++j) {
      auto u = (j + 0.5) / (width - 1);
      auto v = (from_top_i + 0.5) / (height - 1);
      auto r = cam.get_ray(u, v);
      auto h = root.hit(r);
      // brute force, primitives visited in their current (BVH-sorted) store order
      std::optional<rd::Hit> bf{};
      double upper = std::numeric_limits<double>::infinity();
      for (auto& p : s->scene.primitives()) {
        auto probe = rd::hit(p, r, 0.001, upper);
        if (probe) { bf = probe; upper = probe->at(); }
      }
      size_t k = static_cast<size_t>(i) * width + j;
      if (h) {
        id[k] = idx.at(&h->what());
        t[k] = h->at();
        nrm[3 * k + 0] = h->normal().x; nrm[3 * k + 1] = h->normal().y; nrm[3 * k + 2] = h->normal().z;
        front[k] = h->front_facing();
      } else {
        id[k] = -1; t[k] = 0; nrm[3 * k + 0] = nrm[3 * k + 1] = nrm[3 * k + 2] = 0; front[k] = 0;
      }
      int bid = bf ? idx.at(&bf->what()) : -1;
      if (bid != id[k]) ++disagreements;
    }
  }
  return disagreements;
}

// Linear-domain statistics: the loops of render.cpp:152-163 on one thread, accumulating per-pixel
// sum and sum of squares of ray_color (3 doubles each) instead of writing a PPM.  Uses the global
// mt19937 from its current state (seed it with ref_seed for independent runs).
void ref_render_linear(ref_scene* s, int width, int height, int spp, int max_child_rays, double* sum,
                       double* sumsq) {
  const auto& cam = s->scene.camera();
  auto root = s->scene.get_root_bvh();
  for (int i = 0; i < height; ++i) {
    auto from_top_i = height - i - 1;
    for (int j = 0; j < width; ++j) {
      size_t k = static_cast<size_t>(i) * width + j;
      for (int q = 0; q < spp; ++q) {
        auto u = (j + rd::random_double()) / (width - 1);
        auto v = (from_top_i + rd::random_double()) / (height - 1);
        auto r = cam.get_ray(u, v);
        auto c = rd::ray_color(r, root, max_child_rays);
        for (int ch = 0; ch < 3; ++ch) { sum[3 * k + ch] += c[ch]; sumsq[3 * k + ch] += c[ch] * c[ch]; }
      }
    }
  }
}

// The reference's own render() (render.cpp:135-191): P3 text on stdout, progress on stderr.
void ref_render_ppm(ref_scene* s, int width, double aspect, int spp, int max_child_rays, int nthreads) {
  rt::Config cfg{};
  cfg.image_width = width; cfg.aspect_ratio = aspect; cfg.samples_per_pixel = spp;
  cfg.max_child_rays = max_child_rays; cfg.nthreads = nthreads;
  rt::render(s->scene, cfg);
  std::cout.flush();
}

}  // extern "C"
