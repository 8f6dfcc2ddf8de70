/* TEST INFRASTRUCTURE ONLY (oracle/) -- see rtw_oracle.h.  Plain C11, double precision, no SIMD, no FMA
 * contraction (built with -ffp-contract=off) so that the arithmetic is the reference's operation for operation.
 *
 * Citations are into /root/reference/src/.  GLM formulas (normalize, reflect, refract, dot, cross) are the
 * published generic ones, restated in oracle/shim/glm/glm.hpp, which is what oracle/_ref is compiled against.
 */
#include "rtw_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* vec3 helpers, each written in the operation order of the GLM generic implementation               */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { double x, y, z; } v3;

static inline v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, double s) { return V(a.x * s, a.y * s, a.z * s); }  /* vec * scalar */
static inline v3 sscale(double s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }  /* scalar * vec */
static inline v3 vdivs(v3 a, double s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* (x+y)+z */
static inline v3 vcross(v3 a, v3 b) {
  return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline double vlength(v3 a) { return sqrt(vdot(a, a)); }
static inline v3 vnormalize(v3 a) { return vscale(a, 1.0 / sqrt(vdot(a, a))); }
static inline v3 vreflect(v3 I, v3 N) { return vsub(I, vscale(vscale(N, vdot(N, I)), 2.0)); }
static inline v3 vrefract(v3 I, v3 N, double eta) {
  double d = vdot(N, I);
  double k = 1.0 - eta * eta * (1.0 - d * d);
  if (k >= 0.0) return vsub(sscale(eta, I), sscale(eta * d + sqrt(k), N));
  return V(0, 0, 0);
}
static inline v3 from3(const double p[3]) { return V(p[0], p[1], p[2]); }
static inline void to3(v3 v, double p[3]) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

/* ------------------------------------------------------------------------------------------------ */
/* Host RNG: std::mt19937 (default seed 5489) behind random-utils.cpp:6-9, with libstdc++'s           */
/* uniform_real_distribution<double> (generate_canonical<double,53>: two 32-bit draws, low word       */
/* first, sum/2^64, clamped below 1) and uniform_int_distribution<int>{0,1} (one draw, Lemire).       */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { uint32_t mt[624]; int idx; } mt19937;

static void mt_seed(mt19937* g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t mt_next(mt19937* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      uint32_t v = g->mt[(i + 397) % 624] ^ (y >> 1);
      if (y & 1u) v ^= 0x9908b0dfu;
      g->mt[i] = v;
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

static mt19937 g_gen;
static int g_gen_init = 0;
static mt19937* gen(void) { /* random-utils.cpp:6-9 */
  if (!g_gen_init) { mt_seed(&g_gen, 5489u); g_gen_init = 1; }
  return &g_gen;
}
void rtwo_seed(uint32_t seed) { mt_seed(&g_gen, seed); g_gen_init = 1; }

static double canonical53(mt19937* g) {
  double sum = 0.0, tmp = 1.0;
  for (int k = 0; k < 2; ++k) { sum += (double)mt_next(g) * tmp; tmp *= 4294967296.0; }
  double r = sum / tmp;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}
/* random-utils.cpp:11-13 */
double rtwo_random_double_range(double a, double b) { return canonical53(gen()) * (b - a) + a; }
double rtwo_random_double(void) { return rtwo_random_double_range(0.0, 1.0); }
/* random-utils.cpp:15-17 with the default range {0,1}: 64-bit product of one draw with 2, high word */
int rtwo_random_int01(void) { return (int)(((uint64_t)mt_next(gen()) * 2u) >> 32); }

/* random-utils.cpp:19-22 (brace-init: x, y, z drawn left to right) */
static v3 random_vec3(double mn, double mx) {
  double x = rtwo_random_double_range(mn, mx);
  double y = rtwo_random_double_range(mn, mx);
  double z = rtwo_random_double_range(mn, mx);
  return V(x, y, z);
}
/* random-utils.cpp:23-33: "random_unit_vector" is a rejection-sampled point of the unit ball restricted to
 * the positive octant (random_vec3() default range [0,1)), NOT normalised. */
static v3 random_unit_vector(void) {
  for (;;) {
    v3 v = random_vec3(0.0, 1.0);
    if (vdot(v, v) >= 1.0) continue;
    return v;
  }
}
/* random-utils.cpp:34-41.  vec3(random_double(-1,1), random_double(-1,1), 0) is a parenthesised constructor
 * call: g++ 13 evaluates its arguments right to left, so the FIRST draw lands in y (pinned by the bit-exact
 * image test against oracle/_ref). */
static int g_disk_first_draw_is_y = 1;
static v3 random_in_unit_disk(void) {
  for (;;) {
    double d0 = rtwo_random_double_range(-1.0, 1.0);
    double d1 = rtwo_random_double_range(-1.0, 1.0);
    v3 p = g_disk_first_draw_is_y ? V(d1, d0, 0.0) : V(d0, d1, 0.0);
    if (vdot(p, p) >= 1.0) continue;
    return p;
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* Scene model                                                                                        */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
  v3 origin, w, u, v, horizontal, vertical, lower_left;
  double lens_radius, t0, t1;
} camera;

struct rtwo_scene {
  rtwo_prim* prims; int nprims, cap_prims;
  rtwo_mat* mats; int nmats, cap_mats;
  rtwo_camera_params cp;
  camera cam;
};

/* common-model.cpp:136-154 */
static camera make_camera(const rtwo_camera_params* p) {
  camera c;
  v3 from = from3(p->lookfrom), at = from3(p->lookat), vup = from3(p->vup);
  c.origin = from;
  c.w = vnormalize(vsub(from, at));
  c.u = vnormalize(vcross(vup, c.w));
  c.v = vnormalize(vcross(c.w, c.u));
  c.lens_radius = p->aperture / 2;
  c.t0 = p->t0; c.t1 = p->t1;
  double viewport_height = 2.0 * tan(p->vfov * 3.141592653589793238462643383279502884 / 180 / 2);
  double viewport_width = p->aspect * viewport_height;
  double fd = p->focus_dist > 0 ? p->focus_dist : vlength(vsub(from, at));
  c.horizontal = sscale(fd * viewport_width, c.u);
  c.vertical = sscale(fd * viewport_height, c.v);
  c.lower_left = vsub(vsub(vsub(c.origin, vdivs(c.horizontal, 2.0)), vdivs(c.vertical, 2.0)), sscale(fd, c.w));
  return c;
}

typedef struct { v3 o, d; double time; } ray;

/* common-model.cpp:156-167 with the two random inputs made explicit */
static ray camera_ray(const camera* c, double s, double t, v3 disk, double when) {
  v3 rd = sscale(c->lens_radius, disk);
  v3 offset = vadd(vscale(c->u, rd.x), vscale(c->v, rd.y));
  ray r;
  r.o = vadd(c->origin, offset);
  r.d = vsub(vadd(vadd(c->lower_left, sscale(s, c->horizontal)), sscale(t, c->vertical)), r.o);
  r.time = when;
  return r;
}

static rtwo_scene* scene_new(const rtwo_camera_params* cp) {
  rtwo_scene* s = (rtwo_scene*)calloc(1, sizeof *s);
  s->cp = *cp;
  s->cam = make_camera(cp);
  return s;
}
static int add_mat(rtwo_scene* s, int kind, v3 albedo, double fuzz, double ior) {
  if (s->nmats == s->cap_mats) { s->cap_mats = s->cap_mats ? 2 * s->cap_mats : 64; s->mats = (rtwo_mat*)realloc(s->mats, sizeof(rtwo_mat) * (size_t)s->cap_mats); }
  rtwo_mat m; memset(&m, 0, sizeof m);
  m.kind = kind; to3(albedo, m.albedo);
  /* common-model.h:132-133,143-144: fuzz clamped to [0,1] in the Metal and Dielectric constructors */
  m.fuzz = fuzz < 0.0 ? 0.0 : (fuzz > 1.0 ? 1.0 : fuzz);
  m.ior = ior;
  s->mats[s->nmats] = m;
  return s->nmats++;
}
static void add_prim(rtwo_scene* s, int kind, int mat, v3 a, v3 b, v3 c, double radius) {
  if (s->nprims == s->cap_prims) { s->cap_prims = s->cap_prims ? 2 * s->cap_prims : 256; s->prims = (rtwo_prim*)realloc(s->prims, sizeof(rtwo_prim) * (size_t)s->cap_prims); }
  rtwo_prim p; memset(&p, 0, sizeof p);
  p.kind = kind; p.material = mat; to3(a, p.a); to3(b, p.b); to3(c, p.c); p.radius = radius;
  s->prims[s->nprims++] = p;
}

/* main.cpp:23-83.  `rt::point center(a + 0.9*random_double(), 0.2, b + 0.9*random_double())` is again a
 * parenthesised constructor call evaluated right to left by g++ 13: the draw after choose_mat goes to z. */
static int g_center_first_draw_is_z = 1;
rtwo_scene* rtwo_scene_cover(int nsqrt, double aspect, int moving) {
  rtwo_camera_params cp = {{13, 2, 3}, {0, 0, 0}, {0, 1, 0}, 20.0, aspect, 0.1, 10.0, 0, 1};
  rtwo_scene* s = scene_new(&cp);
  int ground = add_mat(s, RTWO_LAMBERTIAN, V(0.5, 0.5, 0.5), 0, 0);
  add_prim(s, RTWO_SPHERE, ground, V(0, -1000, 0), V(0, -1000, 0), V(0, 0, 0), 1000.0);
  for (int a = -nsqrt; a < nsqrt; a++) {
    for (int b = -nsqrt; b < nsqrt; b++) {
      double choose_mat = rtwo_random_double();
      double d0 = rtwo_random_double();
      double d1 = rtwo_random_double();
      v3 center = g_center_first_draw_is_z ? V(a + 0.9 * d1, 0.2, b + 0.9 * d0) : V(a + 0.9 * d0, 0.2, b + 0.9 * d1);
      if (vlength(vsub(center, V(4, 0.2, 0))) > 0.9) {
        if (choose_mat < 0.8) {
          /* operands of `random_vec3() * random_vec3()`: the product is commutative per component */
          v3 r1 = random_vec3(0, 1), r2 = random_vec3(0, 1);
          int m = add_mat(s, RTWO_LAMBERTIAN, vmul(r1, r2), 0, 0);
          if (moving) {
            v3 center2 = vadd(center, V(0, rtwo_random_double_range(0, .5), 0));
            add_prim(s, RTWO_MOVING_SPHERE, m, center, center2, V(0, 0, 0), 0.2);
          } else {
            add_prim(s, RTWO_SPHERE, m, center, center, V(0, 0, 0), 0.2);
          }
        } else if (choose_mat < 0.95) {
          v3 albedo = random_vec3(0.5, 1);
          double fuzz = rtwo_random_double_range(0, 0.5);
          int m = add_mat(s, RTWO_METAL, albedo, fuzz, 0);
          add_prim(s, RTWO_SPHERE, m, center, center, V(0, 0, 0), 0.2);
        } else {
          int m = add_mat(s, RTWO_DIELECTRIC, V(1, 1, 1), 0, 1.5);
          add_prim(s, RTWO_SPHERE, m, center, center, V(0, 0, 0), 0.2);
        }
      }
    }
  }
  int glass = add_mat(s, RTWO_DIELECTRIC, V(1, 1, 1), 0, 1.5);
  int reddish = add_mat(s, RTWO_LAMBERTIAN, V(0.4, 0.2, 0.1), 0, 0);
  int reddish_metal = add_mat(s, RTWO_METAL, V(0.7, 0.6, 0.5), 0, 0);
  add_prim(s, RTWO_SPHERE, glass, V(0, 1, 0), V(0, 1, 0), V(0, 0, 0), 1.0);
  add_prim(s, RTWO_SPHERE, reddish, V(-4, 1, 0), V(-4, 1, 0), V(0, 0, 0), 1.0);
  add_prim(s, RTWO_SPHERE, reddish_metal, V(4, 1, 0), V(4, 1, 0), V(0, 0, 0), 1.0);
  return s;
}

/* tinyobjloader 1.0.6 real parser restated (see oracle/shim/tiny_obj_loader.h for the description) */
static int parse_real(const char* s, const char* e, double* out) {
  if (s >= e) return 0;
  double m = 0.0; int ex = 0, neg = 0, exneg = 0, nread = 0;
  const char* p = s;
  if (*p == '+' || *p == '-') { neg = (*p == '-'); ++p; }
  else if (!(*p >= '0' && *p <= '9')) return 0;
  while (p != e && *p >= '0' && *p <= '9') { m *= 10; m += (int)(*p - '0'); ++p; ++nread; }
  if (nread == 0) return 0;
  if (p != e) {
    int go_exp = 0;
    if (*p == '.') {
      static const double lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
      int k = 1; ++p;
      while (p != e && *p >= '0' && *p <= '9') { m += (int)(*p - '0') * (k < 8 ? lut[k] : pow(10.0, -k)); ++k; ++p; }
      go_exp = (p != e);
    } else if (*p == 'e' || *p == 'E') go_exp = 1;
    if (go_exp && (*p == 'e' || *p == 'E')) {
      int n = 0; ++p;
      if (p != e && (*p == '+' || *p == '-')) { exneg = (*p == '-'); ++p; }
      else if (p == e || !(*p >= '0' && *p <= '9')) return 0;
      while (p != e && *p >= '0' && *p <= '9') { ex *= 10; ex += (int)(*p - '0'); ++p; ++n; }
      if (exneg) ex = -ex;
      if (n == 0) return 0;
    }
  }
  *out = (neg ? -1 : 1) * (ex ? ldexp(m * pow(5.0, ex), ex) : m);
  return 1;
}

/* main.cpp:85-136: one grey Lambertian, camera (1,0,-1)->(0,0,0) fov 35 aperture .01 focus |from-at|, every
 * face of the FIRST shape as a Triangle; a non-triangular face is an error (tinyobj triangulates by default,
 * so polygons arrive as fans).  random_int() at main.cpp:86 consumes one draw. */
rtwo_scene* rtwo_scene_obj(const char* path, double aspect) {
  (void)rtwo_random_int01();
  rtwo_camera_params cp = {{1, 0, -1}, {0, 0, 0}, {0, 1, 0}, 35.0, aspect, 0.01, -1.0, 0, 1};
  FILE* f = fopen(path, "r");
  if (!f) return NULL;
  rtwo_scene* s = scene_new(&cp);
  int grey = add_mat(s, RTWO_LAMBERTIAN, V(0.5, 0.5, 0.5), 0, 0);
  double* verts = NULL; size_t nv = 0, capv = 0;
  char line[4096];
  int shape_has_faces = 0, shape_closed = 0;
  while (fgets(line, sizeof line, f)) {
    char* p = line + strspn(line, " \t");
    if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
      p += 2;
      for (int k = 0; k < 3; ++k) {
        p += strspn(p, " \t");
        char* e = p + strcspn(p, " \t\r\n");
        double val = 0.0; parse_real(p, e, &val); p = e;
        if (nv == capv) { capv = capv ? 2 * capv : 4096; verts = (double*)realloc(verts, capv * sizeof(double)); }
        verts[nv++] = val;
      }
    } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
      if (shape_closed) continue; /* only shapes[0] is read (main.cpp:117) */
      p += 2;
      int idx[64], n = 0;
      for (;;) {
        p += strspn(p, " \t\r\n");
        if (!*p) break;
        int i = atoi(p);
        int nvert = (int)(nv / 3);
        if (n < 64) idx[n++] = i > 0 ? i - 1 : nvert + i;
        p += strcspn(p, " \t\r\n");
      }
      for (int k = 2; k < n; ++k) {
        const double *A = verts + 3 * idx[0], *B = verts + 3 * idx[k - 1], *C = verts + 3 * idx[k];
        add_prim(s, RTWO_TRIANGLE, grey, from3(A), from3(B), from3(C), 0.0);
      }
      shape_has_faces = 1;
    } else if ((p[0] == 'g' || p[0] == 'o') && (p[1] == ' ' || p[1] == '\t' || p[1] == '\n' || p[1] == '\r')) {
      if (shape_has_faces) shape_closed = 1;
    }
  }
  fclose(f);
  free(verts);
  return s;
}

rtwo_scene* rtwo_scene_custom(const rtwo_prim* prims, int nprims, const rtwo_mat* mats, int nmats,
                              const rtwo_camera_params* cam) {
  rtwo_scene* s = scene_new(cam);
  for (int i = 0; i < nmats; ++i) add_mat(s, mats[i].kind, from3(mats[i].albedo), mats[i].fuzz, mats[i].ior);
  for (int i = 0; i < nprims; ++i)
    add_prim(s, prims[i].kind, prims[i].material, from3(prims[i].a), from3(prims[i].b), from3(prims[i].c), prims[i].radius);
  return s;
}
void rtwo_scene_free(rtwo_scene* s) { if (s) { free(s->prims); free(s->mats); free(s); } }
int rtwo_scene_nprims(const rtwo_scene* s) { return s->nprims; }
int rtwo_scene_nmats(const rtwo_scene* s) { return s->nmats; }
void rtwo_scene_dump(const rtwo_scene* s, rtwo_prim* prims, rtwo_mat* mats) {
  memcpy(prims, s->prims, sizeof(rtwo_prim) * (size_t)s->nprims);
  memcpy(mats, s->mats, sizeof(rtwo_mat) * (size_t)s->nmats);
}
void rtwo_scene_camera(const rtwo_scene* s, rtwo_camera_params* out) { *out = s->cp; }
void rtwo_scene_camera_derived(const rtwo_scene* s, double out[21]) {
  const camera* c = &s->cam;
  to3(c->origin, out); to3(c->lower_left, out + 3); to3(c->horizontal, out + 6); to3(c->vertical, out + 9);
  to3(c->u, out + 12); to3(c->v, out + 15);
  out[18] = c->lens_radius; out[19] = c->t0; out[20] = c->t1;
}

/* ------------------------------------------------------------------------------------------------ */
/* Intersection                                                                                       */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int prim; v3 p; double t; v3 n; int front; } hitrec;

/* common-model.cpp:64-91 */
static int sphere_hit(const ray* r, double tmin, double tmax, v3 center, double radius, hitrec* h) {
  v3 oc = vsub(r->o, center);
  double a = vdot(r->d, r->d);
  double hh = vdot(oc, r->d);
  double c = vdot(oc, oc) - radius * radius;
  double disc = hh * hh - a * c;
  if (disc < 0.0) return 0;
  double root = (-hh - sqrt(disc)) / a;
  if (root < tmin || root > tmax) {
    root = (-hh + sqrt(disc)) / a;
    if (root < tmin || root > tmax) return 0;
  }
  v3 p = vadd(r->o, vscale(r->d, root));
  v3 n = vnormalize(vsub(p, center));
  int front = (vdot(r->d, n) < 0) ^ (radius < 0);
  h->p = p; h->t = root; h->n = front ? n : vneg(n); h->front = front;
  return 1;
}
/* oo-primitives.h:64-66 with t0_=0, t1_=1 (oo-primitives.h:51-52) */
static v3 moving_center(const rtwo_prim* p, double time) {
  v3 c0 = from3(p->a), c1 = from3(p->b);
  return vadd(c0, sscale((time - 0.0) / (1.0 - 0.0), vsub(c1, c0)));
}
/* common-model.cpp:103-125: un-normalised geometric normal, front_facing always true, det >= 1e-6 culls
 * back faces */
static int triangle_hit(const ray* r, double tmin, double tmax, v3 A, v3 B, v3 C, hitrec* h) {
  v3 e1 = vsub(B, A), e2 = vsub(C, A);
  v3 n = vcross(e1, e2);
  double det = -vdot(r->d, n);
  double invdet = 1.0 / det;
  v3 ao = vsub(r->o, A);
  v3 dao = vcross(ao, r->d);
  double u = vdot(e2, dao) * invdet;
  double v = -vdot(e1, dao) * invdet;
  double t = vdot(ao, n) * invdet;
  if (det >= 1e-6 && t >= tmin && t <= tmax && u >= 0.0 && v >= 0.0 && (u + v) <= 1.0) {
    h->p = vadd(r->o, vscale(r->d, t)); h->t = t; h->n = n; h->front = 1;
    return 1;
  }
  return 0;
}
static int prim_hit(const rtwo_prim* p, const ray* r, double tmin, double tmax, hitrec* h) {
  switch (p->kind) {
    case RTWO_SPHERE: return sphere_hit(r, tmin, tmax, from3(p->a), p->radius, h);          /* common-model.cpp:93-96 */
    case RTWO_MOVING_SPHERE: return sphere_hit(r, tmin, tmax, moving_center(p, r->time), p->radius, h); /* :98-101 */
    default: return triangle_hit(r, tmin, tmax, from3(p->a), from3(p->b), from3(p->c), h);
  }
}
/* Closest hit.  The reference walks a median-split BVH (render.cpp:52-110); its leaf loop and child order
 * implement "ordered closest hit with a shrinking upper bound, later primitive wins exact ties"
 * (render.cpp:55-70).  Any exact closest-hit search returns the same primitive except on exact-t ties, so
 * the oracle scans the list in insertion order with the same accept rule. */
static int closest_hit(const rtwo_scene* s, const ray* r, hitrec* out) {
  double upper = INFINITY;
  int found = 0;
  hitrec h;
  for (int i = 0; i < s->nprims; ++i) {
    if (prim_hit(&s->prims[i], r, 0.001, upper, &h)) { h.prim = i; *out = h; upper = h.t; found = 1; }
  }
  return found;
}

/* ------------------------------------------------------------------------------------------------ */
/* Materials                                                                                          */
/* ------------------------------------------------------------------------------------------------ */
/* common-model.cpp:33-38 */
static double reflectance(double cosine, double ref_idx) {
  double r0 = (1 - ref_idx) / (1 + ref_idx);
  r0 = r0 * r0;
  return r0 + (1 - r0) * pow((1 - cosine), 5);
}
/* common-model.cpp:13-62 with the random inputs passed in.  `need_coin` tells the caller whether the
 * Schlick coin is consumed (short-circuit at common-model.cpp:53-54): call with coin < 0 first. */
static int dielectric_needs_coin(const rtwo_mat* m, v3 d_in, v3 n, int front) {
  v3 unit = vnormalize(d_in);
  double cos_theta = vdot(vneg(unit), n);
  double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
  double ratio = front ? (1.0 / m->ior) : m->ior;
  return !(ratio * sin_theta > 1.0);
}
static int scatter(const rtwo_mat* m, v3 d_in, v3 n, int front, v3 ball, double coin, v3* d_out, v3* att) {
  if (m->kind == RTWO_LAMBERTIAN) { /* common-model.cpp:13-22 */
    if (fabs(n.x - ball.x) < 1e-8 && fabs(n.y - ball.y) < 1e-8 && fabs(n.z - ball.z) < 1e-8) return 0;
    *d_out = vadd(n, ball);
    *att = from3(m->albedo);
    return 1;
  }
  if (m->kind == RTWO_METAL) { /* common-model.cpp:24-31: always scatters */
    v3 reflected = vreflect(d_in, n);
    *d_out = vadd(reflected, sscale(m->fuzz, ball));
    *att = from3(m->albedo);
    return 1;
  }
  /* common-model.cpp:40-62 */
  v3 unit = vnormalize(d_in);
  double cos_theta = vdot(vneg(unit), n);
  double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
  double ratio = front ? (1.0 / m->ior) : m->ior;
  int cannot_refract = ratio * sin_theta > 1.0;
  v3 dir;
  if (cannot_refract || reflectance(cos_theta, ratio) > coin) dir = vreflect(unit, n);
  else dir = vrefract(unit, n, ratio);
  *d_out = vadd(dir, sscale(m->fuzz, ball));
  *att = V(1.0, 1.0, 1.0);
  return 1;
}

/* render.cpp:125-128 */
static v3 sky(v3 d) {
  v3 unit = vnormalize(d);
  double t = 0.5 * (unit.y + +1.0);
  return vadd(sscale(1.0 - t, V(1.0, 1.0, 1.0)), sscale(t, V(0.5, 0.7, 1.0)));
}

/* Random-number source for one path: either the global mt19937 with the reference's rejection loops, or the
 * Philox contract of the new renderer. */
typedef struct {
  int philox;
  uint32_t key[2], pixel, sample;
  int scatter_index;
} rsrc;

void rtwo_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline double u01(uint32_t x) { return (double)(x >> 8) * (1.0 / 16777216.0); }
static void philox_dim(const rsrc* rs, uint32_t dim, double xi[4]) {
  uint32_t ctr[4] = {rs->pixel, rs->sample, dim, 0u}, o[4];
  rtwo_philox4x32_10(ctr, rs->key, o);
  for (int k = 0; k < 4; ++k) xi[k] = u01(o[k]);
}
/* Direct inversion onto the distribution of random-utils.cpp:23-33 (uniform in the positive octant of the unit
 * ball): z uniform in [0,1), azimuth uniform in [0,pi/2), radius = cbrt(xi). */
static v3 octant_ball(double xz, double xphi, double xrho) {
  double rho = cbrt(xrho);
  double sn = sqrt(1.0 - xz * xz);
  double phi = 1.5707963267948966 * xphi;
  return V(rho * sn * cos(phi), rho * sn * sin(phi), rho * xz);
}

/* render.cpp:112-129 (recursive, so that attenuation products associate exactly as in the reference).
 * max_depth is size_t there: `max_depth <= 0` means == 0. */
static v3 ray_color(const rtwo_scene* s, const ray* r, int max_depth, rsrc* rs, uint64_t* nrays) {
  hitrec h;
  ++*nrays;
  if (closest_hit(s, r, &h)) {
    if (max_depth <= 0) return V(0, 0, 0);
    const rtwo_mat* m = &s->mats[s->prims[h.prim].material];
    v3 ball; double coin = 2.0;
    if (rs->philox) {
      double xi[4];
      philox_dim(rs, 2u + (uint32_t)rs->scatter_index, xi);
      rs->scatter_index++;
      ball = octant_ball(xi[0], xi[1], xi[2]);
      coin = xi[3];
    } else {
      /* draw order inside each scatter(): Lambertian/Metal draw only the ball vector; Dielectric draws the coin
       * (only when refraction is possible) and then the ball vector (common-model.cpp:53-60) */
      if (m->kind == RTWO_DIELECTRIC && dielectric_needs_coin(m, r->d, h.n, h.front)) coin = rtwo_random_double();
      ball = random_unit_vector();
    }
    v3 d_out, att;
    if (scatter(m, r->d, h.n, h.front, ball, coin, &d_out, &att)) {
      ray child; child.o = h.p; child.d = d_out; child.time = r->time;
      return vmul(att, ray_color(s, &child, max_depth - 1, rs, nrays));
    }
    return V(0, 0, 0);
  }
  return sky(r->d);
}

/* ------------------------------------------------------------------------------------------------ */
/* Public drivers                                                                                     */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
  const rtwo_scene* s; const camera* cam; int width, height, row0, row1; double time;
  int32_t* id; double* t; double* nrm; uint8_t* front;
} primary_job;

static void* primary_worker(void* arg) {
  primary_job* jb = (primary_job*)arg;
  for (int i = jb->row0; i < jb->row1; ++i) {
    int from_top_i = jb->height - i - 1;
    for (int j = 0; j < jb->width; ++j) {
      double u = (j + 0.5) / (jb->width - 1);
      double v = (from_top_i + 0.5) / (jb->height - 1);
      ray r = camera_ray(jb->cam, u, v, V(0, 0, 0), jb->time);
      size_t k = (size_t)i * (size_t)jb->width + (size_t)j;
      hitrec h;
      if (closest_hit(jb->s, &r, &h)) {
        jb->id[k] = h.prim; jb->t[k] = h.t; jb->nrm[3 * k] = h.n.x; jb->nrm[3 * k + 1] = h.n.y; jb->nrm[3 * k + 2] = h.n.z;
        jb->front[k] = (uint8_t)h.front;
      } else {
        jb->id[k] = -1; jb->t[k] = 0; jb->nrm[3 * k] = jb->nrm[3 * k + 1] = jb->nrm[3 * k + 2] = 0; jb->front[k] = 0;
      }
    }
  }
  return NULL;
}

/* rows split over nthreads (the rays are independent; brute force over a million triangles needs the cores) */
void rtwo_primary_hits_mt(const rtwo_scene* s, int width, int height, double time, int nthreads, int32_t* id, double* t,
                          double* nrm, uint8_t* front) {
  rtwo_camera_params cp = s->cp;
  cp.aperture = 0.0; cp.t0 = time; cp.t1 = time;
  camera cam = make_camera(&cp);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (nthreads > height) nthreads = height > 0 ? height : 1;
  primary_job jobs[256]; pthread_t th[256];
  for (int k = 0; k < nthreads; ++k) {
    primary_job jb = {s, &cam, width, height, (int)((long long)height * k / nthreads), (int)((long long)height * (k + 1) / nthreads), time,
                      id, t, nrm, front};
    jobs[k] = jb;
    pthread_create(&th[k], NULL, primary_worker, &jobs[k]);
  }
  for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
}

void rtwo_primary_hits(const rtwo_scene* s, int width, int height, double time, int32_t* id, double* t,
                       double* nrm, uint8_t* front) {
  rtwo_primary_hits_mt(s, width, height, time, 1, id, t, nrm, front);
}

/* render.cpp:152-163, one thread */
void rtwo_render_linear(const rtwo_scene* s, int width, int height, int spp, int max_child_rays, double* sum,
                        double* sumsq, uint64_t* nrays) {
  rsrc rs; memset(&rs, 0, sizeof rs);
  uint64_t rays = 0;
  for (int i = 0; i < height; ++i) {
    int from_top_i = height - i - 1;
    for (int j = 0; j < width; ++j) {
      size_t k = (size_t)i * (size_t)width + (size_t)j;
      v3 pixel = V(0, 0, 0), pixel2 = V(0, 0, 0);
      for (int q = 0; q < spp; ++q) {
        double u = (j + rtwo_random_double()) / (width - 1);
        double v = (from_top_i + rtwo_random_double()) / (height - 1);
        v3 disk = random_in_unit_disk();
        double when = rtwo_random_double_range(s->cam.t0, s->cam.t1);
        ray r = camera_ray(&s->cam, u, v, disk, when);
        v3 c = ray_color(s, &r, max_child_rays, &rs, &rays);
        pixel = vadd(pixel, c);
        pixel2 = vadd(pixel2, vmul(c, c));
      }
      sum[3 * k] += pixel.x; sum[3 * k + 1] += pixel.y; sum[3 * k + 2] += pixel.z;
      if (sumsq) { sumsq[3 * k] += pixel2.x; sumsq[3 * k + 1] += pixel2.y; sumsq[3 * k + 2] += pixel2.z; }
    }
  }
  if (nrays) *nrays += rays;
}

/* render.cpp:11-20 */
void rtwo_quantize(const double* sum, int npixels, int spp, uint8_t* rgb) {
  for (int k = 0; k < 3 * npixels; ++k) {
    double c = sqrt(sum[k] / (double)spp);
    double cl = c < 0.0 ? 0.0 : (c > 0.999 ? 0.999 : c); /* std::clamp: NaN compares false twice -> NaN -> int UB; not reached */
    rgb[k] = (uint8_t)(int)(256 * cl);
  }
}

typedef struct {
  const rtwo_scene* s; int width, height, s0, s1, depth; uint64_t seed; int row0, row1;
  double *sum, *sumsq; uint64_t nrays;
} philox_job;

static void* philox_worker(void* arg) {
  philox_job* jb = (philox_job*)arg;
  const rtwo_scene* s = jb->s;
  for (int i = jb->row0; i < jb->row1; ++i) {
    int from_top_i = jb->height - i - 1;
    for (int j = 0; j < jb->width; ++j) {
      size_t k = (size_t)i * (size_t)jb->width + (size_t)j;
      for (int q = jb->s0; q < jb->s1; ++q) {
        rsrc rs; rs.philox = 1; rs.key[0] = (uint32_t)jb->seed; rs.key[1] = (uint32_t)(jb->seed >> 32);
        rs.pixel = (uint32_t)k; rs.sample = (uint32_t)q; rs.scatter_index = 0;
        /* dimension 0 carries five 24-bit uniforms: the top 24 bits of the four words, and a fifth one assembled from the
         * low bytes of words 0..2 (shutter time); dimension 1 is unused */
        double x0[4], x_time;
        {
          uint32_t ctr[4] = {rs.pixel, rs.sample, 0u, 0u}, o[4];
          rtwo_philox4x32_10(ctr, rs.key, o);
          for (int m = 0; m < 4; ++m) x0[m] = u01(o[m]);
          x_time = (double)(((o[0] & 0xffu) << 16) | ((o[1] & 0xffu) << 8) | (o[2] & 0xffu)) * (1.0 / 16777216.0);
        }
        double u = (j + x0[0]) / (jb->width - 1);
        double v = (from_top_i + x0[1]) / (jb->height - 1);
        double rr = sqrt(x0[2]), ph = 6.283185307179586 * x0[3];
        v3 disk = V(rr * cos(ph), rr * sin(ph), 0.0);
        double when = s->cam.t0 + x_time * (s->cam.t1 - s->cam.t0);
        ray r = camera_ray(&s->cam, u, v, disk, when);
        v3 c = ray_color(s, &r, jb->depth, &rs, &jb->nrays);
        jb->sum[3 * k] += c.x; jb->sum[3 * k + 1] += c.y; jb->sum[3 * k + 2] += c.z;
        if (jb->sumsq) { jb->sumsq[3 * k] += c.x * c.x; jb->sumsq[3 * k + 1] += c.y * c.y; jb->sumsq[3 * k + 2] += c.z * c.z; }
      }
    }
  }
  return NULL;
}

void rtwo_render_philox(const rtwo_scene* s, int width, int height, int sample_begin, int sample_end,
                        int max_child_rays, uint64_t seed, int nthreads, double* sum, double* sumsq,
                        uint64_t* nrays) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (nthreads > height) nthreads = height > 0 ? height : 1;
  philox_job jobs[256]; pthread_t th[256];
  for (int t = 0; t < nthreads; ++t) {
    philox_job jb = {s, width, height, sample_begin, sample_end, max_child_rays, seed,
                     (int)((long long)height * t / nthreads), (int)((long long)height * (t + 1) / nthreads), sum, sumsq, 0};
    jobs[t] = jb;
    pthread_create(&th[t], NULL, philox_worker, &jobs[t]);
  }
  uint64_t rays = 0;
  for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); rays += jobs[t].nrays; }
  if (nrays) *nrays += rays;
}

/* ------------------------------------------------------------------------------------------------ */
/* Unit-level entry points                                                                            */
/* ------------------------------------------------------------------------------------------------ */
int rtwo_hit_sphere(const double org[3], const double dir[3], double tmin, double tmax, const double center[3],
                    double radius, double* t, double point[3], double normal[3], int* front) {
  ray r = {from3(org), from3(dir), 0.0}; hitrec h;
  if (!sphere_hit(&r, tmin, tmax, from3(center), radius, &h)) return 0;
  *t = h.t; to3(h.p, point); to3(h.n, normal); *front = h.front;
  return 1;
}
int rtwo_hit_triangle(const double org[3], const double dir[3], double tmin, double tmax, const double a[3],
                      const double b[3], const double c[3], double* t, double point[3], double normal[3]) {
  ray r = {from3(org), from3(dir), 0.0}; hitrec h;
  if (!triangle_hit(&r, tmin, tmax, from3(a), from3(b), from3(c), &h)) return 0;
  *t = h.t; to3(h.p, point); to3(h.n, normal);
  return 1;
}
int rtwo_scatter(const rtwo_mat* m, const double dir_in[3], const double normal[3], int front,
                 const double ball[3], double coin, double dir_out[3], double attenuation[3]) {
  v3 d, a;
  rtwo_mat mm = *m;
  mm.fuzz = mm.fuzz < 0.0 ? 0.0 : (mm.fuzz > 1.0 ? 1.0 : mm.fuzz);
  if (!scatter(&mm, from3(dir_in), from3(normal), front, from3(ball), coin, &d, &a)) return 0;
  to3(d, dir_out); to3(a, attenuation);
  return 1;
}
void rtwo_sky(const double dir[3], double rgb[3]) { to3(sky(from3(dir)), rgb); }
