/* TEST INFRASTRUCTURE ONLY (oracle/).  Never linked, imported or executed by the product path;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Plain-C, double-precision, CPU restatement of the reference's per-pixel / per-sample path-tracing
 * loop (joaotavora/raytracing-one-weekend, src/render.cpp + common-model.cpp + random-utils.cpp +
 * the scene builders of main.cpp).  Every function in rtw_oracle.c cites the reference file:line it
 * restates.  PARITY PIN: tests/test_oracle_pin.py checks this port against the unmodified reference
 * compiled into oracle/_ref/ (bit-identical P3 image at -t 1, identical scenes, identical primary
 * hits) and against the golden fixtures generated from it under tests/golden/.
 */
#ifndef RTW_ORACLE_H
#define RTW_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { RTWO_SPHERE = 0, RTWO_MOVING_SPHERE = 1, RTWO_TRIANGLE = 2 };
enum { RTWO_LAMBERTIAN = 0, RTWO_METAL = 1, RTWO_DIELECTRIC = 2 };

typedef struct rtwo_prim {
  int32_t kind;     /* RTWO_SPHERE / RTWO_MOVING_SPHERE / RTWO_TRIANGLE */
  int32_t material; /* index into the material list */
  double a[3];      /* sphere: centre at t=0; triangle: vertex a */
  double b[3];      /* moving sphere: centre at t=1; triangle: vertex b */
  double c[3];      /* triangle: vertex c */
  double radius;
} rtwo_prim;

typedef struct rtwo_mat {
  int32_t kind; /* RTWO_LAMBERTIAN / RTWO_METAL / RTWO_DIELECTRIC */
  int32_t pad;
  double albedo[3];
  double fuzz;
  double ior;
} rtwo_mat;

typedef struct rtwo_camera_params {
  double lookfrom[3], lookat[3], vup[3];
  double vfov, aspect, aperture, focus_dist /* <= 0: |lookfrom - lookat| */, t0, t1;
} rtwo_camera_params;

typedef struct rtwo_scene rtwo_scene;

/* --- host RNG of the reference: one global std::mt19937 + libstdc++ distributions --- */
void rtwo_seed(uint32_t seed);
double rtwo_random_double(void);
double rtwo_random_double_range(double a, double b);
int rtwo_random_int01(void);

/* --- scenes --- */
rtwo_scene* rtwo_scene_cover(int nsqrt, double aspect, int moving);
rtwo_scene* rtwo_scene_obj(const char* path, double aspect);
rtwo_scene* rtwo_scene_custom(const rtwo_prim* prims, int nprims, const rtwo_mat* mats, int nmats,
                              const rtwo_camera_params* cam);
void rtwo_scene_free(rtwo_scene* s);
int rtwo_scene_nprims(const rtwo_scene* s);
int rtwo_scene_nmats(const rtwo_scene* s);
void rtwo_scene_dump(const rtwo_scene* s, rtwo_prim* prims, rtwo_mat* mats);
void rtwo_scene_camera(const rtwo_scene* s, rtwo_camera_params* out);
/* derived camera block: origin, lower_left, horizontal, vertical, u, v (3 doubles each), lens_radius, t0, t1 */
void rtwo_scene_camera_derived(const rtwo_scene* s, double out[21]);

/* --- deterministic primary-ray mode (aperture 0, shutter [time,time], pixel centres) --- */
void rtwo_primary_hits_mt(const rtwo_scene* s, int width, int height, double time, int nthreads, int32_t* id, double* t,
                          double* nrm, uint8_t* front);
void rtwo_primary_hits(const rtwo_scene* s, int width, int height, double time, int32_t* id, double* t,
                       double* nrm, uint8_t* front);

/* --- renders --- */
/* reference sampling order on the global mt19937 (render.cpp:152-163), one thread; sums are += */
void rtwo_render_linear(const rtwo_scene* s, int width, int height, int spp, int max_child_rays, double* sum,
                        double* sumsq, uint64_t* nrays);
/* write_color (render.cpp:11-20) applied to linear sums: rgb8 = int(256*clamp(sqrt(sum/spp),0,.999)) */
void rtwo_quantize(const double* sum, int npixels, int spp, uint8_t* rgb);
/* the new renderer's sampling contract (Philox4x32-10 keyed on pixel, global sample index, dimension; direct
 * inversion instead of rejection loops) evaluated in double: same random numbers as the CUDA kernels */
void rtwo_render_philox(const rtwo_scene* s, int width, int height, int sample_begin, int sample_end,
                        int max_child_rays, uint64_t seed, int nthreads, double* sum, double* sumsq,
                        uint64_t* nrays);
void rtwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* --- unit-level entry points for closed-form tests (each returns 1 on hit / scatter) --- */
int rtwo_hit_sphere(const double org[3], const double dir[3], double tmin, double tmax, const double center[3],
                    double radius, double* t, double point[3], double normal[3], int* front);
int rtwo_hit_triangle(const double org[3], const double dir[3], double tmin, double tmax, const double a[3],
                      const double b[3], const double c[3], double* t, double point[3], double normal[3]);
/* scatter with explicit random inputs: ball[3] plays random_unit_vector(), coin plays random_double() */
int rtwo_scatter(const rtwo_mat* m, const double dir_in[3], const double normal[3], int front,
                 const double ball[3], double coin, double dir_out[3], double attenuation[3]);
void rtwo_sky(const double dir[3], double rgb[3]);

#ifdef __cplusplus
}
#endif
#endif
