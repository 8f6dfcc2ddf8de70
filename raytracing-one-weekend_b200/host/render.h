// render.h -- the drop-in boundary: rtweekend::render(const Scene&, const Config&).
// Field-for-field / signature-for-signature the reference's public surface (src/render.h:11-44): Config (8 fields,
// same defaults), Scene (camera(), primitives(), boutique()), render(), operator<<(ostream&, Config).
// Scene::get_root_bvh() is gone: the BVH is a device structure now and no caller outside render.cpp used it.
// New, additive: DeviceOptions (which GPUs, Philox seed, kernel choice) and render_accum() for callers that want the
// linear accumulation buffer instead of P3 text.
#pragma once
#include <cstdint>
#include <iosfwd>
#include <optional>
#include <string>
#include <vector>

#include "primitive-model.h"

namespace rtweekend::detail {

struct Config {
  int number_of_balls_sqrt = 11;
  double aspect_ratio = 3.0 / 2.0;
  int image_width = 200;
  int samples_per_pixel = 20;
  bool moving_spheres = true;
  int max_child_rays = 20;
  int nthreads = 4;  // kept for CLI/API compatibility: only its effect on the sample count survives
  std::optional<std::string> model = {};
};

struct DeviceOptions {
  int ngpus = 1;            // samples-per-pixel are sharded over GPUs 0..ngpus-1, one NCCL reduce at the end
  std::uint64_t seed = 0;   // Philox key
  int kernel = RTW_KERNEL_AUTO;
  int device = 0;           // first device (single-GPU renders)
  bool stats = false;       // count rays / tests (slower)
  bool split_rows = false;  // multi-GPU: interleaved row tiles + one gather instead of the sample split + one reduce
  int tile_rows = 8;        // rows per tile of the row split
  bool obj_all_shapes = false;  // -l: take the faces of every shape of the OBJ file, not only shapes[0] as the reference does
  bool binary_ppm = false;  // P6 (binary) instead of the reference's P3 text (the pixels are quantised on the device either way)
  int bvh_build = 0;        // 0 auto, 1 host (binned SAH), 2 device (linear BVH): RTW_FLAG_BVH_BUILD_HOST / _GPU of the C ABI
  std::string checkpoint;   // progressive rendering: file holding the accumulation buffer + sample cursor (see render_progressive)
  int checkpoint_every = 0; // samples per slice between checkpoint writes (0: one slice)
};
DeviceOptions& device_options();  // process-wide knobs, also settable through RTW_GPUS / RTW_SEED / RTW_KERNEL

class Scene {
  PrimitiveStore_t primitives_;
  MaterialStore_t boutique_;
  Camera cam_;

 public:
  template <typename T>
  explicit Scene(T&& camera) : cam_{std::forward<T>(camera)} {}
  [[nodiscard]] const Camera& camera() const { return cam_; }
  auto& primitives() { return primitives_; }
  auto& boutique() { return boutique_; }
  [[nodiscard]] const auto& primitives() const { return primitives_; }
  [[nodiscard]] const auto& boutique() const { return boutique_; }

  // The variant/virtual primitive list as the flat, insertion-ordered arrays of the C ABI.
  struct Flat {
    std::vector<rtw_primitive> prims;
    std::vector<rtw_material> mats;
    rtw_scene_desc desc{};
  };
  [[nodiscard]] Flat flatten() const;
};

struct Accum {
  int width = 0, height = 0, spp = 0;  // spp = effective samples per pixel
  std::vector<float> rgba;             // width*height*4: sum r, sum g, sum b, samples
  rtw_stats stats{};
};
struct Image8 {                        // the picture as the reference prints it: write_color (render.cpp:11-20) per channel
  int width = 0, height = 0, spp = 0;
  std::vector<std::uint8_t> rgb;       // width*height*3, top row first
  rtw_stats stats{};
};

int image_height(const Config& cfg);      // int(width / aspect), render.cpp:137
int effective_spp(const Config& cfg);     // spp / nthreads * nthreads, render.cpp:174,185
Accum render_accum(const Scene& world, const Config& cfg);   // linear accumulation buffer (float sums) on the host
Image8 render_rgb8(const Scene& world, const Config& cfg);   // write_color evaluated on the device from the exact sums: 3 bytes per pixel come back
void write_ppm(std::ostream& out, const Accum& img);  // render.cpp:11-20,182-186 evaluated on the host from the float sums
void write_ppm(std::ostream& out, const Image8& img); // P3 text of an already quantised image (same format)
void write_ppm_binary(std::ostream& out, const Image8& img);  // the same pixels as binary P6
Image8 quantise(const Accum& img, int device = 0);    // write_color on the device for a host-resident accumulation buffer (rtw_finalize_rgb8)
// Progressive / resumable accumulation: renders the samples [cursor, spp) in slices of `every`, adding into the
// buffer stored in `path` (created when absent) and rewriting it after every slice.  Samples are keyed by their global index, so
// a render resumed from its checkpoint produces the same bytes as the uninterrupted progressive render with the same slice size.
// Honours DeviceOptions::ngpus; the scene stays on the device(s) between slices (one flatten + BVH build per render); the
// checkpoint is bound to the scene (rtw_scene_hash), the image size, the depth and the seed.
Accum render_progressive(const Scene& world, const Config& cfg, const std::string& path, int every);
void render(const Scene& world, const Config& cfg);   // P3 text on stdout, progress on stderr

std::ostream& operator<<(std::ostream& o, const Config& c);
}  // namespace rtweekend::detail

namespace rtweekend {
using detail::Accum;
using detail::Image8;
using detail::render_rgb8;
using detail::quantise;
using detail::Config;
using detail::device_options;
using detail::DeviceOptions;
using detail::render;
using detail::render_accum;
using detail::Scene;
using detail::write_ppm;
using detail::write_ppm_binary;
using detail::render_progressive;
}  // namespace rtweekend
