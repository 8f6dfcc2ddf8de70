// render.cpp -- host side of the drop-in render(): flatten the scene, cross the C ABI, format the PPM.
// Replaces the body of the reference's render() (render.cpp:135-191).  The three nested loops, ray_color, the BVH and
// the thread fan-out/sum all run behind rtw_render / rtw_render_multi_gpu; what stays on the host is exactly what the
// reference does outside its hot loop: image height from the aspect ratio (:137), the effective sample count (:174,185),
// gamma + quantisation + P3 text (:11-20,182-186) and the timing line (:188-190).
#include "render.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace rtweekend::detail {

DeviceOptions& device_options() {
  static DeviceOptions opts = [] {
    DeviceOptions o;
    if (const char* e = std::getenv("RTW_GPUS")) o.ngpus = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RTW_SEED")) o.seed = std::strtoull(e, nullptr, 0);
    if (const char* e = std::getenv("RTW_KERNEL")) o.kernel = std::atoi(e);
    if (const char* e = std::getenv("RTW_DEVICE")) o.device = std::atoi(e);
    if (const char* e = std::getenv("RTW_STATS")) o.stats = std::atoi(e) != 0;
    if (const char* e = std::getenv("RTW_SPLIT")) o.split_rows = std::string(e) == "rows";
    if (const char* e = std::getenv("RTW_TILE_ROWS")) o.tile_rows = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RTW_BVH_BUILD")) o.bvh_build = std::string(e) == "host" ? 1 : (std::string(e) == "gpu" ? 2 : 0);
    return o;
  }();
  return opts;
}

Scene::Flat Scene::flatten() const {
  Flat f;
  std::unordered_map<const Material*, int> index;
  f.mats.reserve(boutique_.size());
  for (const auto& m : boutique_) {
    index.emplace(m.get(), static_cast<int>(f.mats.size()));
    f.mats.push_back(m->flat());
  }
  f.prims.reserve(primitives_.size());
  for (const auto& p : primitives_) {  // detail::flat / detail::material_of: virtual call or std::visit (primitive-model.h)
    rtw_primitive q = flat(p);
    const Material* m = &material_of(p);
    auto it = index.find(m);
    if (it == index.end()) {  // material owned by someone else: append it
      it = index.emplace(m, static_cast<int>(f.mats.size())).first;
      f.mats.push_back(m->flat());
    }
    q.material = it->second;
    f.prims.push_back(q);
  }
  f.desc.prims = f.prims.data();
  f.desc.nprims = static_cast<int64_t>(f.prims.size());
  f.desc.mats = f.mats.data();
  f.desc.nmats = static_cast<int64_t>(f.mats.size());
  f.desc.camera = cam_.block();
  return f;
}

int image_height(const Config& cfg) { return static_cast<int>(cfg.image_width / cfg.aspect_ratio); }

int effective_spp(const Config& cfg) {
  if (cfg.nthreads <= 0) throw std::invalid_argument("nthreads must be positive");
  return cfg.samples_per_pixel / cfg.nthreads * cfg.nthreads;
}

namespace {
struct Prepared {
  int width = 0, height = 0, spp = 0;
  rtw_render_cfg rc{};
};
Prepared prepare(const Config& cfg, const DeviceOptions& opt) {
  Prepared p;
  p.width = cfg.image_width;
  p.height = image_height(cfg);
  p.spp = effective_spp(cfg);
  // The reference divides by zero here (nthreads > spp) and prints a NaN image; refuse instead.
  if (p.spp <= 0) throw std::invalid_argument("samples_per_pixel / nthreads * nthreads is 0: nothing to render");
  if (p.width < 2 || p.height < 2) throw std::invalid_argument("image must be at least 2x2");
  p.rc.width = p.width;
  p.rc.height = p.height;
  p.rc.sample_begin = 0;
  p.rc.sample_end = p.spp;
  p.rc.max_child_rays = cfg.max_child_rays;
  p.rc.kernel = opt.kernel;
  p.rc.seed = opt.seed;
  p.rc.device = opt.device;
  p.rc.flags = (opt.stats ? RTW_FLAG_STATS : 0) | (opt.split_rows && opt.ngpus > 1 ? RTW_FLAG_SPLIT_ROWS : 0) |
               (opt.bvh_build == 1 ? RTW_FLAG_BVH_BUILD_HOST : 0) | (opt.bvh_build == 2 ? RTW_FLAG_BVH_BUILD_GPU : 0);
  p.rc.row_tile_rows = opt.tile_rows;
  return p;
}
[[noreturn]] void throw_last() { throw std::runtime_error(std::string("rtw_b200: ") + rtw_last_error()); }
}  // namespace

Accum render_accum(const Scene& world, const Config& cfg) {
  const DeviceOptions& opt = device_options();
  const Prepared p = prepare(cfg, opt);
  rtw_prewarm(opt.ngpus > 1 ? 0 : opt.device, opt.ngpus);   // contexts come up while the scene is flattened
  Accum img;
  img.width = p.width; img.height = p.height; img.spp = p.spp;
  const Scene::Flat flat = world.flatten();
  img.rgba.assign(static_cast<std::size_t>(img.width) * static_cast<std::size_t>(img.height) * 4, 0.0f);
  const int status = opt.ngpus > 1 ? rtw_render_multi_gpu(&flat.desc, &p.rc, opt.ngpus, img.rgba.data(), &img.stats)
                                   : rtw_render(&flat.desc, &p.rc, img.rgba.data(), &img.stats);
  if (status != 0) throw_last();
  return img;
}

Image8 render_rgb8(const Scene& world, const Config& cfg) {
  const DeviceOptions& opt = device_options();
  const Prepared p = prepare(cfg, opt);
  rtw_prewarm(opt.ngpus > 1 ? 0 : opt.device, opt.ngpus);
  Image8 img;
  img.width = p.width; img.height = p.height; img.spp = p.spp;
  const Scene::Flat flat = world.flatten();
  img.rgb.resize(static_cast<std::size_t>(img.width) * static_cast<std::size_t>(img.height) * 3);
  const int status = opt.ngpus > 1 ? rtw_render_multi_gpu_rgb8(&flat.desc, &p.rc, opt.ngpus, img.rgb.data(), &img.stats)
                                   : rtw_render_rgb8(&flat.desc, &p.rc, img.rgb.data(), &img.stats);
  if (status != 0) throw_last();
  return img;
}

void write_ppm(std::ostream& out, const Accum& img) {
  // c = sqrt(sum / spp); int(256 * clamp(c, 0, 0.999)) per channel, "r g b\n" per pixel, top row first
  std::string text;
  text.reserve(static_cast<std::size_t>(img.width) * static_cast<std::size_t>(img.height) * 12 + 32);
  text += "P3\n" + std::to_string(img.width) + ' ' + std::to_string(img.height) + "\n255\n";
  const std::size_t npix = static_cast<std::size_t>(img.width) * static_cast<std::size_t>(img.height);
  const double spp = static_cast<double>(img.spp);
  char buf[4];
  for (std::size_t k = 0; k < npix; ++k) {
    for (int ch = 0; ch < 3; ++ch) {
      const double c = std::sqrt(static_cast<double>(img.rgba[4 * k + ch]) / spp);
      int v = static_cast<int>(256 * std::clamp(c, 0.0, 0.999));
      int n = 0;
      do { buf[n++] = static_cast<char>('0' + v % 10); v /= 10; } while (v > 0);
      while (n > 0) text += buf[--n];
      text += ch == 2 ? '\n' : ' ';
    }
  }
  out.write(text.data(), static_cast<std::streamsize>(text.size()));
}

void write_ppm(std::ostream& out, const Image8& img) {
  // the same text from already quantised channels: "r g b\n" per pixel, top row first (render.cpp:182-186)
  char digits[256][4];
  int len[256];
  for (int v = 0; v < 256; ++v) len[v] = std::snprintf(digits[v], sizeof digits[v], "%d", v);
  const std::size_t npix = static_cast<std::size_t>(img.width) * static_cast<std::size_t>(img.height);
  std::string text;
  text.resize(npix * 12 + 32);
  char* w = text.data();
  w += std::snprintf(w, 32, "P3\n%d %d\n255\n", img.width, img.height);
  const std::uint8_t* px = img.rgb.data();
  for (std::size_t k = 0; k < npix; ++k, px += 3) {
    for (int ch = 0; ch < 3; ++ch) {
      const int v = px[ch];
      std::memcpy(w, digits[v], static_cast<std::size_t>(len[v]));
      w += len[v];
      *w++ = ch == 2 ? '\n' : ' ';
    }
  }
  out.write(text.data(), w - text.data());
}

void write_ppm_binary(std::ostream& out, const Image8& img) {
  out << "P6\n" << img.width << ' ' << img.height << "\n255\n";
  out.write(reinterpret_cast<const char*>(img.rgb.data()), static_cast<std::streamsize>(img.rgb.size()));
}

Image8 quantise(const Accum& img, int device) {
  Image8 q;
  q.width = img.width; q.height = img.height; q.spp = img.spp; q.stats = img.stats;
  const std::size_t npix = static_cast<std::size_t>(img.width) * static_cast<std::size_t>(img.height);
  q.rgb.resize(npix * 3);
  if (rtw_finalize_rgb8(img.rgba.data(), static_cast<int64_t>(npix), img.spp, device, q.rgb.data()) != 0) throw_last();
  return q;
}

namespace {
// checkpoint file: header + width*height*4 floats (sums + sample counts).  The header binds the file to its render: image size,
// sample target, depth, Philox seed and the hash of the flattened scene (prims, mats, camera).
struct CheckpointHeader {
  char magic[8];
  std::int32_t width, height, spp, cursor, max_child_rays, reserved;
  std::uint64_t seed;
  std::uint64_t scene_hash;
};
constexpr char kCheckpointMagic[8] = {'R', 'T', 'W', 'C', 'K', 'P', 'T', '2'};

// Reads the header, checks it against the render that wants to resume (`want`: everything but the cursor), and only then sizes and
// reads the buffer.  Returns false when there is no file.
bool read_checkpoint(const std::string& path, const CheckpointHeader& want, CheckpointHeader& h, std::vector<float>& rgba) {
  std::ifstream in(path, std::ios::binary);
  if (!in) return false;
  in.read(reinterpret_cast<char*>(&h), sizeof h);
  if (!in || std::memcmp(h.magic, kCheckpointMagic, 8) != 0) throw std::runtime_error("checkpoint " + path + ": not a checkpoint file");
  if (h.width != want.width || h.height != want.height || h.max_child_rays != want.max_child_rays || h.seed != want.seed)
    throw std::invalid_argument("checkpoint " + path + " belongs to a different render (size, depth or seed differ)");
  if (h.scene_hash != want.scene_hash) throw std::invalid_argument("checkpoint " + path + " belongs to a different scene");
  if (h.cursor < 0 || h.cursor > h.spp) throw std::runtime_error("checkpoint " + path + ": sample cursor out of range");
  rgba.resize(static_cast<std::size_t>(want.width) * static_cast<std::size_t>(want.height) * 4);
  in.read(reinterpret_cast<char*>(rgba.data()), static_cast<std::streamsize>(rgba.size() * sizeof(float)));
  if (!in) throw std::runtime_error("checkpoint " + path + ": truncated");
  return true;
}
void write_checkpoint(const std::string& path, const CheckpointHeader& h, const std::vector<float>& rgba) {
  const std::string tmp = path + ".tmp";
  {
    std::ofstream out(tmp, std::ios::binary | std::ios::trunc);
    out.write(reinterpret_cast<const char*>(&h), sizeof h);
    out.write(reinterpret_cast<const char*>(rgba.data()), static_cast<std::streamsize>(rgba.size() * sizeof(float)));
    if (!out) throw std::runtime_error("checkpoint " + tmp + ": write failed");
  }
  if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("checkpoint " + path + ": rename failed");
}
}  // namespace

Accum render_progressive(const Scene& world, const Config& cfg, const std::string& path, int every) {
  const DeviceOptions& opt = device_options();
  const Prepared p = prepare(cfg, opt);
  rtw_prewarm(opt.ngpus > 1 ? 0 : opt.device, opt.ngpus);
  Accum img;
  img.width = p.width; img.height = p.height; img.spp = p.spp;
  const Scene::Flat flat = world.flatten();
  CheckpointHeader h{};
  std::memcpy(h.magic, kCheckpointMagic, 8);
  h.width = img.width; h.height = img.height; h.spp = img.spp; h.cursor = 0; h.max_child_rays = cfg.max_child_rays; h.seed = opt.seed;
  if (rtw_scene_hash(&flat.desc, &h.scene_hash) != 0) throw_last();
  CheckpointHeader old{};
  if (read_checkpoint(path, h, old, img.rgba)) {
    h.cursor = std::min(old.cursor, img.spp);
    std::cerr << "resuming at sample " << h.cursor << " of " << img.spp << "\n";
  } else {
    img.rgba.assign(static_cast<std::size_t>(img.width) * static_cast<std::size_t>(img.height) * 4, 0.0f);
  }
  std::vector<float> slice(img.rgba.size());
  const int step = every > 0 ? every : img.spp;
  while (h.cursor < img.spp) {
    rtw_render_cfg rc = p.rc;
    rc.sample_begin = h.cursor; rc.sample_end = std::min(img.spp, h.cursor + step);
    rtw_stats st{};
    // the device(s) keep the scene from the previous slice: only the first slice pays for flatten + BVH build + upload
    const int status = opt.ngpus > 1 ? rtw_render_multi_gpu(&flat.desc, &rc, opt.ngpus, slice.data(), &st) : rtw_render(&flat.desc, &rc, slice.data(), &st);
    if (status != 0) throw_last();
    // every slice is an exact 2^-32 fixed-point sum converted to float; slices add in double so that the running total does not
    // depend on where the render was interrupted beyond one float rounding per slice
    for (std::size_t k = 0; k < img.rgba.size(); ++k) img.rgba[k] = static_cast<float>(static_cast<double>(img.rgba[k]) + static_cast<double>(slice[k]));
    img.stats.paths += st.paths; img.stats.rays += st.rays; img.stats.kernel_ms += st.kernel_ms;
    img.stats.h2d_ms += st.h2d_ms; img.stats.d2h_ms += st.d2h_ms;
    h.cursor = rc.sample_end;
    write_checkpoint(path, h, img.rgba);
    std::cerr << "\rsamples " << h.cursor << "/" << img.spp << std::flush;
  }
  std::cerr << "\n";
  return img;
}

void render(const Scene& world, const Config& cfg) {
  namespace khr = std::chrono;
  const DeviceOptions& opt = device_options();
  std::cerr << "Started rendering on " << opt.ngpus << " GPU(s)\n";
  const auto start = khr::steady_clock::now();
  const auto ms_since = [](khr::steady_clock::time_point t0) { return khr::duration<double, std::milli>(khr::steady_clock::now() - t0).count(); };
  // one-shot renders bring back the quantised image (3 bytes per pixel); progressive ones keep the linear buffer for the checkpoint
  const Image8 img = opt.checkpoint.empty() ? render_rgb8(world, cfg) : quantise(render_progressive(world, cfg, opt.checkpoint, opt.checkpoint_every), opt.device);
  const double render_ms = ms_since(start);
  const auto t_out = khr::steady_clock::now();
  if (opt.binary_ppm) write_ppm_binary(std::cout, img);
  else write_ppm(std::cout, img);
  std::cout.flush();
  const double out_ms = ms_since(t_out);
  const auto took = khr::duration_cast<khr::milliseconds>(khr::steady_clock::now() - start);
  const double paths = static_cast<double>(img.stats.paths), rays = static_cast<double>(img.stats.rays);
  std::cerr << "kernel " << img.stats.kernel_ms << " ms: " << paths / (img.stats.kernel_ms * 1e3) << " Mpaths/s, "
            << rays / (img.stats.kernel_ms * 1e3) << " Mrays/s (" << rays / std::max(paths, 1.0) << " rays/path)\n";
  // where the wall time went: the first CUDA call of a process creates the context(s) (0.3-1 s per GPU on a cold box)
  std::cerr << "host: render call " << render_ms << " ms (contexts + flatten + upload " << img.stats.h2d_ms;
  if (img.stats.bvh_build_gpu_ms > 0) std::cerr << " of which BVH build on the device " << img.stats.bvh_build_gpu_ms;
  std::cerr << ", combine + download " << img.stats.d2h_ms << "), image output " << out_ms << " ms\n";
  std::cerr << "\nDone in " << took.count() << "ms\n";
}

std::ostream& operator<<(std::ostream& o, const Config& c) {
  // same seven lines, same order, as the reference's printer (render.cpp:193-203; `model` is not printed there either)
  o << "Config {\n";
  o << "aspect_ratio: " << c.aspect_ratio << "\n";
  o << "number_of_balls_sqrt: " << c.number_of_balls_sqrt << "\n";
  o << "moving_spheres: " << c.moving_spheres << "\n";
  o << "image_width: " << c.image_width << "\n";
  o << "samples_per_pixel: " << c.samples_per_pixel << "\n";
  o << "max_child_rays: " << c.max_child_rays << "\n";
  o << "nthreads: " << c.nthreads << "\n";
  return o << "}\n";
}

}  // namespace rtweekend::detail
