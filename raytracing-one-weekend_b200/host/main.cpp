// main.cpp -- command line of the B200 renderer.  Same options as the reference executable (main.cpp:147-159):
//   -t,--threads  -w,--image-width  -s,--samples-per-pixel  -c,--max-child-rays  -a,--aspect-ratio  -n,--balls_sqrt
//   -m,--moving-spheres  -q,--quick  --dry-run  -l,--load
// plus device options: --gpus N, --split spp|rows, --tile-rows R, --seed S, --kernel auto|spheres|bvh, --stats,
// --bvh-build auto|host|gpu, --format p3|p6, --checkpoint FILE [--checkpoint-every N] (progressive / resumable), --scene cover|model|mesh-on-ground,
// and a mesh utility: --make-mesh OUT.obj --rounds K (high-poly stand-in generated from the -l model).
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "obj-loader.h"
#include "render.h"
#include "scenes.h"

namespace rt = rtweekend;

namespace {

struct Option {
  bool takes_value;
  std::function<void(const std::string&)> apply;
};

template <typename T>
T parse_number(const std::string& name, const std::string& text) {
  try {
    std::size_t used = 0;
    T v;
    if constexpr (std::is_same_v<T, double>) v = std::stod(text, &used);
    else if constexpr (std::is_same_v<T, std::uint64_t>) v = std::stoull(text, &used, 0);
    else v = static_cast<T>(std::stoi(text, &used));
    if (used != text.size()) throw std::invalid_argument(text);
    return v;
  } catch (const std::exception&) {
    throw std::invalid_argument("Could not convert '" + text + "' for " + name);
  }
}

}  // namespace

int main(int argc, char* argv[]) {
  rt::Config cfg{};
  rt::DeviceOptions& dev = rt::device_options();
  bool quick = false, dry_run = false;
  std::string scene_name, make_mesh;
  int rounds = 5;

  std::map<std::string, Option> table;
  auto number = [&](std::initializer_list<const char*> names, auto& target) {
    using T = std::remove_reference_t<decltype(target)>;
    for (const char* n : names) table[n] = {true, [&target, n](const std::string& v) { target = parse_number<T>(n, v); }};
  };
  auto flag = [&](std::initializer_list<const char*> names, bool& target) {
    for (const char* n : names) table[n] = {false, [&target](const std::string&) { target = true; }};
  };
  number({"-t", "--threads"}, cfg.nthreads);
  number({"-w", "--image-width"}, cfg.image_width);
  number({"-s", "--samples-per-pixel"}, cfg.samples_per_pixel);
  number({"-c", "--max-child-rays"}, cfg.max_child_rays);
  number({"-a", "--aspect-ratio"}, cfg.aspect_ratio);
  number({"-n", "--balls_sqrt"}, cfg.number_of_balls_sqrt);
  flag({"-m", "--moving-spheres"}, cfg.moving_spheres);
  flag({"-q", "--quick"}, quick);
  flag({"--dry-run"}, dry_run);
  for (const char* n : {"-l", "--load"}) table[n] = {true, [&cfg](const std::string& v) { cfg.model = v; }};
  number({"--gpus"}, dev.ngpus);
  number({"--seed"}, dev.seed);
  number({"--device"}, dev.device);
  flag({"--stats"}, dev.stats);
  flag({"--all-shapes"}, dev.obj_all_shapes);
  number({"--rounds"}, rounds);
  number({"--tile-rows"}, dev.tile_rows);
  number({"--checkpoint-every"}, dev.checkpoint_every);
  table["--checkpoint"] = {true, [&dev](const std::string& v) { dev.checkpoint = v; }};
  table["--format"] = {true, [&dev](const std::string& v) {
    if (v == "p3") dev.binary_ppm = false;
    else if (v == "p6") dev.binary_ppm = true;
    else throw std::invalid_argument("--format must be p3 or p6");
  }};
  table["--split"] = {true, [&dev](const std::string& v) {
    if (v == "spp") dev.split_rows = false;
    else if (v == "rows") dev.split_rows = true;
    else throw std::invalid_argument("--split must be spp or rows");
  }};
  table["--bvh-build"] = {true, [&dev](const std::string& v) {
    if (v == "auto") dev.bvh_build = 0;
    else if (v == "host") dev.bvh_build = 1;
    else if (v == "gpu") dev.bvh_build = 2;
    else throw std::invalid_argument("--bvh-build must be auto, host or gpu");
  }};
  table["--static-spheres"] = {false, [&cfg](const std::string&) { cfg.moving_spheres = false; }};
  table["--scene"] = {true, [&scene_name](const std::string& v) { scene_name = v; }};
  table["--make-mesh"] = {true, [&make_mesh](const std::string& v) { make_mesh = v; }};
  table["--kernel"] = {true, [&dev](const std::string& v) {
    if (v == "auto") dev.kernel = RTW_KERNEL_AUTO;
    else if (v == "spheres") dev.kernel = RTW_KERNEL_SPHERES_SMEM;
    else if (v == "bvh") dev.kernel = RTW_KERNEL_BVH;
    else if (v == "bvh-perlane") dev.kernel = RTW_KERNEL_BVH_PERLANE;
    else throw std::invalid_argument("--kernel must be auto, spheres, bvh or bvh-perlane");
  }};

  try {
    for (int i = 1; i < argc; ++i) {
      std::string arg = argv[i], value;
      bool has_value = false;
      if (arg == "-h" || arg == "--help") {
        std::cout << "Raytracing one weekend/week/restoflife (B200)\nOptions:";
        for (const auto& [name, o] : table) std::cout << ' ' << name << (o.takes_value ? " <v>" : "");
        std::cout << "\n";
        return 0;
      }
      if (arg.rfind("--", 0) == 0) {
        if (const auto eq = arg.find('='); eq != std::string::npos) { value = arg.substr(eq + 1); arg.erase(eq); has_value = true; }
      } else if (arg.size() > 2 && arg[0] == '-') {
        value = arg.substr(2); arg.erase(2); has_value = true;
      }
      const auto it = table.find(arg);
      if (it == table.end()) throw std::invalid_argument("The following argument was not expected: " + std::string(argv[i]));
      if (it->second.takes_value && !has_value) {
        if (i + 1 >= argc) throw std::invalid_argument(arg + ": 1 required");
        value = argv[++i];
      }
      it->second.apply(value);
    }
  } catch (const std::invalid_argument& e) {
    std::cerr << e.what() << "\nRun with --help for more information.\n";
    return 105;
  }

  if (dry_run) { std::cout << cfg; return 0; }
  // The first CUDA call of a process initialises EVERY visible GPU (about 0.7 s each on a box without the persistence daemon: 5.5 s on
  // an 8-GPU node before the first kernel): make only the GPUs this render uses visible, unless the user already chose.
  if (make_mesh.empty() && !std::getenv("CUDA_VISIBLE_DEVICES") && dev.ngpus >= 1) {
    const int first = dev.ngpus > 1 ? 0 : dev.device;
    std::string list;
    for (int g = 0; g < dev.ngpus; ++g) list += (g ? "," : "") + std::to_string(first + g);
    setenv("CUDA_VISIBLE_DEVICES", list.c_str(), 1);
    dev.device = 0;   // ordinals now count the visible devices
  }
  // the CUDA contexts come up on background threads while the scene is built
  if (make_mesh.empty()) rtw_prewarm(dev.ngpus > 1 ? 0 : dev.device, dev.ngpus);

  if (!make_mesh.empty()) {
    if (!cfg.model) { std::cerr << "--make-mesh needs -l <base.obj>\n"; return 105; }
    const auto base = rt::detail::load_obj(*cfg.model);
    const auto fine = rt::detail::subdivide_displace(base, rounds, 20221018u, 0.08);
    rt::detail::save_obj(make_mesh, fine);
    std::cerr << "wrote " << make_mesh << ": " << fine.vertices.size() << " vertices, " << fine.faces.size() << " triangles\n";
    return 0;
  }

  if (scene_name.empty()) scene_name = cfg.model ? "model" : "cover";
  if (scene_name == "model") rt::render(rt::foo(cfg), cfg);
  else if (scene_name == "mesh-on-ground") rt::render(rt::mesh_on_ground(cfg), cfg);
  else if (scene_name == "cover") rt::render(rt::lots_of_balls(cfg), cfg);
  else { std::cerr << "--scene must be cover, model or mesh-on-ground\n"; return 105; }
  return 0;
}
