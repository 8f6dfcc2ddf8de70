// obj-loader.h -- the slice of Wavefront OBJ the reference consumes through tinyobjloader 1.0.6 (main.cpp:103-133):
// `v` positions and `f` faces of the FIRST shape, triangles only (polygons arrive fan-triangulated because LoadObj's
// `triangulate` defaults to true).  Reals are parsed with tinyobj's digit-accumulating algorithm, not strtod, so
// vertices carry the same doubles the reference sees.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "vec3.h"

namespace rtweekend::detail {
struct ObjMesh {
  std::vector<point> vertices;
  std::vector<std::array<int, 3>> faces;  // zero-based vertex indices, first shape only
};
// all_shapes = false: the reference's behaviour (faces of shapes[0] only, main.cpp:117); true: every `f` of the file
ObjMesh load_obj(const std::string& path, bool all_shapes = false);  // throws std::runtime_error("Can't load because ...")
void save_obj(const std::string& path, const ObjMesh& mesh);
// deterministic high-poly stand-in for the missing dragon.obj: `rounds` 1->4 midpoint subdivisions with a
// fixed-seed radial displacement
ObjMesh subdivide_displace(const ObjMesh& mesh, int rounds, std::uint32_t seed, double amplitude);
}  // namespace rtweekend::detail
