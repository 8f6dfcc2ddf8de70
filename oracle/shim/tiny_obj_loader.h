// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the part of tinyobjloader 1.0.6
// (conanfile.txt:4; built with TINYOBJLOADER_USE_DOUBLE, main.cpp:9) that main.cpp:103-133
// consumes: attrib.vertices, shapes[0].mesh.num_face_vertices, shapes[0].mesh.indices[].vertex_index.
//
// Restated behaviour of the published 1.0.6 loader for that subset:
//  * `v x y z` appends three reals; reals are parsed with tinyobj's own digit-accumulating
//    routine (integer digits: m = m*10 + d; fraction digit k: m += d * 10^-k using a short
//    table then pow(); optional exponent assembled with ldexp(m*5^e, e)), NOT strtod, so the
//    doubles carry tinyobj's last-bit rounding;
//  * `f` accepts i, i/j, i//k, i/j/k with 1-based or negative (relative) indices and, because
//    LoadObj's `triangulate` argument defaults to true, polygons are fan-triangulated;
//  * a new shape starts at a `g`/`o` line once the current one holds faces;
//  * vt / vn / usemtl / mtllib / s lines do not affect the subset the reference reads.
#pragma once
#include <cmath>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

namespace tinyobj {

#ifdef TINYOBJLOADER_USE_DOUBLE
using real_t = double;
#else
using real_t = float;
#endif

struct index_t { int vertex_index; int normal_index; int texcoord_index; };
struct mesh_t {
  std::vector<index_t> indices;
  std::vector<unsigned char> num_face_vertices;
  std::vector<int> material_ids;
};
struct shape_t { std::string name; mesh_t mesh; };
struct material_t { std::string name; };
struct attrib_t {
  std::vector<real_t> vertices;
  std::vector<real_t> normals;
  std::vector<real_t> texcoords;
};

bool LoadObj(attrib_t* attrib, std::vector<shape_t>* shapes, std::vector<material_t>* materials,
             std::string* err, const char* filename, const char* mtl_basedir = nullptr,
             bool triangulate = true);

#ifdef TINYOBJLOADER_IMPLEMENTATION
namespace shim_detail {
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

inline bool parse_real_tinyobj(const char* s, const char* s_end, double* out) {
  if (s >= s_end) return false;
  double m = 0.0; int ex = 0; bool neg = false, exneg = false; int nread = 0;
  const char* p = s;
  if (*p == '+' || *p == '-') { neg = (*p == '-'); ++p; }
  else if (!is_digit(*p)) return false;
  while (p != s_end && is_digit(*p)) { m *= 10; m += static_cast<int>(*p - '0'); ++p; ++nread; }
  if (nread == 0) return false;
  if (p != s_end) {
    bool go_exp = false;
    if (*p == '.') {
      ++p; int k = 1;
      static const double lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
      const int nlut = static_cast<int>(sizeof lut / sizeof lut[0]);
      while (p != s_end && is_digit(*p)) {
        m += static_cast<int>(*p - '0') * (k < nlut ? lut[k] : std::pow(10.0, -k));
        ++k; ++p;
      }
      go_exp = (p != s_end);
    } else if (*p == 'e' || *p == 'E') {
      go_exp = true;
    }
    if (go_exp && (*p == 'e' || *p == 'E')) {
      ++p;
      if (p != s_end && (*p == '+' || *p == '-')) { exneg = (*p == '-'); ++p; }
      else if (p == s_end || !is_digit(*p)) return false;
      int n = 0;
      while (p != s_end && is_digit(*p)) { ex *= 10; ex += static_cast<int>(*p - '0'); ++p; ++n; }
      if (exneg) ex = -ex;
      if (n == 0) return false;
    }
  }
  *out = (neg ? -1 : 1) * (ex ? std::ldexp(m * std::pow(5.0, ex), ex) : m);
  return true;
}

inline double next_real(const char*& p) {
  p += std::strspn(p, " \t");
  const char* e = p + std::strcspn(p, " \t\r");
  double v = 0.0;
  parse_real_tinyobj(p, e, &v);
  p = e;
  return v;
}
inline int fix_index(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); }
}  // namespace shim_detail

inline bool LoadObj(attrib_t* attrib, std::vector<shape_t>* shapes, std::vector<material_t>* materials,
                    std::string* err, const char* filename, const char*, bool triangulate) {
  attrib->vertices.clear(); attrib->normals.clear(); attrib->texcoords.clear();
  shapes->clear(); if (materials) materials->clear();
  std::ifstream in(filename);
  if (!in) { if (err) *err = std::string("Cannot open file [") + filename + "]\n"; return false; }
  shape_t cur; std::string line;
  auto flush = [&]() { if (!cur.mesh.num_face_vertices.empty()) shapes->push_back(cur); cur = shape_t{}; };
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    const char* p = line.c_str();
    p += std::strspn(p, " \t");
    if (*p == '\0' || *p == '#') continue;
    if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
      p += 2;
      for (int k = 0; k < 3; ++k) attrib->vertices.push_back(static_cast<real_t>(shim_detail::next_real(p)));
    } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
      p += 2;
      std::vector<index_t> face;
      const int nv = static_cast<int>(attrib->vertices.size() / 3);
      while (true) {
        p += std::strspn(p, " \t");
        if (*p == '\0') break;
        index_t ix{-1, -1, -1};
        ix.vertex_index = shim_detail::fix_index(std::atoi(p), nv);
        p += std::strcspn(p, "/ \t");
        if (*p == '/') { ++p; if (*p != '/') { ix.texcoord_index = std::atoi(p) - 1; p += std::strcspn(p, "/ \t"); }
          if (*p == '/') { ++p; ix.normal_index = std::atoi(p) - 1; p += std::strcspn(p, " \t"); } }
        face.push_back(ix);
      }
      if (triangulate && face.size() > 3) {
        for (size_t k = 2; k < face.size(); ++k) {
          cur.mesh.indices.push_back(face[0]); cur.mesh.indices.push_back(face[k - 1]); cur.mesh.indices.push_back(face[k]);
          cur.mesh.num_face_vertices.push_back(3); cur.mesh.material_ids.push_back(-1);
        }
      } else {
        for (auto& ix : face) cur.mesh.indices.push_back(ix);
        cur.mesh.num_face_vertices.push_back(static_cast<unsigned char>(face.size()));
        cur.mesh.material_ids.push_back(-1);
      }
    } else if ((p[0] == 'g' || p[0] == 'o') && (p[1] == ' ' || p[1] == '\t' || p[1] == '\0')) {
      flush();
      cur.name = (p[1] ? std::string(p + 2) : std::string());
    }
  }
  flush();
  return true;
}
#endif  // TINYOBJLOADER_IMPLEMENTATION
}  // namespace tinyobj
