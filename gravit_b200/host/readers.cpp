// readers.cpp -- OBJ/MTL, PLY and .conf ingestion behind the reference's reader classes (see readers.h).
//
// The OBJ path restates what tinyobjloader (third-party/tinyobjloader-master/tiny_obj_loader.cc, the version vendored
// by the reference) produces for ObjReader.cpp:129-203, so that vertex order, face order, material ids and every float
// bit match the reference's Mesh:
//   * numbers: sign, integer digits accumulated as m = m*10 + d, fraction digits as m += d * pow(10, -k), optional
//     exponent e, value = sign * ldexp(m * pow(5, e), e), then rounded to float (tiny_obj_loader.cc:127-264) -- NOT strtod
//   * indices: 1-based -> 0-based, negative = relative to the vertices read so far, 0 stays 0 (:79-83)
//   * faces: polygons become triangle fans (f0, f[k-1], f[k]) (:389-414)
//   * shapes: the pending faces are flushed into a new shape at every `usemtl`, `g` and `o` (and at the end of the file);
//     inside one shape the (v, vt, vn) index triples are de-duplicated in first-use order and the shape gets its own
//     copy of the positions (:296-341; the cache is passed by value, so it never survives a shape)
//   * materials: `mtllib` appends the library's materials; LoadMtl always emits the material it is holding at the end of
//     the file, even an unnamed one (:466-487,700-703); `usemtl` of an unknown name gives id -1
// ObjReader then concatenates the shapes: vertices in shape order, faces offset by the vertices of the earlier shapes
// (ObjReader.cpp:171-203).
#include "readers.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <glob.h>
#include <map>
#include <sstream>
#include <stdexcept>
#include <sys/stat.h>
#include <tuple>

#include "api.h"

namespace {

inline bool is_blank(char c) { return c == ' ' || c == '\t'; }
inline bool is_eol(char c) { return c == '\r' || c == '\n' || c == '\0'; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// the reference's number grammar and arithmetic (see the header comment); false = not a number, *out untouched
bool parse_number(const char *s, const char *end, double *out) {
  if (s >= end) return false;
  double mant = 0.0;
  int expo = 0;
  bool neg = false, eneg = false;
  const char *p = s;
  if (*p == '+' || *p == '-') {
    neg = (*p == '-');
    ++p;
  } else if (!is_digit(*p)) {
    return false;
  }
  int read = 0;
  while (p != end && is_digit(*p)) {
    mant *= 10;
    mant += (int)(*p - '0');
    ++p, ++read;
  }
  if (read == 0) return false;
  bool more = (p != end);
  if (more) {
    bool exp_follows = false;
    if (*p == '.') {
      ++p;
      read = 1;
      while (p != end && is_digit(*p)) {
        mant += (int)(*p - '0') * std::pow(10.0, -read);
        ++read, ++p;
      }
      exp_follows = (p != end);
    } else if (*p == 'e' || *p == 'E') {
      exp_follows = true;
    }
    if (exp_follows && (*p == 'e' || *p == 'E')) {
      ++p;
      if (p != end && (*p == '+' || *p == '-')) {
        eneg = (*p == '-');
        ++p;
      } else if (!(p != end && is_digit(*p))) {
        return false; // an empty exponent is not a number
      }
      read = 0;
      while (p != end && is_digit(*p)) {
        expo *= 10;
        expo += (int)(*p - '0');
        ++p, ++read;
      }
      if (eneg) expo = -expo;
      if (read == 0) return false;
    }
  }
  *out = (neg ? -1 : 1) * std::ldexp(mant * std::pow(5.0, expo), expo);
  return true;
}

float next_float(const char *&tok) {
  tok += strspn(tok, " \t");
  const char *end = tok + strcspn(tok, " \t\r");
  double v = 0.0;
  parse_number(tok, end, &v);
  tok = end;
  return (float)v;
}
std::string next_word(const char *tok) { // sscanf("%s")
  tok += strspn(tok, " \t\r\n");
  size_t n = strcspn(tok, " \t\r\n");
  return std::string(tok, n);
}
inline int fix_index(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); }

struct Corner {
  int v = -1, vt = -1, vn = -1;
  bool operator<(const Corner &o) const { // tiny_obj_loader.cc:55-65: v, then vn, then vt
    return std::tie(v, vn, vt) < std::tie(o.v, o.vn, o.vt);
  }
};
Corner parse_corner(const char *&tok, int nv, int nvn, int nvt) { // i, i/j, i//k, i/j/k
  Corner c;
  c.v = fix_index(atoi(tok), nv);
  tok += strcspn(tok, "/ \t\r");
  if (tok[0] != '/') return c;
  ++tok;
  if (tok[0] == '/') {
    ++tok;
    c.vn = fix_index(atoi(tok), nvn);
    tok += strcspn(tok, "/ \t\r");
    return c;
  }
  c.vt = fix_index(atoi(tok), nvt);
  tok += strcspn(tok, "/ \t\r");
  if (tok[0] != '/') return c;
  ++tok;
  c.vn = fix_index(atoi(tok), nvn);
  tok += strcspn(tok, "/ \t\r");
  return c;
}

struct ObjMaterial {
  std::string name;
  float ambient[3] = { 0, 0, 0 }, diffuse[3] = { 0, 0, 0 }, specular[3] = { 0, 0, 0 };
  float shininess = 1.f, dissolve = 1.f;
};
struct ObjShape {
  std::vector<float> positions, normals;
  std::vector<unsigned> indices;
  std::vector<int> material_ids;
};

bool read_line(std::istream &in, std::string &line) {
  if (in.peek() == -1) return false;
  std::getline(in, line);
  if (!line.empty() && line.back() == '\n') line.pop_back();
  if (!line.empty() && line.back() == '\r') line.pop_back();
  return true;
}
inline bool key(const char *tok, const char *k) {
  size_t n = strlen(k);
  return strncmp(tok, k, n) == 0 && is_blank(tok[n]);
}

void load_mtl(std::istream &in, std::map<std::string, int> &by_name, std::vector<ObjMaterial> &out) {
  ObjMaterial cur;
  std::string line;
  while (read_line(in, line)) {
    if (line.empty()) continue;
    const char *tok = line.c_str();
    tok += strspn(tok, " \t");
    if (tok[0] == '\0' || tok[0] == '#') continue;
    if (key(tok, "newmtl")) {
      if (!cur.name.empty()) {
        by_name.insert(std::make_pair(cur.name, (int)out.size()));
        out.push_back(cur);
      }
      cur = ObjMaterial();
      cur.name = next_word(tok + 7);
      continue;
    }
    auto rgb = [&](float *dst) {
      tok += 2;
      dst[0] = next_float(tok), dst[1] = next_float(tok), dst[2] = next_float(tok);
    };
    if (key(tok, "Ka")) {
      rgb(cur.ambient);
    } else if (key(tok, "Kd")) {
      rgb(cur.diffuse);
    } else if (key(tok, "Ks")) {
      rgb(cur.specular);
    } else if (key(tok, "Ns")) {
      tok += 2;
      cur.shininess = next_float(tok);
    } else if (key(tok, "d")) {
      tok += 1;
      cur.dissolve = next_float(tok);
    } else if (key(tok, "Tr")) {
      tok += 2;
      cur.dissolve = 1.0f - next_float(tok);
    }
  }
  by_name.insert(std::make_pair(cur.name, (int)out.size())); // the held material is always emitted
  out.push_back(cur);
}

void flush_shape(std::vector<ObjShape> &shapes, std::vector<std::vector<Corner> > &group, int material, const std::vector<float> &v,
                 const std::vector<float> &vn) {
  if (group.empty()) return;
  ObjShape sh;
  std::map<Corner, unsigned> cache;
  auto vertex = [&](const Corner &c) -> unsigned {
    auto it = cache.find(c);
    if (it != cache.end()) return it->second;
    if ((size_t)(3 * c.v + 2) >= v.size() || c.v < 0) throw std::runtime_error("OBJ: vertex index out of range");
    sh.positions.insert(sh.positions.end(), { v[3 * c.v], v[3 * c.v + 1], v[3 * c.v + 2] });
    if (c.vn >= 0) {
      if ((size_t)(3 * c.vn + 2) >= vn.size()) throw std::runtime_error("OBJ: normal index out of range");
      sh.normals.insert(sh.normals.end(), { vn[3 * c.vn], vn[3 * c.vn + 1], vn[3 * c.vn + 2] });
    }
    unsigned idx = (unsigned)(sh.positions.size() / 3 - 1);
    cache[c] = idx;
    return idx;
  };
  for (const auto &face : group) {
    for (size_t k = 2; k < face.size(); ++k) { // fan
      unsigned a = vertex(face[0]), b = vertex(face[k - 1]), c = vertex(face[k]);
      sh.indices.insert(sh.indices.end(), { a, b, c });
      sh.material_ids.push_back(material);
    }
  }
  shapes.push_back(std::move(sh));
  group.clear();
}

void load_obj(const std::string &filename, const std::string &basepath, std::vector<ObjShape> &shapes, std::vector<ObjMaterial> &materials) {
  std::ifstream in(filename.c_str());
  if (!in) throw std::runtime_error("Cannot open file [" + filename + "]");
  std::vector<float> v, vn;
  size_t n_vt = 0;
  std::vector<std::vector<Corner> > group;
  std::map<std::string, int> by_name;
  int material = -1;
  std::string line;
  while (read_line(in, line)) {
    if (line.empty()) continue;
    const char *tok = line.c_str();
    tok += strspn(tok, " \t");
    if (tok[0] == '\0' || tok[0] == '#') continue;
    if (tok[0] == 'v' && is_blank(tok[1])) {
      tok += 2;
      float x = next_float(tok), y = next_float(tok), z = next_float(tok);
      v.insert(v.end(), { x, y, z });
    } else if (tok[0] == 'v' && tok[1] == 'n' && is_blank(tok[2])) {
      tok += 3;
      float x = next_float(tok), y = next_float(tok), z = next_float(tok);
      vn.insert(vn.end(), { x, y, z });
    } else if (tok[0] == 'v' && tok[1] == 't' && is_blank(tok[2])) {
      ++n_vt;
    } else if (tok[0] == 'f' && is_blank(tok[1])) {
      tok += 2;
      tok += strspn(tok, " \t");
      std::vector<Corner> face;
      while (!is_eol(tok[0])) {
        face.push_back(parse_corner(tok, (int)(v.size() / 3), (int)(vn.size() / 3), (int)n_vt));
        tok += strspn(tok, " \t\r");
      }
      group.push_back(face);
    } else if (key(tok, "usemtl")) {
      std::string name = next_word(tok + 7);
      flush_shape(shapes, group, material, v, vn);
      auto it = by_name.find(name);
      material = (it != by_name.end()) ? it->second : -1;
    } else if (key(tok, "mtllib")) {
      std::string name = next_word(tok + 7);
      std::ifstream mtl((basepath + name).c_str());
      load_mtl(mtl, by_name, materials); // a missing library still yields one default (all-zero) material
    } else if ((tok[0] == 'g' || tok[0] == 'o') && is_blank(tok[1])) {
      flush_shape(shapes, group, material, v, vn);
    }
  }
  flush_shape(shapes, group, material, v, vn);
}

inline float fastmin(float a, float b) { return a < b ? a : b; }
inline float fastmax(float a, float b) { return a > b ? a : b; }
void bounds_of(gvt::render::data::domain::reader::MeshData &m) { // Mesh::computeBoundingBox, Mesh.cpp:186-199
  for (int i = 0; i < 3; ++i) m.bounds_min[i] = FLT_MAX, m.bounds_max[i] = -FLT_MAX;
  for (size_t i = 0; i + 2 < m.vertices.size(); i += 3)
    for (int a = 0; a < 3; ++a) {
      m.bounds_min[a] = fastmin(m.bounds_min[a], m.vertices[i + a]);
      m.bounds_max[a] = fastmax(m.bounds_max[a], m.vertices[i + a]);
    }
}

// ---- PLY --------------------------------------------------------------------------------------------------------------
enum PlyType { T_NONE, T_I8, T_U8, T_I16, T_U16, T_I32, T_U32, T_F32, T_F64 };
PlyType ply_type(const std::string &s) { // ply.c type_names + old_type_names
  static const std::map<std::string, PlyType> names = {
    { "char", T_I8 },   { "int8", T_I8 },   { "uchar", T_U8 },    { "uint8", T_U8 },    { "short", T_I16 },  { "int16", T_I16 },
    { "ushort", T_U16 }, { "uint16", T_U16 }, { "int", T_I32 },     { "int32", T_I32 },   { "uint", T_U32 },   { "uint32", T_U32 },
    { "float", T_F32 },  { "float32", T_F32 }, { "double", T_F64 }, { "float64", T_F64 },
  };
  auto it = names.find(s);
  return it == names.end() ? T_NONE : it->second;
}
int ply_size(PlyType t) {
  switch (t) {
  case T_I8: case T_U8: return 1;
  case T_I16: case T_U16: return 2;
  case T_I32: case T_U32: case T_F32: return 4;
  case T_F64: return 8;
  default: return 0;
  }
}
struct PlyProp {
  std::string name;
  PlyType type = T_NONE, count_type = T_NONE; // count_type != NONE: a list
};
struct PlyElem {
  std::string name;
  size_t count = 0;
  std::vector<PlyProp> props;
};
struct PlyIn {
  std::ifstream f;
  int format = 0; // 0 ascii, 1 little endian, 2 big endian
  double scalar(PlyType t) {
    if (format == 0) {
      std::string w;
      if (!(f >> w)) throw std::runtime_error("PLY: unexpected end of file");
      return (t == T_F32 || t == T_F64) ? atof(w.c_str()) : (double)atoll(w.c_str()); // ply.c get_ascii_item
    }
    unsigned char b[8];
    int n = ply_size(t);
    f.read((char *)b, n);
    if (!f) throw std::runtime_error("PLY: unexpected end of file");
    if (format == 2) std::reverse(b, b + n);
    switch (t) {
    case T_I8: return (double)(int8_t)b[0];
    case T_U8: return (double)b[0];
    case T_I16: { int16_t x; memcpy(&x, b, 2); return x; }
    case T_U16: { uint16_t x; memcpy(&x, b, 2); return x; }
    case T_I32: { int32_t x; memcpy(&x, b, 4); return x; }
    case T_U32: { uint32_t x; memcpy(&x, b, 4); return x; }
    case T_F32: { float x; memcpy(&x, b, 4); return x; }
    case T_F64: { double x; memcpy(&x, b, 8); return x; }
    default: return 0;
    }
  }
};

bool path_exists(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0;
}
bool path_isdir(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
std::vector<std::string> find_ply(const std::string &dir) { // PlyReader.h findply: glob(dir/*.ply, GLOB_TILDE)
  glob_t g;
  std::vector<std::string> out;
  if (glob((dir + "/*.ply").c_str(), GLOB_TILDE, nullptr, &g) == 0)
    for (size_t i = 0; i < g.gl_pathc; ++i) out.push_back(g.gl_pathv[i]);
  globfree(&g);
  return out;
}

} // namespace

namespace gvt {
namespace render {
namespace data {
namespace domain {
namespace reader {

ObjReader::ObjReader(const std::string filename, int material_type) {
  std::size_t found = filename.find_last_of("/");
  std::string path = filename.substr(0, found + 1); // npos + 1 == 0: no directory part
  std::vector<ObjShape> shapes;
  std::vector<ObjMaterial> materials;
  load_obj(filename, path, shapes, materials);
  for (const ObjMaterial &m : materials) { // ObjReader.cpp:158-169
    MeshMaterial mm;
    memcpy(mm.ka, m.ambient, 12), memcpy(mm.kd, m.diffuse, 12), memcpy(mm.ks, m.specular, 12);
    mm.type = material_type;
    objMesh.materials.push_back(mm);
  }
  size_t offset = 0;
  for (const ObjShape &s : shapes) { // ObjReader.cpp:171-203
    const size_t nv = s.positions.size() / 3;
    objMesh.vertices.insert(objMesh.vertices.end(), s.positions.begin(), s.positions.end());
    if (!s.normals.empty()) {
      for (size_t v = 0; v < nv && 3 * v + 2 < s.normals.size(); ++v) {
        float x = s.normals[3 * v], y = s.normals[3 * v + 1], z = s.normals[3 * v + 2];
        float inv = 1.0f / std::sqrt(x * x + y * y + z * z); // glm::normalize
        objMesh.normals.insert(objMesh.normals.end(), { x * inv, y * inv, z * inv });
      }
    }
    for (size_t f = 0; f < s.indices.size() / 3; ++f) {
      for (int k = 0; k < 3; ++k) objMesh.faces.push_back((unsigned)(offset + s.indices[3 * f + k]));
      if (!materials.empty()) {
        // the reference indexes materials[-1] for an unknown usemtl (undefined behaviour); here that is an error
        if (s.material_ids[f] < 0) throw std::runtime_error("OBJ " + filename + ": face uses a material that no mtllib defines");
        objMesh.faces_to_materials.push_back(s.material_ids[f]);
      }
    }
    offset += nv;
  }
  bounds_of(objMesh);
}
ObjReader::~ObjReader() {}

void ObjReader::addToScene(const std::string &meshname) const {
  const MeshData &m = objMesh;
  api::createMesh(meshname);
  api::addMeshVertices(meshname, (unsigned)(m.vertices.size() / 3), m.vertices.data());
  api::addMeshTriangles0(meshname, (unsigned)(m.faces.size() / 3), m.faces.data());
  if (!m.materials.empty()) {
    const size_t nf = m.faces_to_materials.size();
    std::vector<unsigned> type(nf);
    std::vector<float> kd(3 * nf), ks(3 * nf), alpha(nf);
    for (size_t f = 0; f < nf; ++f) {
      const MeshMaterial &mm = m.materials[(size_t)m.faces_to_materials[f]];
      type[f] = (unsigned)mm.type, alpha[f] = mm.alpha;
      memcpy(&kd[3 * f], mm.kd, 12), memcpy(&ks[3 * f], mm.ks, 12);
    }
    api::addMeshMaterials(meshname, (unsigned)nf, type.data(), kd.data(), ks.data(), alpha.data());
  }
  api::finishMesh(meshname, true); // mesh->generateNormals(); computeBoundingBox() (SimpleFileLoadApp.cpp:156-158)
}

void PlyReader::readFile(const std::string &path, MeshData &out) {
  PlyIn in;
  in.f.open(path.c_str(), std::ios::in | std::ios::binary);
  if (!in.f) throw std::runtime_error("Cannot open file [" + path + "]");
  std::string line;
  std::getline(in.f, line);
  if (line.compare(0, 3, "ply") != 0) throw std::runtime_error("PLY " + path + ": missing magic");
  std::vector<PlyElem> elems;
  bool ended = false;
  while (std::getline(in.f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ss(line);
    std::string w;
    ss >> w;
    if (w == "format") {
      ss >> w;
      if (w == "ascii") in.format = 0;
      else if (w == "binary_little_endian") in.format = 1;
      else if (w == "binary_big_endian") in.format = 2;
      else throw std::runtime_error("PLY " + path + ": unknown format " + w);
    } else if (w == "element") {
      PlyElem e;
      ss >> e.name >> e.count;
      elems.push_back(e);
    } else if (w == "property") {
      if (elems.empty()) throw std::runtime_error("PLY " + path + ": property before element");
      PlyProp p;
      ss >> w;
      if (w == "list") {
        std::string ct, it;
        ss >> ct >> it >> p.name;
        p.count_type = ply_type(ct), p.type = ply_type(it);
      } else {
        p.type = ply_type(w);
        ss >> p.name;
      }
      if (p.type == T_NONE) throw std::runtime_error("PLY " + path + ": unknown property type");
      elems.back().props.push_back(p);
    } else if (w == "end_header") {
      ended = true;
      break;
    }
  }
  if (!ended) throw std::runtime_error("PLY " + path + ": no end_header");
  out = MeshData();
  for (const PlyElem &e : elems) {
    const bool is_vertex = (e.name == "vertex"), is_face = (e.name == "face");
    // PlyReader.cpp:113-118: colours are read when the vertex element has more than five properties
    const bool has_color = is_vertex && e.props.size() > 5;
    if (is_vertex) {
      out.vertices.resize(3 * e.count);
      if (has_color) out.vertex_colors.assign(3 * e.count, 0.f);
    }
    for (size_t i = 0; i < e.count; ++i) {
      for (const PlyProp &p : e.props) {
        if (p.count_type != T_NONE) {
          const int n = (int)in.scalar(p.count_type);
          const bool take = is_face && (p.name == "vertex_indices" || p.name == "vertex_index");
          long long idx[3] = { 0, 0, 0 };
          for (int k = 0; k < n; ++k) {
            double x = in.scalar(p.type);
            if (take && k < 3) idx[k] = (long long)x; // PlyReader.cpp:165-167: the first three of the list
          }
          if (take && n >= 3) {
            for (int k = 0; k < 3; ++k) { // the upper bound is checked by Mesh::addFace when the mesh is built
              if (idx[k] < 0) throw std::runtime_error("PLY " + path + ": negative vertex index");
              out.faces.push_back((unsigned)idx[k]);
            }
          }
        } else {
          const double x = in.scalar(p.type);
          if (!is_vertex) continue;
          if (p.name == "x") out.vertices[3 * i] = (float)x;
          else if (p.name == "y") out.vertices[3 * i + 1] = (float)x;
          else if (p.name == "z") out.vertices[3 * i + 2] = (float)x;
          else if (has_color && p.name == "red") out.vertex_colors[3 * i] = (float)(unsigned char)x / 255.f;
          else if (has_color && p.name == "green") out.vertex_colors[3 * i + 1] = (float)(unsigned char)x / 255.f;
          else if (has_color && p.name == "blue") out.vertex_colors[3 * i + 2] = (float)(unsigned char)x / 255.f;
        }
      }
    }
  }
  bounds_of(out);
}

PlyReader::PlyReader(const std::string rootdir, bool) {
  if (!path_exists(rootdir)) throw std::runtime_error("File \"" + rootdir + "\" does not exist.");
  if (!path_isdir(rootdir)) throw std::runtime_error("File \"" + rootdir + "\" is not a directory.");
  std::vector<std::string> files = find_ply(rootdir);
  if (files.empty()) throw std::runtime_error("Directory \"" + rootdir + "\" contains no .ply files.");
  int k = 0;
  for (const std::string &file : files) {
    std::string meshname = "Mesh" + std::to_string(k++);
    MeshData m;
    readFile(file, m);
    api::createMesh(meshname);
    api::addMeshVertices(meshname, (unsigned)(m.vertices.size() / 3), m.vertices.data());
    std::vector<unsigned> one(m.faces.size());
    for (size_t i = 0; i < one.size(); ++i) one[i] = m.faces[i] + 1; // api::addMeshTriangles is 1-based (Mesh.cpp:102-112)
    api::addMeshTriangles(meshname, (unsigned)(one.size() / 3), one.data());
    if (!m.vertex_colors.empty()) api::addMeshVertexColor(meshname, (unsigned)(m.vertex_colors.size() / 3), m.vertex_colors.data());
    api::finishMesh(meshname);
    meshes.push_back(meshname);
  }
}
PlyReader::~PlyReader() {}

void PlyReader::addInstances() const {
  const float identity[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
  for (size_t k = 0; k < meshes.size(); ++k) api::addInstance("inst" + std::to_string(k), meshes[k], identity);
}

} // namespace reader
} // namespace domain
} // namespace data
} // namespace render
} // namespace gvt

namespace gvtapps {
namespace render {

namespace {
std::vector<std::string> split(const std::string &s, char delim) { // ConfigFileLoader.cpp:51-60: empty items dropped
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, delim))
    if (!item.empty()) out.push_back(item);
  return out;
}
float num(const std::vector<std::string> &e, size_t i) {
  if (i >= e.size()) throw std::runtime_error("conf: line '" + e[0] + "' has too few fields");
  return (float)std::atof(e[i].c_str());
}
} // namespace

ConfigFileLoader::ConfigFileLoader(const std::string filename) {
  if (filename.empty()) throw std::runtime_error("Error filename not specified");
  std::ifstream file(filename.c_str());
  if (!file.good()) throw std::runtime_error("Error loading config file " + filename);
  std::size_t slash = filename.find_last_of("/");
  const std::string confdir = filename.substr(0, slash + 1);

  std::map<std::string, std::string> mesh_of_file; // one mesh (and one adapter / BLAS) per distinct OBJ path (:139-166)
  float eye[3] = { 0, 0, 1 }, focus[3] = { 0, 0, 0 }, up[3] = { 0, 1, 0 }, fov = (float)(45.0 * M_PI / 180.0);
  bool have_camera = false;
  std::string line;
  while (std::getline(file, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.find("#") == 0) continue;
    for (char &c : line)
      if (c == '\t') c = ' ';
    std::vector<std::string> e = split(line, ' ');
    if (e.empty()) continue;
    if (e[0] == "F") {
      width = std::atoi(e.at(1).c_str()), height = std::atoi(e.at(2).c_str());
    } else if (e[0] == "C") {
      for (int i = 0; i < 3; ++i) eye[i] = num(e, 1 + i), focus[i] = num(e, 4 + i), up[i] = num(e, 7 + i);
      // the file's angle is honoured; the (compiled-out) reference parses it and then stores 45 degrees "TODO" (:107,123)
      fov = (float)(std::atof(e.at(10).c_str()) * M_PI / 180.0);
      have_camera = true;
    } else if (e[0] == "G") {
      if (e.size() < 11) throw std::runtime_error("conf: G needs a file and 9 numbers");
      if (e[1].find(".obj") >= e[1].size()) throw std::runtime_error("conf: G " + e[1] + ": only .obj datasets are loadable (:205-207)");
      std::string path = e[1];
      if (path[0] != '/' && !path_exists(path)) path = confdir + path; // relative to the conf file when not to the cwd
      auto it = mesh_of_file.find(path);
      if (it == mesh_of_file.end()) {
        const std::string meshname = "Mesh" + std::to_string(mesh_of_file.size());
        gvt::render::data::domain::reader::ObjReader reader(path);
        reader.addToScene(meshname);
        it = mesh_of_file.insert(std::make_pair(path, meshname)).first;
      }
      // glm::translate then glm::scale on the identity (:170-203); the rotation triple is parsed and has no effect,
      // as in the reference (mAA is computed and dropped, :184-191)
      float m[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
      const float t[3] = { num(e, 2), num(e, 3), num(e, 4) }, s[3] = { num(e, 8), num(e, 9), num(e, 10) };
      if (std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) > 0.0f) m[12] = t[0], m[13] = t[1], m[14] = t[2];
      if (std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) > 0.0f) m[0] *= s[0], m[5] *= s[1], m[10] *= s[2];
      api::addInstance("inst" + std::to_string(n_instances++), it->second, m);
    } else if (e[0] == "LP") {
      const float pos[3] = { num(e, 1), num(e, 2), num(e, 3) }, color[3] = { num(e, 4), num(e, 5), num(e, 6) };
      api::addPointLight("PointLight" + std::to_string(n_lights++), pos, color);
    } else if (e[0] == "LA") {
      const float pos[3] = { num(e, 1), num(e, 2), num(e, 3) }, n[3] = { num(e, 4), num(e, 5), num(e, 6) };
      const float w = num(e, 7), h = num(e, 8), color[3] = { num(e, 9), num(e, 10), num(e, 11) };
      api::addAreaLight("AreaLight" + std::to_string(n_lights++), pos, color, n, w, h);
    } else if (e[0] == "DT" && e.size() > 1) {
      if (e[1] == "OPTIX") domain_type = 1;
      if (e[1] == "EMBREE") domain_type = 2;
      if (e[1] == "B200") domain_type = 3;
    } else if (e[0] == "ST" && e.size() > 1) {
      if (e[1] == "DOMAIN") scheduler_type = 1;
      if (e[1] == "HYBRID") scheduler_type = 2;
    } else if (e[0] == "AT" && e.size() > 1) {
      if (e[1] == "BVH") accel_type = BVH;
    }
  }
  if (have_camera) api::addCamera("Camera", eye, focus, up, fov, 1, 1, 0.f);
  if (width > 0 && height > 0) api::addFilm("Film", width, height, "output");
  if (have_camera && width > 0 && height > 0)
    api::addRenderer("Renderer", gvt::render::adapter::B200,
                     scheduler_type == 1 ? gvt::render::scheduler::Domain : gvt::render::scheduler::Image, "Camera", "Film");
}
ConfigFileLoader::~ConfigFileLoader() {}

} // namespace render
} // namespace gvtapps

// ---- C entry points (ctypes / FFI): 0 = ok, -1 = exception, text via gvtapi_reader_last_error -------------------------
namespace {
thread_local std::string g_reader_err;
template <class F> int guard(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_reader_err = e.what();
    return -1;
  }
}
using gvt::render::data::domain::reader::MeshData;
template <class T> void copy_out(const std::vector<T> &v, T *dst) {
  if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(T));
}
} // namespace
#define GAPI extern "C" __attribute__((visibility("default")))
GAPI const char *gvtapi_reader_last_error() { return g_reader_err.c_str(); }

// Two-call protocol: pass null arrays to get the sizes (n[0] vertices, n[1] faces, n[2] materials, n[3] normals,
// n[4] vertex colours), then arrays of that size.  materials: 10 floats each (ka, kd, ks, type).
static int mesh_out(const MeshData &m, unsigned n[5], float *vertices, unsigned *faces, int *face_mat, float *materials, float *normals,
                    float *colors, float *bounds) {
  if (n) {
    n[0] = (unsigned)(m.vertices.size() / 3), n[1] = (unsigned)(m.faces.size() / 3), n[2] = (unsigned)m.materials.size();
    n[3] = (unsigned)(m.normals.size() / 3), n[4] = (unsigned)(m.vertex_colors.size() / 3);
  }
  copy_out(m.vertices, vertices), copy_out(m.faces, faces), copy_out(m.faces_to_materials, face_mat);
  copy_out(m.normals, normals), copy_out(m.vertex_colors, colors);
  if (materials)
    for (size_t i = 0; i < m.materials.size(); ++i) {
      memcpy(materials + 10 * i, m.materials[i].ka, 12), memcpy(materials + 10 * i + 3, m.materials[i].kd, 12);
      memcpy(materials + 10 * i + 6, m.materials[i].ks, 12), materials[10 * i + 9] = (float)m.materials[i].type;
    }
  if (bounds) memcpy(bounds, m.bounds_min, 12), memcpy(bounds + 3, m.bounds_max, 12);
  return 0;
}
GAPI int gvtapi_obj_read(const char *path, int material_type, unsigned n[5], float *vertices, unsigned *faces, int *face_mat,
                         float *materials, float *normals, float *bounds) {
  return guard([&] {
    gvt::render::data::domain::reader::ObjReader r(path, material_type);
    mesh_out(*r.getMesh(), n, vertices, faces, face_mat, materials, normals, nullptr, bounds);
  });
}
GAPI int gvtapi_ply_read(const char *path, unsigned n[5], float *vertices, unsigned *faces, float *colors, float *bounds) {
  return guard([&] {
    MeshData m;
    gvt::render::data::domain::reader::PlyReader::readFile(path, m);
    mesh_out(m, n, vertices, faces, nullptr, nullptr, nullptr, colors, bounds);
  });
}
GAPI int gvtapi_parse_float(const char *text, float *out) { // the OBJ number grammar (parity hook)
  const char *tok = text;
  *out = next_float(tok);
  return 0;
}
GAPI int gvtapi_load_obj(const char *path, int material_type, const char *meshname) {
  return guard([&] { gvt::render::data::domain::reader::ObjReader(path, material_type).addToScene(meshname); });
}
GAPI int gvtapi_load_ply_dir(const char *dir, int add_instances, int *n_meshes) {
  return guard([&] {
    gvt::render::data::domain::reader::PlyReader r(dir);
    if (add_instances) r.addInstances();
    if (n_meshes) *n_meshes = (int)r.getMeshes().size();
  });
}
// info: scheduler_type, domain_type, accel_type, width, height, n_instances, n_lights
GAPI int gvtapi_load_conf(const char *path, int info[7]) {
  return guard([&] {
    gvtapps::render::ConfigFileLoader c(path);
    if (info) {
      info[0] = c.scheduler_type, info[1] = c.domain_type, info[2] = (int)c.accel_type, info[3] = c.width, info[4] = c.height;
      info[5] = c.n_instances, info[6] = c.n_lights;
    }
  });
}
