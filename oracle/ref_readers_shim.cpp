// ref_readers_shim.cpp -- C entry points over the REAL third-party readers the reference links (tinyobjloader and Greg
// Turk's ply.c, compiled where they lie under /root/reference/third-party by oracle/Makefile into
// oracle/_ref/libgvt_ref_readers.so).  TEST INFRASTRUCTURE ONLY: validates gravit_b200/host/readers.cpp.  This file is
// ours; no reference source is copied into this repo.
//
// gvtref_obj_* flatten tinyobj's shapes exactly like ObjReader.cpp:171-203 (vertices in shape order, faces offset by the
// vertices of the earlier shapes, material id per face).  gvtref_ply_* read vertex x/y/z (+ red/green/blue when the
// vertex element has more than five properties) and the face list like PlyReader.cpp:100-132.
#include <tiny_obj_loader.h>

#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
#include <ply.h>
}

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
struct Flat {
  std::vector<float> vertices, normals, materials; // materials: ka kd ks (9 floats)
  std::vector<unsigned> faces;
  std::vector<int> face_mat;
  std::vector<float> colors;
};
Flat g;
} // namespace

// n: vertices, faces, materials, normal vectors; returns 0 / -1
REF_API int gvtref_obj_load(const char *filename, unsigned n[4]) {
  g = Flat();
  std::string fn(filename);
  std::size_t found = fn.find_last_of("/");
  std::string path = fn.substr(0, found + 1);
  std::vector<tinyobj::shape_t> shapes;
  std::vector<tinyobj::material_t> materials;
  std::string err;
  if (!tinyobj::LoadObj(shapes, materials, err, filename, path.c_str())) return -1;
  for (const auto &m : materials) {
    g.materials.insert(g.materials.end(), m.ambient, m.ambient + 3);
    g.materials.insert(g.materials.end(), m.diffuse, m.diffuse + 3);
    g.materials.insert(g.materials.end(), m.specular, m.specular + 3);
  }
  size_t offset = 0;
  for (size_t i = 0; i < shapes.size(); i++) {
    const tinyobj::mesh_t &m = shapes[i].mesh;
    g.vertices.insert(g.vertices.end(), m.positions.begin(), m.positions.end());
    g.normals.insert(g.normals.end(), m.normals.begin(), m.normals.end());
    for (size_t f = 0; f < m.indices.size() / 3; f++) {
      for (int k = 0; k < 3; ++k) g.faces.push_back((unsigned)(offset + m.indices[3 * f + k]));
      g.face_mat.push_back(m.material_ids[f]);
    }
    offset += m.positions.size() / 3;
  }
  n[0] = (unsigned)(g.vertices.size() / 3), n[1] = (unsigned)(g.faces.size() / 3), n[2] = (unsigned)materials.size();
  n[3] = (unsigned)(g.normals.size() / 3);
  return 0;
}
REF_API void gvtref_obj_get(float *vertices, unsigned *faces, int *face_mat, float *materials, float *normals) {
  if (vertices) memcpy(vertices, g.vertices.data(), g.vertices.size() * 4);
  if (faces) memcpy(faces, g.faces.data(), g.faces.size() * 4);
  if (face_mat) memcpy(face_mat, g.face_mat.data(), g.face_mat.size() * 4);
  if (materials) memcpy(materials, g.materials.data(), g.materials.size() * 4);
  if (normals) memcpy(normals, g.normals.data(), g.normals.size() * 4);
}

namespace {
struct PVertex {
  float x, y, z;
  unsigned char cx, cy, cz;
  void *other_props;
};
struct PFace {
  unsigned char nverts;
  int *verts;
  void *other_props;
};
PlyProperty vert_props[] = {
  { (char *)"x", Float32, Float32, offsetof(PVertex, x), 0, 0, 0, 0 },   { (char *)"y", Float32, Float32, offsetof(PVertex, y), 0, 0, 0, 0 },
  { (char *)"z", Float32, Float32, offsetof(PVertex, z), 0, 0, 0, 0 },   { (char *)"red", Uint8, Uint8, offsetof(PVertex, cx), 0, 0, 0, 0 },
  { (char *)"green", Uint8, Uint8, offsetof(PVertex, cy), 0, 0, 0, 0 }, { (char *)"blue", Uint8, Uint8, offsetof(PVertex, cz), 0, 0, 0, 0 },
};
PlyProperty face_props[] = {
  { (char *)"vertex_indices", Int32, Int32, offsetof(PFace, verts), 1, Uint8, Uint8, offsetof(PFace, nverts) },
};
} // namespace

// n: vertices, faces, has_color
REF_API int gvtref_ply_load(const char *filename, unsigned n[3]) {
  g = Flat();
  FILE *fp = fopen(filename, "r");
  if (!fp) return -1;
  PlyFile *in = read_ply(fp);
  if (!in) return -1;
  bool has_color = false;
  for (int i = 0; i < in->num_elem_types; i++) {
    int count = 0;
    std::string name = setup_element_read_ply(in, i, &count);
    if (name == "vertex") {
      setup_property_ply(in, &vert_props[0]);
      setup_property_ply(in, &vert_props[1]);
      setup_property_ply(in, &vert_props[2]);
      if (in->elems[i]->nprops > 5) {
        has_color = true;
        setup_property_ply(in, &vert_props[3]);
        setup_property_ply(in, &vert_props[4]);
        setup_property_ply(in, &vert_props[5]);
      }
      for (int j = 0; j < count; j++) {
        PVertex v;
        memset(&v, 0, sizeof v);
        get_element_ply(in, (void *)&v);
        g.vertices.push_back(v.x), g.vertices.push_back(v.y), g.vertices.push_back(v.z);
        if (has_color) g.colors.push_back(float(v.cx) / 255.f), g.colors.push_back(float(v.cy) / 255.f), g.colors.push_back(float(v.cz) / 255.f);
      }
    } else if (name == "face") {
      setup_property_ply(in, &face_props[0]);
      for (int j = 0; j < count; j++) {
        PFace f;
        memset(&f, 0, sizeof f);
        get_element_ply(in, (void *)&f);
        if (f.nverts >= 3 && f.verts) g.faces.push_back(f.verts[0]), g.faces.push_back(f.verts[1]), g.faces.push_back(f.verts[2]);
      }
    } else {
      // other elements are skipped by reading them with no properties selected
      for (int j = 0; j < count; j++) get_element_ply(in, nullptr);
    }
  }
  close_ply(in);
  n[0] = (unsigned)(g.vertices.size() / 3), n[1] = (unsigned)(g.faces.size() / 3), n[2] = has_color ? 1u : 0u;
  return 0;
}
REF_API void gvtref_ply_get(float *vertices, unsigned *faces, float *colors) {
  if (vertices) memcpy(vertices, g.vertices.data(), g.vertices.size() * 4);
  if (faces) memcpy(faces, g.faces.data(), g.faces.size() * 4);
  if (colors) memcpy(colors, g.colors.data(), g.colors.size() * 4);
}
