// ref_shim.cpp -- C entry points over the REAL reference code, for validating the oracle restatement.
// TEST INFRASTRUCTURE ONLY.  This file is ours; it #includes the reference's headers from where they lie
// (/root/reference/src, /root/reference/third-party/glm) and is linked with the reference's own
// Ray.cpp, BBox.cpp, Light.cpp, Mesh.cpp, Material.cpp compiled in place by oracle/Makefile into
// oracle/_ref/libgvt_ref.so.  No reference source is copied into this repo.
//
// What the reference can give us standalone: Shade(), Box3D, RayPacketIntersection<8>,
// Mesh::addFace/generateNormals/computeBoundingBox, PointLight::contribution, the Ray record, GLM's
// inverse/transpose.  BVH.cpp / gvtCamera.cpp / the tracers need MPI+TBB and Embree is absent.
#include <gvt/render/actor/Ray.h>
#include <gvt/render/actor/RayPacket.h>
#include <gvt/render/data/primitives/BBox.h>
#include <gvt/render/data/primitives/Material.h>
#include <gvt/render/data/primitives/Mesh.h>
#include <gvt/render/data/scene/Light.h>

#include <cstring>

using namespace gvt::render::actor;
using namespace gvt::render::data::primitives;
using namespace gvt::render::data::scene;

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API int gvtref_sizeof_ray() { return (int)sizeof(Ray); }
REF_API int gvtref_sizeof_box() { return (int)sizeof(Box3D); }
REF_API int gvtref_vec3_length() { return (int)glm::vec3().length(); }

// Shade (Material.cpp:90-139) with a PointLight (Light.cpp:58-62)
REF_API int gvtref_shade(int mtype, const float *kd, const float *ks, float alpha, const void *ray80, const float *N,
                         const float *lpos, const float *lcolor, float *out) {
  Material m;
  m.type = mtype;
  m.kd = glm::vec3(kd[0], kd[1], kd[2]);
  m.ks = glm::vec3(ks[0], ks[1], ks[2]);
  m.alpha = alpha;
  Ray r((const unsigned char *)ray80);
  PointLight L(glm::vec3(lpos[0], lpos[1], lpos[2]), glm::vec3(lcolor[0], lcolor[1], lcolor[2]));
  glm::vec3 c(0.f);
  bool ok = Shade(&m, r, glm::vec3(N[0], N[1], N[2]), &L, L.position, c);
  out[0] = c[0], out[1] = c[1], out[2] = c[2];
  return ok ? 1 : 0;
}

// RayPacketIntersection<8>::intersect (RayPacket.h:111-211) on up to 8 rays; returns hit[] and the updated t[]
REF_API void gvtref_packet_box(const void *rays80, int n, const float *box6, int update, int *hit, float *t) {
  RayVector rv(n);
  for (int i = 0; i < n; ++i) rv[i] = Ray((const unsigned char *)rays80 + 80 * i);
  RayPacketIntersection<8> rp(rv.begin(), rv.end());
  Box3D bb(glm::vec3(box6[0], box6[1], box6[2]), glm::vec3(box6[3], box6[4], box6[5]));
  rp.intersect(bb, hit, update != 0);
  for (int i = 0; i < 8; ++i) t[i] = rp.t[i];
}

// Mesh::addFace filter + generateNormals + computeBoundingBox (Mesh.cpp:102-154,186-199)
REF_API int gvtref_mesh(const float *verts, int nv, const unsigned *tris1, int nt, int *faces_out, float *normals_out,
                        float *bbox_out) {
  Mesh m;
  for (int i = 0; i < nv; ++i) m.addVertex(glm::vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
  for (int i = 0; i < nt; ++i) m.addFace(tris1[3 * i], tris1[3 * i + 1], tris1[3 * i + 2]);
  m.computeBoundingBox();
  m.generateNormals();
  int nf = (int)m.faces.size();
  for (int i = 0; i < nf; ++i) {
    faces_out[3 * i] = std::get<0>(m.faces[i]);
    faces_out[3 * i + 1] = std::get<1>(m.faces[i]);
    faces_out[3 * i + 2] = std::get<2>(m.faces[i]);
  }
  for (int i = 0; i < nv; ++i)
    for (int k = 0; k < 3; ++k) normals_out[3 * i + k] = m.normals[i][k];
  for (int k = 0; k < 3; ++k) bbox_out[k] = m.boundingBox.bounds_min[k], bbox_out[3 + k] = m.boundingBox.bounds_max[k];
  return nf;
}

// api::addInstance arithmetic (api.cpp:300-321): minv, normi, instance box from two corners, centroid
REF_API void gvtref_instance(const float *am, const float *meshbox6, float *minv16, float *normi9, float *ibox6,
                             float *centroid3) {
  glm::mat4 m = glm::make_mat4(am);
  glm::mat4 minv = glm::inverse(m);
  glm::mat3 normi = glm::transpose(glm::inverse(glm::mat3(m)));
  auto il = glm::vec3(m * glm::vec4(meshbox6[0], meshbox6[1], meshbox6[2], 1.f));
  auto ih = glm::vec3(m * glm::vec4(meshbox6[3], meshbox6[4], meshbox6[5], 1.f));
  Box3D ibox(il, ih);
  memcpy(minv16, glm::value_ptr(minv), 64);
  memcpy(normi9, glm::value_ptr(normi), 36);
  for (int k = 0; k < 3; ++k) ibox6[k] = ibox.bounds_min[k], ibox6[3 + k] = ibox.bounds_max[k];
  glm::vec3 c = ibox.centroid();
  for (int k = 0; k < 3; ++k) centroid3[k] = c[k];
}

// Box3D::merge / surfaceArea / wideRangingBoxDir (BBox.cpp:55-63,118-133), used by the top-level BVH build
REF_API float gvtref_box_merge_area(const float *a6, const float *b6, float *out6, int *widest) {
  Box3D a(glm::vec3(a6[0], a6[1], a6[2]), glm::vec3(a6[3], a6[4], a6[5]));
  Box3D b(glm::vec3(b6[0], b6[1], b6[2]), glm::vec3(b6[3], b6[4], b6[5]));
  a.merge(b);
  for (int k = 0; k < 3; ++k) out6[k] = a.bounds_min[k], out6[3 + k] = a.bounds_max[k];
  *widest = a.wideRangingBoxDir();
  return a.surfaceArea();
}

// glm::normalize(mat3 * v) as used for normals (EmbreeMeshAdapter.cpp:504,518)
REF_API void gvtref_normal(const float *normi9, const float *v3in, float *out3) {
  glm::mat3 n = glm::make_mat3(normi9);
  glm::vec3 r = glm::normalize(n * glm::vec3(v3in[0], v3in[1], v3in[2]));
  out3[0] = r[0], out3[1] = r[1], out3[2] = r[2];
}
