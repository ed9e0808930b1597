// gvt_oracle.cpp -- CPU ORACLE for the GraviT ray-trace hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is a strict-IEEE (no fast-math, -ffp-contract=off, explicit fmaf only where written)
// restatement of the reference's algorithm for the path
//     camera rays -> top-level instance routing -> per-instance trace/shade/shadow -> framebuffer
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load it.
// The product (gravit_b200/csrc, gravit_b200/host) never links, imports or calls anything in oracle/.
//
// Parity pinning: the whole pipeline reproduces the reference's own golden images
// Test/CTESTtest/data/simple.ppm and bunny.ppm (smooth-normal mode) within the reference's CTest
// tolerance (sum-abs-diff < 300, CMakeLists.txt:666); see tests/test_oracle_golden.py.  The pieces that
// compile from the reference's own sources (Shade, Box3D, RayPacketIntersection, Mesh::generateNormals,
// PointLight::contribution, GLM inverse) are cross-checked against oracle/_ref (built by oracle/Makefile).
// The ray/triangle arithmetic lives in Embree 2.15 (third-party/embree @ c9e873db, absent): per-ray
// t/u/v/primID are "parity unpinned" by the reference and are pinned by the goldens at image level only.
//
// Every function cites the reference file:line it follows (paths relative to src/gvt/render/).
// Vector math restates GLM 0.9.8.1 operation order (third-party/glm/glm/detail/*.inl).
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define OR_API extern "C" __attribute__((visibility("default")))

namespace {

// ---------------------------------------------------------------------------------------------
// GLM-order vector math
// ---------------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
  float &operator[](int i) { return (&x)[i]; }
  float operator[](int i) const { return (&x)[i]; }
};
inline V3 v3(float x, float y, float z) { return V3{ x, y, z }; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
// func_geometric.inl:54-60  tmp = x*y; tmp.x + tmp.y + tmp.z
inline float dot(V3 a, V3 b) {
  V3 t = a * b;
  return t.x + t.y + t.z;
}
// func_geometric.inl:76-84
inline V3 cross(V3 x, V3 y) { return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
// func_geometric.inl:94 v * inversesqrt(dot(v,v)); func_exponential.inl:46 1/sqrt(x)
inline V3 normalize(V3 v) { return v * (1.0f / sqrtf(dot(v, v))); }
inline float length(V3 v) { return sqrtf(dot(v, v)); }

// RayPacket.h:50,58  fastmin/fastmax: NaN-order-sensitive selects, NOT fminf/fmaxf
inline float fastmin(float a, float b) { return (a < b) ? a : b; }
inline float fastmax(float a, float b) { return (a > b) ? a : b; }

struct M4 {
  float c[4][4]; // c[col][row], GLM column-major
};
struct M3 {
  float c[3][3];
};

// type_mat4x4.inl:532-543  (m0*x + m1*y) + (m2*z + m3*w)
inline void m4_mul_v4(const M4 &m, const float v[4], float out[4]) {
  for (int r = 0; r < 4; ++r) {
    float mul0 = m.c[0][r] * v[0], mul1 = m.c[1][r] * v[1];
    float add0 = mul0 + mul1;
    float mul2 = m.c[2][r] * v[2], mul3 = m.c[3][r] * v[3];
    float add1 = mul2 + mul3;
    out[r] = add0 + add1;
  }
}
// type_mat3x3.inl:436-439
inline V3 m3_mul_v3(const M3 &m, V3 v) {
  return v3(m.c[0][0] * v.x + m.c[1][0] * v.y + m.c[2][0] * v.z, m.c[0][1] * v.x + m.c[1][1] * v.y + m.c[2][1] * v.z,
            m.c[0][2] * v.x + m.c[1][2] * v.y + m.c[2][2] * v.z);
}
// func_matrix.inl:272-293
inline M3 m3_inverse(const M3 &mm) {
  const float(*m)[3] = mm.c;
  float ood = 1.0f / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) +
                      m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
  M3 I;
  I.c[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * ood;
  I.c[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * ood;
  I.c[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * ood;
  I.c[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * ood;
  I.c[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * ood;
  I.c[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * ood;
  I.c[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * ood;
  I.c[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * ood;
  I.c[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * ood;
  return I;
}
inline M3 m3_transpose(const M3 &m) {
  M3 t;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t.c[i][j] = m.c[j][i];
  return t;
}
// func_matrix.inl:297-354
inline M4 m4_inverse(const M4 &mm) {
  const float(*m)[4] = mm.c;
  float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
  float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
  float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
  float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
  float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
  float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
  float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
  float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
  float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
  float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
  float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
  float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
  float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
  float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
  float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
  float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
  float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
  float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
  float Fac0[4] = { Coef00, Coef00, Coef02, Coef03 }, Fac1[4] = { Coef04, Coef04, Coef06, Coef07 };
  float Fac2[4] = { Coef08, Coef08, Coef10, Coef11 }, Fac3[4] = { Coef12, Coef12, Coef14, Coef15 };
  float Fac4[4] = { Coef16, Coef16, Coef18, Coef19 }, Fac5[4] = { Coef20, Coef20, Coef22, Coef23 };
  float Vec0[4] = { m[1][0], m[0][0], m[0][0], m[0][0] }, Vec1[4] = { m[1][1], m[0][1], m[0][1], m[0][1] };
  float Vec2[4] = { m[1][2], m[0][2], m[0][2], m[0][2] }, Vec3[4] = { m[1][3], m[0][3], m[0][3], m[0][3] };
  const float SignA[4] = { +1, -1, +1, -1 }, SignB[4] = { -1, +1, -1, +1 };
  M4 Inv;
  for (int i = 0; i < 4; ++i) {
    float Inv0 = Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i] + Vec3[i] * Fac2[i];
    float Inv1 = Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i] + Vec3[i] * Fac4[i];
    float Inv2 = Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i] + Vec3[i] * Fac5[i];
    float Inv3 = Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i] + Vec2[i] * Fac5[i];
    Inv.c[0][i] = Inv0 * SignA[i];
    Inv.c[1][i] = Inv1 * SignB[i];
    Inv.c[2][i] = Inv2 * SignA[i];
    Inv.c[3][i] = Inv3 * SignB[i];
  }
  float Dot0[4] = { m[0][0] * Inv.c[0][0], m[0][1] * Inv.c[1][0], m[0][2] * Inv.c[2][0], m[0][3] * Inv.c[3][0] };
  float Dot1 = (Dot0[0] + Dot0[1]) + (Dot0[2] + Dot0[3]);
  float ood = 1.0f / Dot1;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) Inv.c[c][r] *= ood;
  return Inv;
}

// ---------------------------------------------------------------------------------------------
// Records shared with the tests (plain C layouts)
// ---------------------------------------------------------------------------------------------
// actor/Ray.h:68-95: 64-byte payload inside a 16-aligned 68-byte union => sizeof(Ray)==80
struct Ray80 {
  V3 o;
  float t_min;
  V3 d;
  float t_max;
  V3 color;
  float t;
  int id;
  int depth;
  float w;
  int type;
  unsigned char pad[16];
};
static_assert(sizeof(Ray80) == 80, "Ray record must be 80 bytes");
enum { RAY_PRIMARY = 0, RAY_SHADOW = 1, RAY_SECONDARY = 2 }; // actor/Ray.h:50-55
const float RAY_EPSILON = 1.e-6; // actor/Ray.cpp:33 (double literal rounded to float)

struct HitRec { // closest-hit record of the first query of each input ray (test hook)
  int prim;     // -1 = miss
  float t, u, v;
};

struct Box {
  V3 lo, hi;
  Box() : lo(v3(FLT_MAX, FLT_MAX, FLT_MAX)), hi(v3(-FLT_MAX, -FLT_MAX, -FLT_MAX)) {} // BBox.cpp:33-38
  void merge(const Box &o) {                                                         // BBox.cpp:55-63
    lo.x = fastmin(o.lo.x, lo.x), lo.y = fastmin(o.lo.y, lo.y), lo.z = fastmin(o.lo.z, lo.z);
    hi.x = fastmax(o.hi.x, hi.x), hi.y = fastmax(o.hi.y, hi.y), hi.z = fastmax(o.hi.z, hi.z);
  }
  void expand(V3 v) { // BBox.cpp:65-73
    lo.x = fastmin(lo.x, v.x), lo.y = fastmin(lo.y, v.y), lo.z = fastmin(lo.z, v.z);
    hi.x = fastmax(hi.x, v.x), hi.y = fastmax(hi.y, v.y), hi.z = fastmax(hi.z, v.z);
  }
  int wideRangingBoxDir() const { // BBox.cpp:118-126
    V3 d = hi - lo;
    if (d.x > d.y && d.x > d.z) return 0;
    if (d.y > d.z) return 1;
    return 2;
  }
  V3 centroid() const { return lo * 0.5f + hi * 0.5f; } // BBox.cpp:128 (0.5f*min + 0.5f*max)
  float surfaceArea() const {                           // BBox.cpp:130-133
    V3 d = hi - lo;
    return 2.f * (d.x * d.y + d.y * d.z + d.z * d.x);
  }
};

struct Material { // data/primitives/Material.h:59-89 (fields used by LAMBERT/PHONG/BLINN)
  int type;
  V3 kd, ks;
  float alpha;
};
enum { MAT_LAMBERT = 0, MAT_PHONG = 1, MAT_BLINN = 2 };

struct Light { // data/scene/Light.h  (type 0 point, 1 ambient, 2 area)
  int type;
  V3 pos, color;
  V3 normal;
  float width, height;
};

// ---------------------------------------------------------------------------------------------
// Mesh + its CPU BVH (the BVH is the oracle's own; closest hit is BVH-independent by the tie rule)
// ---------------------------------------------------------------------------------------------
struct BNode {
  Box box;
  int left, right; // children, or -1
  int first, count;
};

struct Mesh {
  std::vector<V3> verts;
  std::vector<int> faces; // 3 per face, 0-based, post-filter (Mesh.cpp:102-112)
  std::vector<V3> normals;
  std::vector<int> face_mat; // optional per-face material index (or empty)
  std::vector<Material> mats;
  std::vector<V3> vcolors;
  Material mat;
  Box bbox;
  // bvh
  std::vector<BNode> nodes;
  std::vector<int> prim_order;

  // Mesh.cpp:116-154
  void generateNormals() {
    normals.assign(verts.size(), v3(0, 0, 0));
    size_t nf = faces.size() / 3;
    for (size_t i = 0; i < nf; ++i) {
      int I = faces[3 * i], J = faces[3 * i + 1], K = faces[3 * i + 2];
      V3 a = verts[I], b = verts[J], c = verts[K];
      V3 u = b - a, v = c - a, n;
      n.x = u.y * v.z - u.z * v.y;
      n.y = u.z * v.x - u.x * v.z;
      n.z = u.x * v.y - u.y * v.x;
      n = normalize(n);
      normals[I] = normals[I] + n;
      normals[J] = normals[J] + n;
      normals[K] = normals[K] + n;
    }
    for (auto &n : normals) n = normalize(n);
  }
  void computeBoundingBox() { // Mesh.cpp:186-199
    Box b;
    for (auto &p : verts) b.expand(p);
    bbox = b;
  }

  Box triBox(int f) const {
    Box b;
    b.expand(verts[faces[3 * f]]);
    b.expand(verts[faces[3 * f + 1]]);
    b.expand(verts[faces[3 * f + 2]]);
    return b;
  }

  int buildNode(int first, int count, std::vector<Box> &tb, std::vector<V3> &tc) {
    int id = (int)nodes.size();
    nodes.push_back(BNode());
    Box b, cb;
    for (int i = first; i < first + count; ++i) {
      b.merge(tb[prim_order[i]]);
      cb.expand(tc[prim_order[i]]);
    }
    nodes[id].box = b;
    nodes[id].left = nodes[id].right = -1;
    nodes[id].first = first;
    nodes[id].count = count;
    if (count <= 4) return id;
    // binned SAH, 16 bins, on the widest centroid axis
    V3 ext = cb.hi - cb.lo;
    int ax = (ext.x > ext.y && ext.x > ext.z) ? 0 : (ext.y > ext.z ? 1 : 2);
    int mid = -1;
    if (ext[ax] > 0) {
      const int NB = 16;
      Box bb[NB];
      int bc[NB] = { 0 };
      float k = NB * (1.f - 1e-6f) / ext[ax];
      for (int i = first; i < first + count; ++i) {
        int p = prim_order[i];
        int bi = std::min(NB - 1, std::max(0, (int)((tc[p][ax] - cb.lo[ax]) * k)));
        bb[bi].merge(tb[p]);
        bc[bi]++;
      }
      float la[NB], ra[NB];
      int lc[NB], rc[NB];
      Box acc;
      int c = 0;
      for (int i = 0; i < NB; ++i) {
        acc.merge(bb[i]);
        c += bc[i];
        la[i] = c ? acc.surfaceArea() : 0, lc[i] = c;
      }
      acc = Box();
      c = 0;
      for (int i = NB - 1; i >= 0; --i) {
        acc.merge(bb[i]);
        c += bc[i];
        ra[i] = c ? acc.surfaceArea() : 0, rc[i] = c;
      }
      float best = FLT_MAX;
      int bs = -1;
      for (int i = 0; i < NB - 1; ++i) {
        if (lc[i] == 0 || rc[i + 1] == 0) continue;
        float cost = la[i] * lc[i] + ra[i + 1] * rc[i + 1];
        if (cost < best) best = cost, bs = i;
      }
      if (bs >= 0) {
        auto it = std::partition(prim_order.begin() + first, prim_order.begin() + first + count, [&](int p) {
          int bi = std::min(NB - 1, std::max(0, (int)((tc[p][ax] - cb.lo[ax]) * k)));
          return bi <= bs;
        });
        mid = (int)(it - prim_order.begin());
      }
    }
    if (mid <= first || mid >= first + count) mid = first + count / 2; // median fallback
    int l = buildNode(first, mid - first, tb, tc);
    int r = buildNode(mid, first + count - mid, tb, tc);
    nodes[id].left = l;
    nodes[id].right = r;
    return id;
  }
  void buildBVH() {
    int nf = (int)(faces.size() / 3);
    nodes.clear();
    prim_order.resize(nf);
    std::vector<Box> tb(nf);
    std::vector<V3> tc(nf);
    for (int i = 0; i < nf; ++i) {
      prim_order[i] = i;
      tb[i] = triBox(i);
      tc[i] = tb[i].lo * 0.5f + tb[i].hi * 0.5f;
    }
    nodes.reserve(nf ? 2 * nf : 1);
    if (nf) buildNode(0, nf, tb, tc);
  }
};

// Canonical ray/triangle test (Moeller-Trumbore, fp32, the FMAs below are the ONLY fused ops; the CUDA
// kernel issues the same sequence with __fmaf_rn/__fmul_rn/... so t,u,v agree bit for bit).
// Stands in for Embree 2.15 rtcIntersect8/rtcOccluded8 (EmbreeMeshAdapter.cpp:474,375); u weights v1,
// v weights v2 as Embree does (EmbreeMeshAdapter.cpp:555-558).  Valid hit: tnear < t < tfar.
inline bool triTest(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float tnear, float tfar, float &t, float &u, float &v) {
  // p = cross(d, e2)
  float px = fmaf(d.y, e2.z, -(d.z * e2.y));
  float py = fmaf(d.z, e2.x, -(d.x * e2.z));
  float pz = fmaf(d.x, e2.y, -(d.y * e2.x));
  float det = fmaf(e1.z, pz, fmaf(e1.y, py, e1.x * px));
  float inv = 1.0f / det;
  float sx = o.x - v0.x, sy = o.y - v0.y, sz = o.z - v0.z;
  float uu = fmaf(sz, pz, fmaf(sy, py, sx * px)) * inv;
  if (!(uu >= 0.f && uu <= 1.f)) return false;
  // q = cross(s, e1)
  float qx = fmaf(sy, e1.z, -(sz * e1.y));
  float qy = fmaf(sz, e1.x, -(sx * e1.z));
  float qz = fmaf(sx, e1.y, -(sy * e1.x));
  float vv = fmaf(d.z, qz, fmaf(d.y, qy, d.x * qx)) * inv;
  if (!(vv >= 0.f && (uu + vv) <= 1.f)) return false;
  float tt = fmaf(e2.z, qz, fmaf(e2.y, qy, e2.x * qx)) * inv;
  if (!(tt > tnear && tt < tfar)) return false;
  t = tt, u = uu, v = vv;
  return true;
}

struct LocalHit {
  int prim = -1;
  float t = FLT_MAX, u = 0, v = 0;
};

// conservative slab test for the oracle's own BVH (padded; exactness is not required here because the
// closest hit is decided by triTest + the (t, prim) tie rule only)
inline bool slab(const Box &b, V3 o, V3 id, float tmax) {
  float pad = 1e-5f;
  V3 ext = b.hi - b.lo;
  float ex = ext.x * pad + 1e-30f, ey = ext.y * pad + 1e-30f, ez = ext.z * pad + 1e-30f;
  float l, h, tn = 0.f, tf = tmax;
  l = (b.lo.x - ex - o.x) * id.x, h = (b.hi.x + ex - o.x) * id.x;
  if (l > h) std::swap(l, h);
  if (!(l != l)) tn = std::max(tn, l);
  if (!(h != h)) tf = std::min(tf, h);
  l = (b.lo.y - ey - o.y) * id.y, h = (b.hi.y + ey - o.y) * id.y;
  if (l > h) std::swap(l, h);
  if (!(l != l)) tn = std::max(tn, l);
  if (!(h != h)) tf = std::min(tf, h);
  l = (b.lo.z - ez - o.z) * id.z, h = (b.hi.z + ez - o.z) * id.z;
  if (l > h) std::swap(l, h);
  if (!(l != l)) tn = std::max(tn, l);
  if (!(h != h)) tf = std::min(tf, h);
  return tn <= tf * 1.00001f + 1e-30f;
}

inline void testPrim(const Mesh &m, int f, V3 o, V3 d, float tnear, float tfar, LocalHit &h) {
  V3 v0 = m.verts[m.faces[3 * f]], v1 = m.verts[m.faces[3 * f + 1]], v2 = m.verts[m.faces[3 * f + 2]];
  V3 e1 = v1 - v0, e2 = v2 - v0;
  float t, u, v;
  if (triTest(o, d, v0, e1, e2, tnear, tfar, t, u, v)) {
    // tie rule: smaller t wins; equal t -> smaller primitive id wins (BVH-order independent)
    if (t < h.t || (t == h.t && f < h.prim)) h.prim = f, h.t = t, h.u = u, h.v = v;
  }
}

LocalHit intersectClosest(const Mesh &m, V3 o, V3 d, float tnear, float tfar, bool brute) {
  LocalHit h;
  int nf = (int)(m.faces.size() / 3);
  if (brute || m.nodes.empty()) {
    for (int f = 0; f < nf; ++f) testPrim(m, f, o, d, tnear, tfar, h);
    return h;
  }
  V3 id = v3(1.f / d.x, 1.f / d.y, 1.f / d.z);
  int stack[128], sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const BNode &n = m.nodes[stack[--sp]];
    // prune with <= (ties must still be visited so that the prim-id tie rule can apply)
    if (!slab(n.box, o, id, std::min(tfar, h.t))) continue;
    if (n.left < 0) {
      for (int i = n.first; i < n.first + n.count; ++i) testPrim(m, m.prim_order[i], o, d, tnear, tfar, h);
    } else {
      stack[sp++] = n.right;
      stack[sp++] = n.left;
    }
  }
  return h;
}

bool intersectAny(const Mesh &m, V3 o, V3 d, float tnear, float tfar, bool brute) {
  int nf = (int)(m.faces.size() / 3);
  LocalHit h;
  if (brute || m.nodes.empty()) {
    for (int f = 0; f < nf && h.prim < 0; ++f) testPrim(m, f, o, d, tnear, tfar, h);
    return h.prim >= 0;
  }
  V3 id = v3(1.f / d.x, 1.f / d.y, 1.f / d.z);
  int stack[128], sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const BNode &n = m.nodes[stack[--sp]];
    if (!slab(n.box, o, id, tfar)) continue;
    if (n.left < 0) {
      for (int i = n.first; i < n.first + n.count; ++i) {
        testPrim(m, m.prim_order[i], o, d, tnear, tfar, h);
        if (h.prim >= 0) return true;
      }
    } else {
      stack[sp++] = n.right;
      stack[sp++] = n.left;
    }
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// Scene
// ---------------------------------------------------------------------------------------------
struct Instance {
  int mesh;
  M4 m, minv;
  M3 normi;
  Box box;
  V3 centroid;
};

// data/accel/BVH.h:137-150
struct TNode {
  Box bbox;
  int left = -1, right = -1;
  int instanceSetIdx = 0, numInstances = 0;
};

enum { FLAG_SMOOTH = 1, FLAG_ENTER_INSIDE = 2, FLAG_BRUTE = 4, FLAG_SHADOW_TO_LIGHT = 8 };

struct Stats {
  long long camera_rays, routed_initial, adapter_calls, rays_traced, primary_hits, shadow_generated, shadow_unoccluded,
      secondary_generated, fb_adds, rounds, rays_sent;
};

struct Scene {
  std::vector<std::unique_ptr<Mesh> > meshes;
  std::vector<Instance> inst;
  std::vector<Light> lights;
  V3 eye, focus, up;
  float fov = 0.785398f;
  int depth = 1, samples = 1;
  int jitterWindowSize = 0; // gvtCamera.h:136 -- an int
  int W = 512, H = 512;
  int flags = 0;
  unsigned seed = 0;
  // TLAS (reference data/accel/BVH)
  std::vector<TNode> tnodes;
  std::vector<int> sortedID; // instanceSetID
  std::vector<Box> sortedBB; // instanceSetBB
  std::vector<int> work;     // instanceSet being partitioned (ids)
  std::vector<int> owner;    // domain scheduler: instance -> rank
  std::vector<long long> route_log; // per exchange round: P*P matrix of rays sent [src*P+dst]
  std::vector<int> call_log;        // per adapter call: (rank, inst, queue size)
};

// ---- TLAS build: data/accel/BVH.cpp:44-216 --------------------------------------------------------
float findSplitPoint(Scene &s, int splitAxis, int start, int end) { // BVH.cpp:173-216
  float minCost = FLT_MAX;
  float splitPoint = 0;
  for (int i = start; i < end; ++i) {
    const Box &ref = s.inst[s.work[i]].box;
    for (int e = 0; e < 2; ++e) {
      float edge = (e == 0) ? ref.lo[splitAxis] : ref.hi[splitAxis];
      Box leftBox, rightBox;
      int leftCount = 0;
      for (int j = start; j < end; ++j) {
        const Box &bb = s.inst[s.work[j]].box;
        if (bb.centroid()[splitAxis] < edge) {
          ++leftCount;
          leftBox.merge(bb);
        } else
          rightBox.merge(bb);
      }
      int rightCount = end - start - leftCount;
      // TRAVERSAL_COST 0.5 is a double literal: the sum is evaluated in double then rounded to float
      float cost = (float)(0.5 + (double)(leftBox.surfaceArea() * leftCount) + (double)(rightBox.surfaceArea() * rightCount));
      if (cost < minCost) {
        minCost = cost;
        splitPoint = edge;
      }
    }
  }
  return splitPoint;
}

int tlasBuild(Scene &s, std::vector<int> &sorted, int start, int end) { // BVH.cpp:77-171
  int id = (int)s.tnodes.size();
  s.tnodes.push_back(TNode());
  Box bbox;
  for (int i = start; i < end; ++i) bbox.merge(s.inst[s.work[i]].box);
  int count = end - start;
  auto makeLeaf = [&]() {
    s.tnodes[id].bbox = bbox;
    s.tnodes[id].instanceSetIdx = (int)sorted.size();
    s.tnodes[id].numInstances = count;
    for (int i = start; i < end; ++i) sorted.push_back(s.work[i]);
  };
  if (count <= 1) { // LEAF_SIZE 1, BVH.cpp:40,95
    makeLeaf();
    return id;
  }
  int splitAxis = bbox.wideRangingBoxDir();
  float splitPoint = findSplitPoint(s, splitAxis, start, end);
  // BVH.cpp:139-140: std::partition (libstdc++) by centroid[axis] < splitPoint; the stored "centroid" child is
  // ibox->centroid() (api.cpp:321), CentroidLessThan reads it (BVH.h:160-172)
  auto bound = std::partition(s.work.begin() + start, s.work.begin() + end,
                              [&](int k) { return s.inst[k].centroid[splitAxis] < splitPoint; });
  int splitIdx = (int)(bound - s.work.begin());
  if (splitIdx == start || splitIdx == end) {
    makeLeaf();
    return id;
  }
  int l = tlasBuild(s, sorted, start, splitIdx);
  int r = tlasBuild(s, sorted, splitIdx, end);
  s.tnodes[id].left = l;
  s.tnodes[id].right = r;
  s.tnodes[id].bbox = bbox;
  s.tnodes[id].numInstances = 0;
  return id;
}

// ---- routing: actor/RayPacket.h:83-211 + data/accel/BVH.h:61-135 ------------------------------------
template <int W> struct Packet {
  float ox[W], oy[W], oz[W], dx[W], dy[W], dz[W], t[W];
  int mask[W];
  Packet(const Ray80 *r, size_t n) {
    size_t i = 0;
    for (; i < n && i < (size_t)W; ++i) {
      ox[i] = r[i].o.x, oy[i] = r[i].o.y, oz[i] = r[i].o.z;
      dx[i] = 1.f / r[i].d.x, dy[i] = 1.f / r[i].d.y, dz[i] = 1.f / r[i].d.z;
      t[i] = r[i].t_max;
      mask[i] = 1;
    }
    for (; i < (size_t)W; ++i) t[i] = -1, mask[i] = -1;
  }
  // enter_inside: documented deviation F6: a box whose slab interval contains the origin
  // (tnear <= eps < tfar) is entered at t = 0.  Only applied when the caller asks (camera rays).
  bool intersect(const Box &bb, int hit[], bool update, bool enter_inside) {
    bool any = false;
    for (int i = 0; i < W; ++i) {
      float lx = (bb.lo.x - ox[i]) * dx[i], ly = (bb.lo.y - oy[i]) * dy[i], lz = (bb.lo.z - oz[i]) * dz[i];
      float ux = (bb.hi.x - ox[i]) * dx[i], uy = (bb.hi.y - oy[i]) * dy[i], uz = (bb.hi.z - oz[i]) * dz[i];
      float minx = fastmin(lx, ux), maxx = fastmax(lx, ux);
      float miny = fastmin(ly, uy), maxy = fastmax(ly, uy);
      float minz = fastmin(lz, uz), maxz = fastmax(lz, uz);
      float tnear = fastmax(fastmax(minx, miny), minz);
      float tfar = fastmin(fastmin(maxx, maxy), maxz);
      bool inside = update && enter_inside && !(tnear > RAY_EPSILON) && tfar > tnear && tfar > RAY_EPSILON;
      if (inside) tnear = 0.f;
      bool ok = tfar > tnear && (!update || tnear > RAY_EPSILON || inside) && t[i] > tnear; // RayPacket.h:194
      hit[i] = ok ? 1 : -1;
      if (ok && update) t[i] = tnear;
      any |= ok;
    }
    return any;
  }
};

template <int W>
void routeT(const Scene &s, const Ray80 *rays, size_t n, int from, bool enter_inside, int *next, float *tout) {
  std::vector<int> stack(s.inst.size() * 2 + 2);
  for (size_t off = 0; off < n; off += W) {
    size_t cnt = std::min((size_t)W, n - off);
    for (size_t i = 0; i < cnt; ++i) next[off + i] = -1, tout[off + i] = FLT_MAX;
    if (s.tnodes.empty()) continue;
    Packet<W> rp(rays + off, cnt);
    int sp = 0;
    stack[sp++] = -1;
    int cur = 0;
    int hit[W];
    while (cur >= 0) {
      const TNode &nd = s.tnodes[cur];
      if (!rp.intersect(nd.bbox, hit, false, false)) {
        cur = stack[--sp];
        continue;
      }
      if (nd.numInstances > 0) {
        for (int i = nd.instanceSetIdx; i < nd.instanceSetIdx + nd.numInstances; ++i) {
          if (from == s.sortedID[i]) continue;
          int h2[W];
          if (rp.intersect(s.sortedBB[i], h2, true, enter_inside)) {
            for (int o = 0; o < W; ++o)
              if (h2[o] == 1 && rp.mask[o] == 1 && tout[off + o] > rp.t[o]) {
                next[off + o] = s.sortedID[i];
                tout[off + o] = rp.t[o];
              }
          }
        }
        cur = stack[--sp];
      } else {
        stack[sp++] = nd.right;
        cur = nd.left;
      }
    }
  }
}

void route(const Scene &s, const Ray80 *rays, size_t n, int from, int simd, int *next, float *t) {
  bool ei = (s.flags & FLAG_ENTER_INSIDE) && from == -1;
  if (simd == 1)
    routeT<1>(s, rays, n, from, ei, next, t);
  else
    routeT<8>(s, rays, n, from, ei, next, t); // GVT_SIMD_WIDTH 8 (CMake/gravit.cmake:49-50)
}

// ---- counter-based RNG (replaces the per-TBB-chunk stateful generator, Q14) -----------------------
inline uint32_t pcg(uint32_t v) {
  uint32_t s = v * 747796405u + 2891336453u;
  uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
  return (w >> 22u) ^ w;
}
inline float rnd01(uint32_t seed, int id, int depth, int dim) {
  uint32_t h = pcg(seed ^ pcg((uint32_t)id * 4u + 0x9e3779b9u * (uint32_t)(depth + 1)) ^ (0x85ebca6bu * (uint32_t)(dim + 1)));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// ---- shading: data/primitives/Material.cpp:50-139, data/scene/Light.cpp:58-62 ----------------------
inline float maxf(float a, float b) { return (a < b) ? b : a; } // std::max(a,b)
inline V3 lightContribution(const Light &L, V3 hit, V3 samplePos) {
  if (L.type == 1) return L.color; // AmbientLight, Light.cpp:72
  V3 p = (L.type == 2) ? samplePos : L.pos;
  float distance = 1.f / length(p - hit);
  distance = (distance > 1.f) ? 1.f : distance;
  return L.color * distance;
}
inline float clamp01(float x) {
  float a = (x < 0.f) ? 0.f : x; // glm::max(x, 0)
  return (1.f < a) ? 1.f : a;    // glm::min(a, 1)
}
bool Shade(const Material &mat, const Ray80 &ray, V3 N, const Light &L, V3 lightPos, V3 &color) {
  V3 hitPoint = ray.o + ray.d * ray.t;
  V3 wi = normalize(lightPos - hitPoint);
  float NdotL = maxf(0.f, dot(N, wi));
  V3 Li = lightContribution(L, hitPoint, lightPos);
  if (NdotL == 0.f || (Li.x == 0.f && Li.y == 0.f && Li.z == 0.f)) return false;
  switch (mat.type) {
  case MAT_LAMBERT: color = mat.kd * (NdotL * ray.w); break;
  case MAT_PHONG: {
    V3 R = ((N * 2.f) * NdotL) - wi;
    float VdotR = maxf(0.f, dot(R, -ray.d));
    float power = VdotR * powf(VdotR, mat.alpha);
    color = mat.kd * (NdotL * ray.w);
    color = color + mat.ks * (power * ray.w);
  } break;
  case MAT_BLINN: {
    V3 Hh = normalize(wi - ray.d);
    float NdotH = maxf(0.f, dot(Hh, N));
    float power = NdotH * powf(NdotH, mat.alpha);
    V3 diffuse = mat.kd * (NdotL * ray.w);
    V3 specular = mat.ks * (power * ray.w);
    color = diffuse + specular;
  } break;
  default: color = v3(0, 0, 0);
  }
  color = color * Li;
  color = v3(clamp01(color.x), clamp01(color.y), clamp01(color.z));
  return true;
}

// EmbreeMeshAdapter.cpp:289-318 with the counter-based RNG; float math (documented: PSNR-level parity)
V3 cosHemisphere(V3 n, float Xi1, float Xi2) {
  float theta = (float)std::acos(std::sqrt(1.0 - (double)Xi1));
  float phi = (float)(2.0 * 3.1415926535897932384626433832795 * (double)Xi2);
  float xs = sinf(theta) * cosf(phi), ys = cosf(theta), zs = sinf(theta) * sinf(phi);
  V3 y = n, h = y;
  if (fabsf(h.x) <= fabsf(h.y) && fabsf(h.x) <= fabsf(h.z))
    h.x = 1.0f;
  else if (fabsf(h.y) <= fabsf(h.x) && fabsf(h.y) <= fabsf(h.z))
    h.y = 1.0f;
  else
    h.z = 1.0f;
  V3 x = cross(h, y), z = cross(x, y);
  V3 dir = x * xs + y * ys + z * zs;
  return normalize(dir);
}

// ---- the adapter: EmbreeMeshAdapter.cpp:320-660 ---------------------------------------------------
struct TraceOut {
  std::vector<Ray80> moved;
  long long hits = 0, shadow_gen = 0, shadow_ok = 0, secondary = 0;
};

inline void toLocal(const M4 &minv, V3 o, V3 d, V3 &lo, V3 &ld) {
  // o' = minv * (o,1), d' = minv * (d,0) in GLM order: (c0*x + c1*y) + (c2*z + c3*w); for w=0 the last
  // term is dropped: (c0*x + c1*y) + c2*z.  t is unchanged because d' is not renormalised.
  for (int r = 0; r < 3; ++r) {
    float a = minv.c[0][r] * o.x + minv.c[1][r] * o.y;
    float b = minv.c[2][r] * o.z + minv.c[3][r];
    lo[r] = a + b;
    float c = minv.c[0][r] * d.x + minv.c[1][r] * d.y;
    ld[r] = c + minv.c[2][r] * d.z;
  }
}

void traceRange(const Scene &s, int instId, Ray80 *rays, size_t begin, size_t end, TraceOut &out, HitRec *hits) {
  const Instance &in = s.inst[instId];
  const Mesh &mesh = *s.meshes[in.mesh];
  const bool brute = s.flags & FLAG_BRUTE;
  const bool smooth = s.flags & FLAG_SMOOTH;
  // documented deviation: occlusion is tested up to the light (t_max = |light - origin|, what
  // EmbreeMeshAdapter.cpp:347 meant to compute) instead of to infinity; off by default = reference behaviour
  const bool to_light = s.flags & FLAG_SHADOW_TO_LIGHT;
  for (size_t ri = begin; ri < end; ++ri) {
    Ray80 &r = rays[ri];
    bool first = true;
    for (;;) { // while(validRayLeft) loop of one lane, EmbreeMeshAdapter.cpp:459-615
      V3 lo, ld;
      toLocal(in.minv, r.o, r.d, lo, ld);
      // prepGVT_EMBREE_PACKET_TYPE :269-286: tnear=RAY_EPSILON, tfar=FLT_MAX, ray's own t_min/t_max ignored
      LocalHit h = intersectClosest(mesh, lo, ld, RAY_EPSILON, (to_light && r.type == RAY_SHADOW) ? r.t_max : FLT_MAX, brute);
      if (first && hits) {
        hits[ri].prim = h.prim;
        hits[ri].t = h.prim >= 0 ? h.t : FLT_MAX;
        hits[ri].u = h.u, hits[ri].v = h.v;
      }
      first = false;
      if (h.prim < 0) { // miss -> localDispatch :604-609
        out.moved.push_back(r);
        break;
      }
      if (r.type == RAY_SHADOW) break; // :487-489 dropped
      out.hits++;
      float t = h.t;
      r.t = t;
      int I = mesh.faces[3 * h.prim], J = mesh.faces[3 * h.prim + 1], K = mesh.faces[3 * h.prim + 2];
      V3 v0 = mesh.verts[I], v1 = mesh.verts[J], v2 = mesh.verts[K];
      // :504 normalflat = normalize(normi * -Ng); -Ng_embree = cross(v1-v0, v2-v0) (pinned by goldens, F5)
      V3 normalflat = normalize(m3_mul_v3(in.normi, cross(v1 - v0, v2 - v0)));
      V3 manualNormal;
      if (smooth) { // :507-519
        V3 a = mesh.normals[J], b = mesh.normals[K], c = mesh.normals[I];
        manualNormal = a * h.u + b * h.v + c * (1.0f - h.u - h.v);
        manualNormal = normalize(m3_mul_v3(in.normi, manualNormal));
      } else
        manualNormal = normalflat; // FLAT_SHADING :75,521-523
      if (dot(-r.d, normalflat) <= 0.f) manualNormal = -manualNormal; // :528-530
      Material mat;
      if (!mesh.vcolors.empty()) { // :536-563
        V3 c0 = mesh.vcolors[I], c1 = mesh.vcolors[J], c2 = mesh.vcolors[K];
        mat.type = MAT_LAMBERT;
        mat.kd = (c0 * (1.f - h.u - h.v)) + (c1 * h.u) + (c2 * h.v);
        mat.ks = v3(.5f, .5f, .5f);
        mat.alpha = 1.f;
      } else if (!mesh.face_mat.empty() && mesh.face_mat[h.prim] >= 0) // :565-566
        mat = mesh.mats[mesh.face_mat[h.prim]];
      else
        mat = mesh.mat; // :568
      if (r.type == RAY_SECONDARY) { // :572-575
        float tt = (t > 1) ? 1.f / t : t;
        r.w = r.w * tt;
      }
      // generateShadowRays :320-358 followed by traceShadowRays :364-385 (same instance, tfar=FLT_MAX)
      int li = 0;
      for (const Light &L : s.lights) {
        V3 lightPos = L.pos;
        if (L.type == 2) { // AreaLight::GetPosition Light.cpp:115-127, counter-based RNG
          float xl = (rnd01(s.seed, r.id, r.depth, 8 + 2 * li) - 0.5f) * L.width;
          float zl = (rnd01(s.seed, r.id, r.depth, 9 + 2 * li) - 0.5f) * L.height;
          V3 vv = L.normal, up = v3(0, 1, 0), uu, ww;
          if (vv.x == up.x && vv.y == up.y && vv.z == up.z)
            uu = v3(1, 0, 0), ww = v3(0, 0, 1);
          else {
            uu = v3(up.y * vv.z - vv.y * up.z, up.z * vv.x - vv.z * up.x, up.x * vv.y - vv.x * up.y);
            ww = v3(vv.y * uu.z - uu.y * vv.z, vv.z * uu.x - uu.z * vv.x, vv.x * uu.y - uu.x * vv.y);
          }
          lightPos = v3(L.pos.x + (xl * uu.x + zl * ww.x), L.pos.y + (xl * uu.y + zl * ww.y), L.pos.z + (xl * uu.z + zl * ww.z));
        }
        ++li;
        V3 c;
        if (!Shade(mat, r, manualNormal, L, lightPos, c)) continue;
        const float multiplier = 1.0f - RAY_EPSILON * 16;
        const float t_shadow = multiplier * r.t;
        V3 origin = r.o + r.d * t_shadow;
        V3 dir = lightPos - origin;
        Ray80 sh;
        memset(&sh, 0, sizeof(sh));
        sh.o = origin;
        sh.d = normalize(dir); // Ray ctor actor/Ray.h:109
        sh.t_min = RAY_EPSILON;
        sh.t_max = to_light ? length(dir) : 3.0f; // glm::vec3::length() == 3 (Q2), EmbreeMeshAdapter.cpp:347,355
        sh.t = r.t;
        sh.id = r.id;
        sh.w = r.w;
        sh.type = RAY_SHADOW;
        sh.depth = r.depth; // reference leaves it unset (Q1); fixed here for determinism
        sh.color = c;
        out.shadow_gen++;
        V3 slo, sld;
        toLocal(in.minv, sh.o, sh.d, slo, sld);
        if (!intersectAny(mesh, slo, sld, RAY_EPSILON, to_light ? sh.t_max : FLT_MAX, brute)) {
          out.shadow_ok++;
          out.moved.push_back(sh);
        }
      }
      int ndepth = r.depth - 1;
      float p = 1.f - rnd01(s.seed, r.id, r.depth, 0);
      if (ndepth > 0 && r.w > p) { // :588-598
        r.type = RAY_SECONDARY;
        const float multiplier = 1.0f - 16.0f * FLT_EPSILON;
        const float t_secondary = multiplier * r.t;
        r.o = r.o + r.d * t_secondary;
        r.d = cosHemisphere(manualNormal, rnd01(s.seed, r.id, r.depth, 1), rnd01(s.seed, r.id, r.depth, 2));
        r.w = r.w * dot(r.d, manualNormal);
        r.depth = ndepth;
        out.secondary++;
        continue;
      }
      break;
    }
  }
}

// parallel_for over chunks; results merged in chunk order (deterministic, unlike the TBB original)
template <typename F> void parallelChunks(size_t n, size_t chunk, int threads, F f) {
  size_t nchunks = (n + chunk - 1) / chunk;
  if (threads <= 1 || nchunks <= 1) {
    for (size_t c = 0; c < nchunks; ++c) f(c, c * chunk, std::min(n, (c + 1) * chunk));
    return;
  }
  std::atomic<size_t> next(0);
  std::vector<std::thread> th;
  int nt = (int)std::min<size_t>(threads, nchunks);
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&]() {
      for (;;) {
        size_t c = next.fetch_add(1);
        if (c >= nchunks) break;
        f(c, c * chunk, std::min(n, (c + 1) * chunk));
      }
    });
  for (auto &t : th) t.join();
}

size_t traceQueue(const Scene &s, int instId, std::vector<Ray80> &q, std::vector<Ray80> &moved, int threads, Stats *st,
                  HitRec *hits) {
  size_t n = q.size();
  size_t chunk = std::max<size_t>(4096, n / (size_t)(std::max(1, threads) * 2)); // EmbreeMeshAdapter.cpp:648
  size_t nchunks = (n + chunk - 1) / chunk;
  std::vector<TraceOut> outs(nchunks);
  parallelChunks(n, chunk, threads,
                 [&](size_t c, size_t b, size_t e) { traceRange(s, instId, q.data(), b, e, outs[c], hits); });
  for (auto &o : outs) {
    moved.insert(moved.end(), o.moved.begin(), o.moved.end());
    if (st) {
      st->primary_hits += o.hits;
      st->shadow_generated += o.shadow_gen;
      st->shadow_unoccluded += o.shadow_ok;
      st->secondary_generated += o.secondary;
    }
  }
  return moved.size();
}

// composite/IceTComposite.cpp:111-117
inline void localAdd(float *fb, size_t idx, V3 c, float alpha) {
  for (int i = 0; i < 3; ++i) {
    fb[idx * 4 + i] += c[i];
    if (fb[idx * 4 + i] > 1.f) fb[idx * 4 + i] = 1.f;
  }
  fb[idx * 4 + 3] += alpha;
}

// algorithm/TracerBase.h:325-414 (non-volume branch :392-401).  keep(next) filters by ownership for
// Tracer<DomainScheduler>::shuffleDropRays (DomainTracer.h:148-181) when drop_nonlocal is set.
void shuffleRays(const Scene &s, std::vector<Ray80> &rays, int domID, std::map<int, std::vector<Ray80> > &queue, float *fb,
                 int threads, Stats *st, bool drop_mode, int rank) {
  size_t n = rays.size();
  std::vector<int> next(n);
  std::vector<float> tt(n);
  size_t chunk = std::max<size_t>(4096, n / (size_t)(std::max(1, threads) * 4));
  chunk = (chunk + 7) & ~(size_t)7; // keep packet boundaries independent of the thread count
  parallelChunks(n, chunk, threads,
                 [&](size_t, size_t b, size_t e) { route(s, rays.data() + b, e - b, domID, 8, next.data() + b, tt.data() + b); });
  for (size_t i = 0; i < n; ++i) {
    Ray80 &r = rays[i];
    if (next[i] != -1) {
      r.o = r.o + r.d * (tt[i] * 0.95f); // TracerBase.h:394
      // FLAG_SHADOW_TO_LIGHT (documented deviation): t_max is the distance left to the light, so it shrinks with the origin
      // advance; otherwise occluders behind the light in a later instance would shadow the pixel
      if ((s.flags & FLAG_SHADOW_TO_LIGHT) && r.type == RAY_SHADOW) r.t_max = r.t_max - tt[i] * 0.95f;
      if (drop_mode && s.owner[next[i]] != rank) continue;
      queue[next[i]].push_back(r);
    } else if (!drop_mode && r.type == RAY_SHADOW && length(r.color) > 0) { // :396-399
      localAdd(fb, (size_t)r.id, r.color * r.w, 1.f);
      if (st) st->fb_adds++;
    }
  }
  rays.clear();
}

// gvtCamera.cpp:89-171 (RIGHT_HAND_CAMERA, gvtCamera.h:46) + :233-312
size_t cameraRays(const Scene &s, Ray80 *rays, int y0, int y1) {
  V3 w = normalize(s.focus - s.eye);
  V3 v = normalize(s.up);
  V3 u;
  u.x = w.y * v.z - w.z * v.y;
  u.y = w.z * v.x - w.x * v.z;
  u.z = w.x * v.y - w.y * v.x;
  u = normalize(u);
  V3 upv;
  upv.x = u.y * w.z - w.y * u.z;
  upv.y = u.z * w.x - w.z * u.x;
  upv.z = u.x * w.y - w.x * u.y;
  v = normalize(upv);
  const int bw = s.W, bh = s.H, samples = s.samples;
  float aspectRatio = float(bw) / float(bh);
  const float vert = tanf((float)(s.fov * 0.5));
  const float horz = tanf((float)(s.fov * 0.5)) * aspectRatio;
  const float divider = samples;
  const float offset = (float)((1.0 / divider) * s.jitterWindowSize);
  const float wmult = 2.f / float(bw - 1), hmult = 2.f / float(bh - 1);
  const float half_sample = samples * 0.5f;
  const size_t samples2 = (size_t)samples * samples;
  const float contri = 1.f / (samples * samples);
  size_t count = 0;
  for (int j = y0; j < y1; ++j) {
    int idx = j * bw;
    for (int i = 0; i < bw; ++i) {
      const float x0 = (float)((double)(float(i) * wmult) - 1.0), y0f = (float)((double)(float(j) * hmult) - 1.0);
      for (int k = 0; k < samples; ++k)
        for (int ww = 0; ww < samples; ++ww) {
          size_t ridx = (size_t)(idx - y0 * bw) * samples2 + k * samples + ww;
          float x = x0 + (ww - half_sample) * offset;
          x *= horz;
          float y = y0f + (k - half_sample) * offset;
          y *= vert;
          V3 dir;
          dir.x = u.x * x + v.x * y + w.x;
          dir.y = u.y * x + v.y * y + w.y;
          dir.z = u.z * x + v.z * y + w.z;
          Ray80 &ray = rays[ridx];
          memset(&ray, 0, sizeof(ray));
          ray.id = idx;
          ray.t_min = RAY_EPSILON;
          ray.o = s.eye;
          ray.d = normalize(dir);
          ray.t_max = FLT_MAX;
          ray.color = v3(0, 0, 0);
          ray.t = FLT_MAX;
          ray.w = contri;
          ray.type = RAY_PRIMARY;
          ray.depth = s.depth;
          ++count;
        }
      idx++;
    }
  }
  return count;
}

// algorithm/ImageTracer.h:127-269 (world_size 1) and algorithm/DomainTracer.h:185-368 (world_size P,
// ranks simulated in-process; exchange = SendRays :370-496; termination = :339-349)
void renderFrame(Scene &s, int schedule, int P, int threads, float *fbout, Stats *st) {
  const size_t npix = (size_t)s.W * s.H;
  if (P < 1) P = 1;
  if (schedule == 0) P = std::max(1, P);
  s.route_log.clear();
  s.call_log.clear();
  std::vector<std::vector<float> > fb(P, std::vector<float>(npix * 4, 0.f)); // clearBuffer
  std::vector<std::map<int, std::vector<Ray80> > > queue(P);
  size_t nrays = npix * s.samples * s.samples;
  if (schedule == 1 && (int)s.owner.size() != (int)s.inst.size()) {
    s.owner.resize(s.inst.size());
    for (size_t k = 0; k < s.inst.size(); ++k) s.owner[k] = (int)(k % P);
  }
  {
    std::vector<Ray80> rays(nrays);
    parallelChunks((size_t)s.H, 16, threads, [&](size_t, size_t b, size_t e) {
      cameraRays(s, rays.data() + (size_t)b * s.W * s.samples * s.samples, (int)b, (int)e);
    });
    if (st) st->camera_rays += nrays;
    for (int r = 0; r < P; ++r) {
      std::vector<Ray80> lr;
      if (schedule == 0) { // ImageTracer.h:111-125 contiguous slices, remainder to the last rank
        size_t portion = nrays / P, b = r * portion, e = (r + 1 == P) ? nrays : (r + 1) * portion;
        lr.assign(rays.begin() + b, rays.begin() + e);
        shuffleRays(s, lr, -1, queue[r], fb[r].data(), threads, st, false, r);
      } else { // every rank generates all rays, keeps those entering a local instance first
        lr = rays;
        shuffleRays(s, lr, -1, queue[r], fb[r].data(), threads, nullptr, true, r);
      }
      if (st)
        for (auto &q : queue[r]) st->routed_initial += q.second.size();
    }
  }
  bool all_done = false;
  while (!all_done) {
    for (int r = 0; r < P; ++r) {
      std::vector<Ray80> moved;
      for (;;) {
        int instTarget = -1;
        size_t instTargetCount = 0;
        for (auto &q : queue[r]) { // largest queue first, ties -> lowest id (Q19)
          bool inRank = schedule == 0 || s.owner[q.first] == r;
          if (inRank && q.second.size() > instTargetCount) instTargetCount = q.second.size(), instTarget = q.first;
        }
        if (instTarget < 0) break;
        if (st) st->adapter_calls++, st->rays_traced += instTargetCount;
        s.call_log.push_back(r), s.call_log.push_back(instTarget), s.call_log.push_back((int)instTargetCount);
        traceQueue(s, instTarget, queue[r][instTarget], moved, threads, st, nullptr);
        queue[r][instTarget].clear();
        shuffleRays(s, moved, instTarget, queue[r], fb[r].data(), threads, st, false, r);
      }
    }
    if (schedule == 0 || P == 1) break;
    // SendRays: every non-local queue moves to its owner
    std::vector<long long> mat((size_t)P * P, 0);
    long long not_done = 0;
    std::vector<std::map<int, std::vector<Ray80> > > inbox(P);
    for (int r = 0; r < P; ++r)
      for (auto &q : queue[r]) {
        int n = s.owner[q.first];
        if (n != r && !q.second.empty()) {
          mat[(size_t)r * P + n] += (long long)q.second.size();
          auto &dst = inbox[n][q.first];
          dst.insert(dst.end(), q.second.begin(), q.second.end());
          if (st) st->rays_sent += q.second.size();
          q.second.clear();
        }
      }
    for (int r = 0; r < P; ++r)
      for (auto &q : inbox[r]) {
        auto &dst = queue[r][q.first];
        dst.insert(dst.end(), q.second.begin(), q.second.end());
      }
    for (int r = 0; r < P; ++r)
      for (auto &q : queue[r]) not_done += (long long)q.second.size();
    s.route_log.insert(s.route_log.end(), mat.begin(), mat.end());
    if (st) st->rounds++;
    all_done = (not_done == 0);
  }
  // composite: the reference uses IceT blend across ranks (IceTComposite.cpp:57-63,96).  The parity target for
  // P>1 is the single-process image: sum over ranks, clamp to 1.
  for (size_t i = 0; i < npix * 4; ++i) {
    float a = 0;
    for (int r = 0; r < P; ++r) a += fb[r][i];
    if ((i & 3) != 3 && a > 1.f) a = 1.f;
    fbout[i] = a;
  }
}

// 30-bit Morton code of a triangle centroid (the product's LBVH key; the reference has no Morton codes --
// this restates DESIGN.md "Morton key" so the GPU build can be checked bit for bit).
inline uint32_t expandBits(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

} // namespace

// =============================================================================================
// C interface (ctypes)
// =============================================================================================
OR_API void *gvto_scene_create() { return new Scene(); }
OR_API void gvto_scene_destroy(void *p) { delete (Scene *)p; }

// mode 1: api::addMeshTriangles semantics (api.cpp:173-180 -> Mesh::addFace Mesh.cpp:102-112): 1-based indices,
//         faces with two identical vertex positions dropped.  mode 0: ObjReader path (ObjReader.cpp:192-194): 0-based, no filter.
OR_API int gvto_add_mesh(void *p, const float *verts, unsigned nv, const unsigned *tris, unsigned nt, int mode) {
  Scene &s = *(Scene *)p;
  std::unique_ptr<Mesh> m(new Mesh());
  m->verts.resize(nv);
  for (unsigned i = 0; i < nv; ++i) m->verts[i] = v3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
  for (unsigned i = 0; i < nt; ++i) {
    int a = (int)tris[3 * i], b = (int)tris[3 * i + 1], c = (int)tris[3 * i + 2];
    if (mode == 1) {
      a -= 1, b -= 1, c -= 1;
      if (a < 0 || b < 0 || c < 0 || a >= (int)nv || b >= (int)nv || c >= (int)nv) return -1;
      auto eq = [&](int i0, int i1) {
        return m->verts[i0].x == m->verts[i1].x && m->verts[i0].y == m->verts[i1].y && m->verts[i0].z == m->verts[i1].z;
      };
      if (eq(a, b) || eq(b, c) || eq(c, a)) continue;
    }
    m->faces.push_back(a), m->faces.push_back(b), m->faces.push_back(c);
  }
  m->mat.type = MAT_LAMBERT; // Material.h:62-75 defaults
  m->mat.kd = v3(.5f, .5f, .5f);
  m->mat.ks = v3(.5f, .5f, .5f);
  m->mat.alpha = 1.f;
  m->computeBoundingBox(); // api::finishMesh api.cpp:182-192
  m->generateNormals();
  m->buildBVH();
  s.meshes.push_back(std::move(m));
  return (int)s.meshes.size() - 1;
}
OR_API void gvto_mesh_material(void *p, int mesh, int type, const float *kd, const float *ks, float alpha) {
  Mesh &m = *((Scene *)p)->meshes[mesh];
  m.mat.type = type;
  m.mat.kd = v3(kd[0], kd[1], kd[2]);
  if (ks) m.mat.ks = v3(ks[0], ks[1], ks[2]);
  m.mat.alpha = alpha;
}
// per-face materials: mats = nm x {type, kd[3], ks[3], alpha} (8 floats, type as float), face_mat = index or -1
OR_API void gvto_mesh_face_materials(void *p, int mesh, const float *mats, int nm, const int *face_mat) {
  Mesh &m = *((Scene *)p)->meshes[mesh];
  m.mats.resize(nm);
  for (int i = 0; i < nm; ++i) {
    const float *q = mats + 8 * i;
    m.mats[i].type = (int)q[0];
    m.mats[i].kd = v3(q[1], q[2], q[3]);
    m.mats[i].ks = v3(q[4], q[5], q[6]);
    m.mats[i].alpha = q[7];
  }
  m.face_mat.assign(face_mat, face_mat + m.faces.size() / 3);
}
OR_API void gvto_mesh_vertex_colors(void *p, int mesh, const float *c) {
  Mesh &m = *((Scene *)p)->meshes[mesh];
  m.vcolors.resize(m.verts.size());
  for (size_t i = 0; i < m.verts.size(); ++i) m.vcolors[i] = v3(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
}
OR_API unsigned gvto_mesh_num_faces(void *p, int mesh) { return (unsigned)(((Scene *)p)->meshes[mesh]->faces.size() / 3); }
OR_API void gvto_mesh_get(void *p, int mesh, int *faces, float *normals, float *bbox) {
  Mesh &m = *((Scene *)p)->meshes[mesh];
  if (faces) memcpy(faces, m.faces.data(), m.faces.size() * sizeof(int));
  if (normals) memcpy(normals, m.normals.data(), m.normals.size() * sizeof(V3));
  if (bbox) memcpy(bbox, &m.bbox, 24);
}
// api::addInstance api.cpp:292-322
OR_API int gvto_add_instance(void *p, int mesh, const float *am) {
  Scene &s = *(Scene *)p;
  Instance in;
  in.mesh = mesh;
  memcpy(in.m.c, am, 64); // glm::make_mat4: column-major
  in.minv = m4_inverse(in.m);
  M3 m3;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) m3.c[c][r] = in.m.c[c][r];
  in.normi = m3_transpose(m3_inverse(m3));
  const Box &mb = s.meshes[mesh]->bbox;
  float lo4[4] = { mb.lo.x, mb.lo.y, mb.lo.z, 1.f }, hi4[4] = { mb.hi.x, mb.hi.y, mb.hi.z, 1.f }, il[4], ih[4];
  m4_mul_v4(in.m, lo4, il);
  m4_mul_v4(in.m, hi4, ih);
  for (int i = 0; i < 3; ++i) { // Box3D(vmin,vmax) BBox.cpp:40-46
    in.box.lo[i] = fastmin(il[i], ih[i]);
    in.box.hi[i] = fastmax(il[i], ih[i]);
  }
  in.centroid = in.box.centroid();
  s.inst.push_back(in);
  return (int)s.inst.size() - 1;
}
OR_API void gvto_instance_get(void *p, int i, float *m, float *minv, float *normi, float *box) {
  Instance &in = ((Scene *)p)->inst[i];
  if (m) memcpy(m, in.m.c, 64);
  if (minv) memcpy(minv, in.minv.c, 64);
  if (normi) memcpy(normi, in.normi.c, 36);
  if (box) memcpy(box, &in.box, 24);
}
OR_API void gvto_add_light(void *p, int type, const float *pos, const float *color, const float *normal, float width,
                           float height) {
  Light L;
  memset(&L, 0, sizeof(L));
  L.type = type;
  if (pos) L.pos = v3(pos[0], pos[1], pos[2]);
  L.color = v3(color[0], color[1], color[2]);
  if (normal) L.normal = v3(normal[0], normal[1], normal[2]);
  L.width = width, L.height = height;
  ((Scene *)p)->lights.push_back(L);
}
OR_API void gvto_clear_lights(void *p) { ((Scene *)p)->lights.clear(); }
OR_API void gvto_set_camera(void *p, const float *eye, const float *focus, const float *up, float fov, int depth, int samples,
                            float jitter) {
  Scene &s = *(Scene *)p;
  s.eye = v3(eye[0], eye[1], eye[2]);
  s.focus = v3(focus[0], focus[1], focus[2]);
  s.up = v3(up[0], up[1], up[2]);
  s.fov = fov, s.depth = depth, s.samples = samples;
  s.jitterWindowSize = (int)jitter; // gvtCamera.cpp:202 setJitterWindowSize(int)
}
OR_API void gvto_set_film(void *p, int w, int h) { ((Scene *)p)->W = w, ((Scene *)p)->H = h; }
OR_API void gvto_set_flags(void *p, int flags, unsigned seed) { ((Scene *)p)->flags = flags, ((Scene *)p)->seed = seed; }
OR_API void gvto_set_owner(void *p, const int *owner, int n) { ((Scene *)p)->owner.assign(owner, owner + n); }
// BVH::BVH BVH.cpp:44-67
OR_API void gvto_commit(void *p) {
  Scene &s = *(Scene *)p;
  s.tnodes.clear();
  s.sortedID.clear();
  s.sortedBB.clear();
  s.work.resize(s.inst.size());
  for (size_t i = 0; i < s.inst.size(); ++i) s.work[i] = (int)i; // ids 0..n-1 in creation order (Q18)
  std::vector<int> sorted;
  if (!s.inst.empty()) tlasBuild(s, sorted, 0, (int)s.inst.size());
  s.sortedID = sorted;
  for (int id : sorted) s.sortedBB.push_back(s.inst[id].box);
}
OR_API int gvto_tlas_get(void *p, int *sorted_ids, int *n_nodes) {
  Scene &s = *(Scene *)p;
  if (sorted_ids) memcpy(sorted_ids, s.sortedID.data(), s.sortedID.size() * sizeof(int));
  if (n_nodes) *n_nodes = (int)s.tnodes.size();
  return (int)s.sortedID.size();
}
OR_API size_t gvto_camera_rays(void *p, void *out) {
  Scene &s = *(Scene *)p;
  return cameraRays(s, (Ray80 *)out, 0, s.H);
}
OR_API void gvto_route(void *p, const void *rays, size_t n, int from, int simd, int *next, float *t) {
  route(*(Scene *)p, (const Ray80 *)rays, n, from, simd, next, t);
}
// Adapter::trace contract (Adapter.h:82-84): rays mutated in place, moved appended. Returns n_moved, or
// (size_t)-1 if cap is too small.
OR_API size_t gvto_trace(void *p, int inst, void *rays, size_t n, void *moved, size_t cap, void *hits, int threads) {
  Scene &s = *(Scene *)p;
  std::vector<Ray80> q((Ray80 *)rays, (Ray80 *)rays + n), mv;
  traceQueue(s, inst, q, mv, threads, nullptr, (HitRec *)hits);
  memcpy(rays, q.data(), n * sizeof(Ray80));
  if (mv.size() > cap) return (size_t)-1;
  memcpy(moved, mv.data(), mv.size() * sizeof(Ray80));
  return mv.size();
}
// closest hit of local-space rays against one mesh (o,d from the Ray80 records), BVH or brute force
OR_API void gvto_intersect(void *p, int mesh, const void *rays, size_t n, int brute, void *hits, int threads) {
  Scene &s = *(Scene *)p;
  const Ray80 *r = (const Ray80 *)rays;
  HitRec *h = (HitRec *)hits;
  parallelChunks(n, 1024, threads, [&](size_t, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      LocalHit lh = intersectClosest(*s.meshes[mesh], r[i].o, r[i].d, RAY_EPSILON, FLT_MAX, brute != 0);
      h[i].prim = lh.prim, h[i].t = lh.prim >= 0 ? lh.t : FLT_MAX, h[i].u = lh.u, h[i].v = lh.v;
    }
  });
}
OR_API void gvto_occluded(void *p, int mesh, const void *rays, size_t n, int brute, int *occ, int threads) {
  Scene &s = *(Scene *)p;
  const Ray80 *r = (const Ray80 *)rays;
  parallelChunks(n, 1024, threads, [&](size_t, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) occ[i] = intersectAny(*s.meshes[mesh], r[i].o, r[i].d, RAY_EPSILON, FLT_MAX, brute != 0);
  });
}
OR_API void gvto_render(void *p, int schedule, int world_size, int threads, float *fb, long long *stats11) {
  Stats st;
  memset(&st, 0, sizeof(st));
  renderFrame(*(Scene *)p, schedule, world_size, threads, fb, &st);
  if (stats11) memcpy(stats11, &st, sizeof(st));
}
OR_API size_t gvto_route_log(void *p, long long *out, size_t cap) {
  Scene &s = *(Scene *)p;
  if (out) memcpy(out, s.route_log.data(), std::min(cap, s.route_log.size()) * sizeof(long long));
  return s.route_log.size();
}
OR_API size_t gvto_call_log(void *p, int *out, size_t cap) {
  Scene &s = *(Scene *)p;
  if (out) memcpy(out, s.call_log.data(), std::min(cap, s.call_log.size()) * sizeof(int));
  return s.call_log.size();
}
// Shade() alone, for cross-checks against oracle/_ref
OR_API int gvto_shade(int mtype, const float *kd, const float *ks, float alpha, const void *ray, const float *N, int ltype,
                      const float *lpos, const float *lcolor, float *out) {
  Material m;
  m.type = mtype, m.kd = v3(kd[0], kd[1], kd[2]), m.ks = v3(ks[0], ks[1], ks[2]), m.alpha = alpha;
  Light L;
  memset(&L, 0, sizeof(L));
  L.type = ltype, L.pos = v3(lpos[0], lpos[1], lpos[2]), L.color = v3(lcolor[0], lcolor[1], lcolor[2]);
  V3 c = v3(0, 0, 0);
  bool ok = Shade(m, *(const Ray80 *)ray, v3(N[0], N[1], N[2]), L, L.pos, c);
  out[0] = c.x, out[1] = c.y, out[2] = c.z;
  return ok;
}
// Box3D::merge / surfaceArea / wideRangingBoxDir (BBox.cpp:55-63,118-133) alone, for cross-checks against oracle/_ref
OR_API float gvto_box_merge_area(const float *a6, const float *b6, float *out6, int *widest) {
  Box a, b;
  a.lo = v3(a6[0], a6[1], a6[2]), a.hi = v3(a6[3], a6[4], a6[5]);
  b.lo = v3(b6[0], b6[1], b6[2]), b.hi = v3(b6[3], b6[4], b6[5]);
  a.merge(b);
  out6[0] = a.lo.x, out6[1] = a.lo.y, out6[2] = a.lo.z, out6[3] = a.hi.x, out6[4] = a.hi.y, out6[5] = a.hi.z;
  *widest = a.wideRangingBoxDir();
  return a.surfaceArea();
}
// normalize(normi * v) as the adapter applies it to normals (EmbreeMeshAdapter.cpp:504,518), for the same cross-check
OR_API void gvto_normal(const float *normi9, const float *v3in, float *out3) {
  M3 n;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) n.c[c][r] = normi9[3 * c + r];
  V3 r = normalize(m3_mul_v3(n, v3(v3in[0], v3in[1], v3in[2])));
  out3[0] = r.x, out3[1] = r.y, out3[2] = r.z;
}
// Morton keys: centroid of the triangle's AABB (0.5f*lo + 0.5f*hi), normalised by the mesh bbox
// (c - lo) / (hi - lo) (IEEE division; 0 where the extent is 0), scaled by 1024, clamped to [0,1023], truncated.
OR_API void gvto_morton(void *p, int mesh, uint32_t *codes) {
  Mesh &m = *((Scene *)p)->meshes[mesh];
  size_t nf = m.faces.size() / 3;
  V3 ext = m.bbox.hi - m.bbox.lo;
  for (size_t f = 0; f < nf; ++f) {
    Box b = m.triBox((int)f);
    V3 c = b.lo * 0.5f + b.hi * 0.5f;
    uint32_t q[3];
    for (int a = 0; a < 3; ++a) {
      float x = ext[a] > 0.f ? (c[a] - m.bbox.lo[a]) / ext[a] : 0.f;
      x = x * 1024.f;
      x = fminf(fmaxf(x, 0.f), 1023.f);
      q[a] = (uint32_t)x;
    }
    codes[f] = (expandBits(q[0]) << 2) | (expandBits(q[1]) << 1) | expandBits(q[2]);
  }
}
// IceTComposite::write IceTComposite.cpp:119-158 -- P6, rows bottom-to-top, (unsigned char)(c*255)
OR_API void gvto_fb_to_ppm_bytes(const float *fb, int w, int h, unsigned char *rgb) {
  size_t k = 0;
  for (int j = h - 1; j >= 0; --j)
    for (int i = 0; i < w; ++i) {
      size_t index = 4 * ((size_t)j * w + i);
      rgb[k++] = (unsigned char)(fb[index + 0] * 255);
      rgb[k++] = (unsigned char)(fb[index + 1] * 255);
      rgb[k++] = (unsigned char)(fb[index + 2] * 255);
    }
}
OR_API int gvto_sizeof_ray() { return (int)sizeof(Ray80); }
