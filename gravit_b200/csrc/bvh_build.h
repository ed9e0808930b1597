// bvh_build.h -- per-thread bodies of the GPU BLAS build:
//   tri_setup   : triangle AABB + 30-bit Morton key of its centroid           (one thread per face)
//   [radix sort by key, stable -- kernels.cu]
//   leaf_setup  : sorted leaf boxes                                            (one thread per sorted face)
//   karras_node : binary radix tree over the sorted keys (Karras, HPG 2012)    (one thread per internal node)
//   refit_up    : bottom-up AABB fit with one atomic flag per internal node    (one thread per leaf)
//   collapse    : binary tree -> 8-wide compressed nodes with quantised child boxes, surface-area-greedy child
//                 expansion, octant-ordered slots; triangles re-laid per wide node     (one thread per wide node,
//                 level by level through a task queue)
// Replaces Embree's rtcNewTriangleMesh + rtcCommit (adapter/embree/EmbreeMeshAdapter.cpp:125-162).
#pragma once
#include "gvt_device.h"

namespace gvt {

struct BuildCtx {
  const float *verts;    // 3 per vertex
  const uint32_t *faces; // 3 per face, 0-based
  unsigned n;            // faces
  V3 mlo, mext;          // mesh bbox min and extent (hi - lo)
  float pad;             // absolute padding of leaf boxes (conservative traversal)
  // sort
  uint32_t *keys, *vals; // Morton key and face index (sorted in place by the radix sort)
  // binary tree: node ids 0..n-2 internal, n-1..2n-2 leaves (leaf k = sorted position k)
  float4 *bn_lo, *bn_hi;
  int2 *bn_child;
  int *bn_parent;
  int2 *bn_range;
  unsigned *bn_flag;
  // wide output
  float4 *wnodes; // 5 per node
  float4 *wtris;  // 3 per triangle
  unsigned *node_count, *tri_count;
  unsigned wnode_cap;
};

// absolute padding of the leaf boxes: the canonical triangle test is evaluated in fp32, so a ray can be accepted while
// passing a few ulp (of the ray-origin-to-triangle distance) outside the exact triangle; the boxes must not cull it
GVT_HD float blas_leaf_pad(V3 mext) { return 4e-6f * (mext.x + mext.y + mext.z) + 1e-30f; }

GVT_HD uint32_t expand_bits10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

GVT_HD void tri_verts(const BuildCtx &c, unsigned f, V3 &a, V3 &b, V3 &cc) {
  unsigned i0 = c.faces[3 * f], i1 = c.faces[3 * f + 1], i2 = c.faces[3 * f + 2];
  a = v3(c.verts[3 * i0], c.verts[3 * i0 + 1], c.verts[3 * i0 + 2]);
  b = v3(c.verts[3 * i1], c.verts[3 * i1 + 1], c.verts[3 * i1 + 2]);
  cc = v3(c.verts[3 * i2], c.verts[3 * i2 + 1], c.verts[3 * i2 + 2]);
}

// Morton key: centroid of the triangle's AABB (0.5f*lo + 0.5f*hi), normalised by the mesh bbox with an IEEE
// division, scaled by 1024, clamped to [0,1023], truncated; x in the top bit of each triple.  Bit-exact with
// oracle/gvt_oracle.cpp gvto_morton.
GVT_HD void tri_setup(const BuildCtx &c, unsigned f) {
  V3 a, b, d;
  tri_verts(c, f, a, b, d);
  V3 lo = v3(fminf(fminf(a.x, b.x), d.x), fminf(fminf(a.y, b.y), d.y), fminf(fminf(a.z, b.z), d.z));
  V3 hi = v3(fmaxf(fmaxf(a.x, b.x), d.x), fmaxf(fmaxf(a.y, b.y), d.y), fmaxf(fmaxf(a.z, b.z), d.z));
  V3 cen = xadd(xmul(lo, 0.5f), xmul(hi, 0.5f));
  float nx = c.mext.x > 0.f ? xdiv(xsub(cen.x, c.mlo.x), c.mext.x) : 0.f;
  float ny = c.mext.y > 0.f ? xdiv(xsub(cen.y, c.mlo.y), c.mext.y) : 0.f;
  float nz = c.mext.z > 0.f ? xdiv(xsub(cen.z, c.mlo.z), c.mext.z) : 0.f;
  uint32_t qx = (uint32_t)fminf(fmaxf(xmul(nx, 1024.f), 0.f), 1023.f);
  uint32_t qy = (uint32_t)fminf(fmaxf(xmul(ny, 1024.f), 0.f), 1023.f);
  uint32_t qz = (uint32_t)fminf(fmaxf(xmul(nz, 1024.f), 0.f), 1023.f);
  c.keys[f] = (expand_bits10(qx) << 2) | (expand_bits10(qy) << 1) | expand_bits10(qz);
  c.vals[f] = f;
}

// leaf k (sorted position) -> padded AABB
GVT_HD void leaf_setup(const BuildCtx &c, unsigned k) {
  V3 a, b, d;
  tri_verts(c, c.vals[k], a, b, d);
  float p = c.pad;
  unsigned id = c.n - 1 + k;
  c.bn_lo[id] = make_float4(fminf(fminf(a.x, b.x), d.x) - p, fminf(fminf(a.y, b.y), d.y) - p, fminf(fminf(a.z, b.z), d.z) - p, 0.f);
  c.bn_hi[id] = make_float4(fmaxf(fmaxf(a.x, b.x), d.x) + p, fmaxf(fmaxf(a.y, b.y), d.y) + p, fmaxf(fmaxf(a.z, b.z), d.z) + p, 0.f);
}

// common-prefix length of sorted keys i and j; ties broken by position so that all keys are distinct
GVT_HD int karras_delta(const BuildCtx &c, int i, int j) {
  if (j < 0 || j >= (int)c.n) return -1;
  uint32_t a = c.keys[i], b = c.keys[j];
  if (a != b) return clz32(a ^ b);
  return 32 + clz32((uint32_t)i ^ (uint32_t)j);
}

GVT_HD void karras_node(const BuildCtx &c, int i) {
  const int n = (int)c.n;
  int d = (karras_delta(c, i, i + 1) - karras_delta(c, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = karras_delta(c, i, i - d);
  int lmax = 2;
  while (karras_delta(c, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (karras_delta(c, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = karras_delta(c, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (karras_delta(c, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  int gamma = i + s * d + (d < 0 ? d : 0);
  int lo = i < j ? i : j, hi = i < j ? j : i;
  int left = (lo == gamma) ? (n - 1 + gamma) : gamma;
  int right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
  c.bn_child[i] = make_int2(left, right);
  c.bn_range[i] = make_int2(lo, hi);
  c.bn_parent[left] = i;
  c.bn_parent[right] = i;
  if (i == 0) c.bn_parent[0] = -1;
}

// ---- wide-node collapse ------------------------------------------------------------------------------------
#define GVT_LEAF_MAX 3 /* triangles per leaf child (unary count in 3 meta bits) */

GVT_HD float box_area(float4 lo, float4 hi) {
  float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}
GVT_HD int bnode_count(const BuildCtx &c, int b) {
  if (b >= (int)c.n - 1) return 1;
  int2 r = c.bn_range[b];
  return r.y - r.x + 1;
}
GVT_HD int bnode_first(const BuildCtx &c, int b) {
  if (b >= (int)c.n - 1) return b - ((int)c.n - 1);
  return c.bn_range[b].x;
}

// biased exponent e such that 255 * 2^(e-127) >= ext (smallest such), clamped to [1,254]
GVT_HD uint32_t quant_exponent(float ext) {
  if (!(ext > 0.f)) return 1u;
  float s = ext * (1.0f / 255.0f) * 1.0001f; // a little slack so that ceil() never needs 256
  uint32_t u = f2u(s);
  uint32_t e = (u >> 23) & 0xFFu;
  if (u & 0x7FFFFFu) e += 1u; // round up to the next power of two
  if (e < 1u) e = 1u;
  if (e > 254u) e = 254u;
  return e;
}

struct CollapseTask {
  int bnode; // binary node to turn into a wide node
  int wnode; // index of the wide node to write
};

// Builds wide node t.wnode from binary subtree t.bnode.  Children that stay internal are appended to `out`
// (returns how many).  Host emulation and the device kernel share this body; the device kernel reserves the
// output range with one atomicAdd.
GVT_HD int collapse_node(const BuildCtx &c, CollapseTask t, int *slot_b /*[8]*/, int &n_internal, int &n_leaf_tris,
                         uint32_t *meta /*[8]*/, unsigned *qlo /*[8][3]*/, unsigned *qhi /*[8][3]*/, float4 &w0) {
  // 1. gather up to 8 children: repeatedly open the child with the largest surface area that holds more than
  //    GVT_LEAF_MAX triangles
  int ch[8];
  int nch = 0;
  if (t.bnode >= (int)c.n - 1) {
    ch[nch++] = t.bnode; // single-triangle mesh: the root is a leaf
  } else {
    int2 k = c.bn_child[t.bnode];
    ch[nch++] = k.x;
    ch[nch++] = k.y;
  }
  while (nch < 8) {
    int best = -1;
    float besta = -1.f;
    for (int i = 0; i < nch; ++i) {
      if (bnode_count(c, ch[i]) <= GVT_LEAF_MAX) continue;
      float a = box_area(c.bn_lo[ch[i]], c.bn_hi[ch[i]]);
      if (a > besta) besta = a, best = i;
    }
    if (best < 0) break;
    int2 k = c.bn_child[ch[best]];
    ch[best] = k.x;
    ch[nch++] = k.y;
  }
  // 2. node frame
  float4 lo = c.bn_lo[ch[0]], hi = c.bn_hi[ch[0]];
  for (int i = 1; i < nch; ++i) {
    float4 l = c.bn_lo[ch[i]], h = c.bn_hi[ch[i]];
    lo.x = fminf(lo.x, l.x), lo.y = fminf(lo.y, l.y), lo.z = fminf(lo.z, l.z);
    hi.x = fmaxf(hi.x, h.x), hi.y = fmaxf(hi.y, h.y), hi.z = fmaxf(hi.z, h.z);
  }
  uint32_t ex = quant_exponent(hi.x - lo.x), ey = quant_exponent(hi.y - lo.y), ez = quant_exponent(hi.z - lo.z);
  float isx = 1.0f / u2f(ex << 23), isy = 1.0f / u2f(ey << 23), isz = 1.0f / u2f(ez << 23);
  // 3. slot assignment: slot s bit2/bit1/bit0 = child sits towards +x/+y/+z of the node centre, so that
  //    visiting slots in descending (s ^ (7 - ray octant)) order is front to back.  Greedy max-cost matching.
  float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
  float cost[8][8];
  for (int i = 0; i < nch; ++i) {
    float4 l = c.bn_lo[ch[i]], h = c.bn_hi[ch[i]];
    float dx = 0.5f * (l.x + h.x) - cx, dy = 0.5f * (l.y + h.y) - cy, dz = 0.5f * (l.z + h.z) - cz;
    for (int s = 0; s < 8; ++s) cost[i][s] = ((s & 4) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 1) ? dz : -dz);
  }
  for (int s = 0; s < 8; ++s) slot_b[s] = -1;
  bool used[8] = { false, false, false, false, false, false, false, false };
  for (int k = 0; k < nch; ++k) {
    int bi = -1, bs = -1;
    float bc = -FLT_MAX;
    for (int i = 0; i < nch; ++i) {
      if (used[i]) continue;
      for (int s = 0; s < 8; ++s) {
        if (slot_b[s] >= 0) continue;
        if (cost[i][s] > bc) bc = cost[i][s], bi = i, bs = s;
      }
    }
    used[bi] = true;
    slot_b[bs] = ch[bi];
  }
  // 4. meta + quantised boxes
  uint32_t imask = 0;
  n_internal = 0, n_leaf_tris = 0;
  for (int s = 0; s < 8; ++s) {
    meta[s] = 0;
    qlo[3 * s] = qlo[3 * s + 1] = qlo[3 * s + 2] = 255u; // empty slot: inverted box
    qhi[3 * s] = qhi[3 * s + 1] = qhi[3 * s + 2] = 0u;
    int b = slot_b[s];
    if (b < 0) continue;
    int cnt = bnode_count(c, b);
    if (cnt > GVT_LEAF_MAX) {
      imask |= 1u << s;
      meta[s] = (1u << 5) | (24u + (uint32_t)s);
      n_internal++;
    } else {
      uint32_t unary = (cnt == 1) ? 1u : (cnt == 2 ? 3u : 7u);
      meta[s] = (unary << 5) | (uint32_t)n_leaf_tris;
      n_leaf_tris += cnt;
    }
    float4 l = c.bn_lo[b], h = c.bn_hi[b];
    // conservative: floor/ceil with a 2e-3 grid-step margin, clamped to the byte range
    float fx = floorf((l.x - lo.x) * isx - 2e-3f), fy = floorf((l.y - lo.y) * isy - 2e-3f), fz = floorf((l.z - lo.z) * isz - 2e-3f);
    float gx = ceilf((h.x - lo.x) * isx + 2e-3f), gy = ceilf((h.y - lo.y) * isy + 2e-3f), gz = ceilf((h.z - lo.z) * isz + 2e-3f);
    qlo[3 * s] = (unsigned)fminf(fmaxf(fx, 0.f), 255.f), qlo[3 * s + 1] = (unsigned)fminf(fmaxf(fy, 0.f), 255.f),
            qlo[3 * s + 2] = (unsigned)fminf(fmaxf(fz, 0.f), 255.f);
    qhi[3 * s] = (unsigned)fminf(fmaxf(gx, 0.f), 255.f), qhi[3 * s + 1] = (unsigned)fminf(fmaxf(gy, 0.f), 255.f),
            qhi[3 * s + 2] = (unsigned)fminf(fmaxf(gz, 0.f), 255.f);
  }
  w0 = make_float4(lo.x, lo.y, lo.z, u2f(ex | (ey << 8) | (ez << 16) | (imask << 24)));
  return nch;
}

GVT_HD uint32_t pack4(const unsigned *q, int base, int comp) {
  return (q[3 * (base + 0) + comp] & 0xFFu) | ((q[3 * (base + 1) + comp] & 0xFFu) << 8) | ((q[3 * (base + 2) + comp] & 0xFFu) << 16) |
         ((q[3 * (base + 3) + comp] & 0xFFu) << 24);
}

// writes the node words and this node's triangle records; child tasks go to out[0..n_internal)
GVT_HD void collapse_emit(const BuildCtx &c, CollapseTask t, const int *slot_b, const uint32_t *meta, const unsigned *qlo,
                          const unsigned *qhi, float4 w0, unsigned child_base, unsigned tri_base, CollapseTask *out) {
  float4 *nd = c.wnodes + 5 * (size_t)t.wnode;
  nd[0] = w0;
  uint32_t m0 = meta[0] | (meta[1] << 8) | (meta[2] << 16) | (meta[3] << 24);
  uint32_t m1 = meta[4] | (meta[5] << 8) | (meta[6] << 16) | (meta[7] << 24);
  nd[1] = make_float4(u2f(child_base), u2f(tri_base), u2f(m0), u2f(m1));
  nd[2] = make_float4(u2f(pack4(qlo, 0, 0)), u2f(pack4(qlo, 4, 0)), u2f(pack4(qlo, 0, 1)), u2f(pack4(qlo, 4, 1)));
  nd[3] = make_float4(u2f(pack4(qlo, 0, 2)), u2f(pack4(qlo, 4, 2)), u2f(pack4(qhi, 0, 0)), u2f(pack4(qhi, 4, 0)));
  nd[4] = make_float4(u2f(pack4(qhi, 0, 1)), u2f(pack4(qhi, 4, 1)), u2f(pack4(qhi, 0, 2)), u2f(pack4(qhi, 4, 2)));
  int ki = 0;
  for (int s = 0; s < 8; ++s) {
    int b = slot_b[s];
    if (b < 0) continue;
    if (bnode_count(c, b) > GVT_LEAF_MAX) { // internal child -> next level task
      out[ki].bnode = b;
      out[ki].wnode = (int)child_base + ki;
      ki++;
    } else {
      int first = bnode_first(c, b), cnt = bnode_count(c, b);
      unsigned off = meta[s] & 0x1Fu;
      for (int k = 0; k < cnt; ++k) {
        unsigned f = c.vals[first + k];
        V3 a, bb, d;
        tri_verts(c, f, a, bb, d);
        V3 e1 = xsub(bb, a), e2 = xsub(d, a);
        float4 *tr = c.wtris + 3 * (size_t)(tri_base + off + k);
        tr[0] = make_float4(a.x, a.y, a.z, u2f(f));
        tr[1] = make_float4(e1.x, e1.y, e1.z, 0.f);
        tr[2] = make_float4(e2.x, e2.y, e2.z, 0.f);
      }
    }
  }
}

} // namespace gvt
